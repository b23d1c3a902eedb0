// rapt_bc.cuh -- BounceCenter.advance (rapt/BounceCenter.py:206-251) for an ensemble: one thread per bounce
// centre, DOPRI5 (scipy "dopri5" with its defaults, fresh call per output row as at BounceCenter.py:246-249)
// on the bounce-averaged drift
//     dR/dt = gamma m v^2 / (q S_b B^2)  gradI x B                                  (BounceCenter.py:236-243)
// where every right-hand side traces FIVE field lines through R and four displaced points
// (flutils.halfbouncepath :254-316, flutils.gradI :153-229 -> flutils.eye :65-151 -> fieldline.py:37-105 ->
// rkf.py:13-143) and runs the reference's quadratures on them (rapt_quad.cuh).  Each lane owns a scratch
// curve in HBM ([max_pts][5] + [max_pts][4] doubles); nothing leaves the device between rows.
//
// Reference behaviour kept: the row label is the START time of its step (`np.arange` value,
// BounceCenter.py:248-249), so row 1 repeats t0; the number of rows is len(np.arange(tcur, tcur+delta, dt));
// tcur afterwards is the last label, not the solver time; the solver is re-created by every advance() call.
// The `isequatorial` branch of the reference indexes a 3-vector as a 4-vector (BounceCenter.py:235) and cannot
// run; it is not offered.
#pragma once
#include "rapt_aux.cuh"

namespace RAPT_NS {
using rapt::BCArgs;

#ifndef RAPT_ST_TRACE           /* same value as include/rapt_b200.h (not visible to NVRTC) */
#define RAPT_ST_TRACE (-7)      /* a field line could not be traced / mirror points not bracketed (the reference raises) */
#endif

struct BcCtx {
    double Bm, coef;            // mirror field; gamma m v^2 / q
    double flres, eyestep;
    double *cv, *bw;            // this lane's scratch
    long long cap;
    int quadrature;
    int err;                    // sticky: RAPT_ST_TRACE / RAPT_ST_ROWCAP
};

// The right-hand side is thousands of field evaluations behind a handful of call sites: the big pieces are real
// functions (one copy each), not inlined into the Runge-Kutta stages.
#define RAPT_BC_FN __device__ __noinline__

template <class F>
RAPT_BC_FN long long bc_trace(const FieldP &f, BcCtx &c, double t, double x, double y, double z)
{
    double ds;
    return fieldline_trace<F>(f, t, x, y, z, c.Bm, c.flres, c.cv, c.bw, c.cap, ds);
}

// flutils.eye(tpos, field, Bm) (flutils.py:95-151)
template <class F>
RAPT_BC_FN double bc_eye(const FieldP &f, BcCtx &c, double t, double x, double y, double z)
{
    const long long n = bc_trace<F>(f, c, t, x, y, z);
    if (n > c.cap) { c.err = RAPT_ST_ROWCAP; return nan(""); }
    int e;
    const double I = eye_curve(c.cv, c.bw, n, c.Bm, &e);
    if (e) c.err = RAPT_ST_TRACE;
    return I;
}

// flutils.halfbouncepath(tpos, field, Bm) (flutils.py:274-316)
template <class F>
RAPT_BC_FN double bc_halfbounce(const FieldP &f, BcCtx &c, double t, double x, double y, double z)
{
    const long long n = bc_trace<F>(f, c, t, x, y, z);
    if (n > c.cap) { c.err = RAPT_ST_ROWCAP; return nan(""); }
    const double S = halfbouncepath_curve(c.cv, c.bw, n, c.Bm, c.quadrature);
    if (!(S == S)) c.err = RAPT_ST_TRACE;
    return S;
}

// flutils.gradI(tpos, field, Bm) (flutils.py:183-229)
template <class F>
RAPT_BC_FN void bc_gradI(const FieldP &f, BcCtx &c, double t, double x, double y, double z, double (&g)[3])
{
    const double d = c.eyestep;
    const double sc = fmax(fabs(x), fabs(y));
    double r[3] = {x / sc, y / sc, 0 / sc};
    double b[3]; F::unitb(f, t, x, y, z, b[0], b[1], b[2]);
    const double rb = dot3(r[0], r[1], r[2], b[0], b[1], b[2]);
    r[0] = r[0] - rb * b[0]; r[1] = r[1] - rb * b[1]; r[2] = r[2] - rb * b[2];
    const double rn = sqrt(dot3(r[0], r[1], r[2], r[0], r[1], r[2]));
    r[0] /= rn; r[1] /= rn; r[2] /= rn;
    const double w[3] = {b[1] * r[2] - b[2] * r[1], b[2] * r[0] - b[0] * r[2], b[0] * r[1] - b[1] * r[0]};   // cross(b, r)
    double dI[2];
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        const double *u = k ? w : r;
        const double I1 = bc_eye<F>(f, c, t, x + d * u[0], y + d * u[1], z + d * u[2]);
        const double I2 = bc_eye<F>(f, c, t, x - d * u[0], y - d * u[1], z - d * u[2]);
        // central difference, forward/backward where one side is beyond the mirror field (flutils.py:210-227)
        if (I1 == 0) dI[k] = (bc_eye<F>(f, c, t, x, y, z) - I2) / d;
        else if (I2 == 0) dI[k] = (I1 - bc_eye<F>(f, c, t, x, y, z)) / d;
        else dI[k] = (I1 - I2) / (2 * d);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) g[i] = dI[0] * r[i] + dI[1] * w[i];
}

// BounceCenter.advance.deriv (BounceCenter.py:232-243)
template <class F>
RAPT_BC_FN void bc_rhs(const FieldP &f, BcCtx &c, double t, const double (&Y)[3], double (&out)[3])
{
    double bx, by, bz; F::B(f, t, Y[0], Y[1], Y[2], bx, by, bz);
    const double magBsq = dot3(bx, by, bz, bx, by, bz);
    const double Sb = bc_halfbounce<F>(f, c, t, Y[0], Y[1], Y[2]);
    double g[3]; bc_gradI<F>(f, c, t, Y[0], Y[1], Y[2], g);
    const double s = c.coef / (Sb * magBsq);
    out[0] = s * (g[1] * bz - g[2] * by);
    out[1] = s * (g[2] * bx - g[0] * bz);
    out[2] = s * (g[0] * by - g[1] * bx);
}

// hinit of dopri5 (iord = 5), Hairer's HINIT as driven by scipy
template <class F>
RAPT_DEV double bc_hinit(const FieldP &f, BcCtx &c, double x, const double (&y)[3], double posneg,
                         const double (&f0)[3], double hmax, double atol, double rtol)
{
    double dnf = 0, dny = 0, y1[3], f1[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double sk = atol + rtol * fabs(y[i]);
        dnf += (f0[i] / sk) * (f0[i] / sk);
        dny += (y[i] / sk) * (y[i] / sk);
    }
    double h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
    h = fmin(h, hmax);
    h = copysign(h, posneg);
#pragma unroll
    for (int i = 0; i < 3; i++) y1[i] = y[i] + h * f0[i];
    bc_rhs<F>(f, c, x + h, y1, f1);
    double der2 = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double sk = atol + rtol * fabs(y[i]);
        der2 += ((f1[i] - f0[i]) / sk) * ((f1[i] - f0[i]) / sk);
    }
    der2 = sqrt(der2) / h;
    const double der12 = fmax(fabs(der2), sqrt(dnf));
    const double h1 = (der12 <= 1e-15) ? fmax(1e-6, fabs(h) * 1e-3) : pow(0.01 / der12, 1.0 / 5);
    h = fmin(fmin(100 * fabs(h), h1), hmax);
    return copysign(h, posneg);
}

// one r.integrate(r.t + dt) of scipy's dopri5 (defaults: nsteps 500, safety 0.9, ifactor 10, dfactor 0.2, beta 0 ->
// 0.04).  Returns idid; cnt = (nfcn, nstep, naccpt, nrejct) accumulated.
template <class F>
RAPT_DEV int bc_dopri5(const FieldP &f, BcCtx &c, double &x, double (&y)[3], double xend, double rtol, double atol, int (&cnt)[4])
{
    const double beta = 0.04, safe = 0.9, fac1 = 0.2, fac2 = 10.0, uround = 2.3e-16;
    const int nmax = 500;
    double facold = 1e-4;
    const double expo1 = 0.2 - beta * 0.75, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const double posneg = copysign(1.0, xend - x), hmax = fabs(xend - x);
    double k1[3], k2[3], k3[3], k4[3], k5[3], k6[3], y1[3], ysti[3];
    bool last = false, reject = false;
    int nstep = 0, naccpt = 0, nrejct = 0, nfcn = 0, idid;
    bc_rhs<F>(f, c, x, y, k1);
    double h = bc_hinit<F>(f, c, x, y, posneg, k1, hmax, atol, rtol);
    nfcn += 2;
    for (;;) {
        if (c.err) { idid = c.err; break; }
        if (nstep > nmax) { idid = RAPT_ST_NMAX; break; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { idid = RAPT_ST_HSMALL; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
        nstep++;
#pragma unroll
        for (int i = 0; i < 3; i++) y1[i] = y[i] + h * D5_A2_1 * k1[i];
        bc_rhs<F>(f, c, x + D5_C2 * h, y1, k2);
#pragma unroll
        for (int i = 0; i < 3; i++) y1[i] = y[i] + h * (D5_A3_1 * k1[i] + D5_A3_2 * k2[i]);
        bc_rhs<F>(f, c, x + D5_C3 * h, y1, k3);
#pragma unroll
        for (int i = 0; i < 3; i++) y1[i] = y[i] + h * (D5_A4_1 * k1[i] + D5_A4_2 * k2[i] + D5_A4_3 * k3[i]);
        bc_rhs<F>(f, c, x + D5_C4 * h, y1, k4);
#pragma unroll
        for (int i = 0; i < 3; i++) y1[i] = y[i] + h * (D5_A5_1 * k1[i] + D5_A5_2 * k2[i] + D5_A5_3 * k3[i] + D5_A5_4 * k4[i]);
        bc_rhs<F>(f, c, x + D5_C5 * h, y1, k5);
#pragma unroll
        for (int i = 0; i < 3; i++) ysti[i] = y[i] + h * (D5_A6_1 * k1[i] + D5_A6_2 * k2[i] + D5_A6_3 * k3[i] + D5_A6_4 * k4[i] + D5_A6_5 * k5[i]);
        const double xph = x + h;
        bc_rhs<F>(f, c, xph, ysti, k6);
#pragma unroll
        for (int i = 0; i < 3; i++) y1[i] = y[i] + h * (D5_A7_1 * k1[i] + D5_A7_3 * k3[i] + D5_A7_4 * k4[i] + D5_A7_5 * k5[i] + D5_A7_6 * k6[i]);
        bc_rhs<F>(f, c, xph, y1, k2);
        nfcn += 6;
        double err = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double e = (D5_E1 * k1[i] + D5_E3 * k3[i] + D5_E4 * k4[i] + D5_E5 * k5[i] + D5_E6 * k6[i] + D5_E7 * k2[i]) * h;
            const double sk = atol + rtol * fmax(fabs(y[i]), fabs(y1[i]));
            err += (e / sk) * (e / sk);
        }
        err = sqrt(err / 3);
        const double fac11 = pow(err, expo1);
        double fac = fac11 / pow(facold, beta);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0) {
            facold = fmax(err, 1e-4);
            naccpt++;
#pragma unroll
            for (int i = 0; i < 3; i++) { k1[i] = k2[i]; y[i] = y1[i]; }
            x = xph;
            if (last) { idid = RAPT_ST_OK; break; }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = false;
        } else {
            // NaN error norm (a failed trace) lands here too; c.err ends the call at the top of the loop
            hnew = h / fmin(facc1, fac11 / safe);
            reject = true;
            if (naccpt >= 1) nrejct++;
            last = false;
        }
        h = hnew;
    }
    cnt[0] += nfcn; cnt[1] += nstep; cnt[2] += naccpt; cnt[3] += nrejct;
    return idid;
}

// op 0: BounceCenter.advance.  op 1: the pieces at given points -- out[i] = (S_b, I, gradI[3], deriv[3]) -- so
// that flutils.halfbouncepath / eye / gradI and the right-hand side are testable on their own.  op 2: out[i] = I only.
template <class F>
__global__ void __launch_bounds__(64) k_bounce_center(const BCArgs a)
{
    const long long lane = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nlanes = (long long)gridDim.x * blockDim.x;
    BcCtx c;
    c.flres = a.flres; c.eyestep = a.eyestep; c.cap = a.max_pts; c.quadrature = a.quadrature;
    c.cv = a.curve + (size_t)lane * a.max_pts * 5;
    c.bw = a.scratch + (size_t)lane * a.max_pts * 4;
    for (long long w = lane; w < a.n; w += nlanes) {
        const long long i = a.order ? a.order[w] : w;
        const double v = a.v[i], mass = a.mass[i], q = a.charge[i];
        const double vc = v / RAPT_C_LIGHT;
        const double gamma = 1.0 / sqrt(1 - vc * vc);                              // BounceCenter.py:226
        c.Bm = a.Bm ? a.Bm[i] : mass * (gamma * gamma) * (v * v) / (2 * a.mu[i]);  // :227
        c.coef = gamma * mass * v * v / q;
        c.err = 0;
        double x = a.t[i], Y[3] = {a.x[i], a.y[i], a.z[i]};
        if (a.op == 2) {                       // flutils.eye only (GuidingCenter.geteye: one value per trajectory row)
            a.out[i] = bc_eye<F>(a.f, c, x, Y[0], Y[1], Y[2]);
            a.status[i] = c.err ? c.err : RAPT_ST_OK;
            continue;
        }
        if (a.op == 1) {
            double *o = a.out + 8 * i, g[3], dv[3];
            o[0] = bc_halfbounce<F>(a.f, c, x, Y[0], Y[1], Y[2]);
            o[1] = bc_eye<F>(a.f, c, x, Y[0], Y[1], Y[2]);
            bc_gradI<F>(a.f, c, x, Y[0], Y[1], Y[2], g);
            bc_rhs<F>(a.f, c, x, Y, dv);
            o[2] = g[0]; o[3] = g[1]; o[4] = g[2]; o[5] = dv[0]; o[6] = dv[1]; o[7] = dv[2];
            a.status[i] = c.err ? c.err : RAPT_ST_OK;
            continue;
        }
        // dt = BCtimestep * bounceperiod(last row) (BounceCenter.py:228-229), unless the caller fixed it
        double dt;
        if (a.dtin) dt = a.dtin[i];
        else dt = a.bctimestep * ((2 / v) * bc_halfbounce<F>(a.f, c, x, Y[0], Y[1], Y[2]));
        if (a.dt_out) a.dt_out[i] = dt;
        int cnt[4] = {0, 0, 0, 0}, st = RAPT_ST_OK;
        long long nrows = 0, nstored = 0;
        double tlabel = x;
        if (c.err || !(dt > 0)) st = c.err ? c.err : RAPT_ST_TRACE;
        else {
            // for t in np.arange(tcur, tcur + delta, dt): numpy's length and fill rule (value_k = start + k * ((start + dt) - start))
            const double t0 = x;
            const long long len = (long long)ceil(((t0 + a.delta) - t0) / dt);
            const double dlt = (t0 + dt) - t0;
            for (long long k = 0; k < len; k++) {
                const int idid = bc_dopri5<F>(a.f, c, x, Y, x + dt, a.rtol, a.atol, cnt);
                if (idid != RAPT_ST_OK) { st = idid; break; }
                tlabel = t0 + (double)k * dlt;
                nrows++;
                if (a.rows && a.store_every > 0 && (nrows % a.store_every) == 0 && nstored < a.max_rows) {
                    double *r = a.rows + ((size_t)i * a.max_rows + nstored) * 4;
                    r[0] = tlabel; r[1] = Y[0]; r[2] = Y[1]; r[3] = Y[2];
                    nstored++;
                }
            }
        }
        a.t[i] = tlabel; a.x[i] = Y[0]; a.y[i] = Y[1]; a.z[i] = Y[2];
        a.nrows[i] = (int)nrows; a.nstored[i] = (int)nstored; a.status[i] = st;
        if (a.tsolver) a.tsolver[i] = x;
#pragma unroll
        for (int k = 0; k < 4; k++) a.counters[4 * i + k] = cnt[k];
    }
}

}  // namespace RAPT_NS
