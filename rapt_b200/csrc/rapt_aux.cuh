// rapt_aux.cuh -- the per-call / per-switch pieces of the path as small kernels:
//   batched _Field operators (fields.py:43-280), GuidingCenter.__init__ (mu, p_par),
//   the Particle <-> GuidingCenter switch transforms (GuidingCenter.py:168-186 + utils.py:251-326,
//   Particle.py:149-164 + utils.py:376-433), the isadiabatic predicates, and the bounce-period set-up
//   (GuidingCenter.py:593-606, fieldline.py:13-105, rkf.py:13-143).
#pragma once
#include "rapt_fields.cuh"
#include "rapt_particle.cuh"
#include "rapt_gc.cuh"
#include "rapt_quad.cuh"

namespace RAPT_NS {
using rapt::OpsArgs;
using rapt::MiscArgs;
using rapt::BounceArgs;
using rapt::AdaptArgs;

// utils.magnetic_moment, utils.py:214-216
template <class F>
RAPT_DEV double magnetic_moment(const FieldP &f, double t, double x, double y, double z, double vpar, double v, double mass)
{
    double vc = v / RAPT_C_LIGHT;
    double gamma = 1.0 / sqrt(1 - vc * vc);
    double Bmag = F::magB(f, t, x, y, z);
    return gamma * gamma * mass * (v - vpar) * (v + vpar) / (2 * Bmag);
}

// GuidingCenter.__init__ :124-133
template <class F>
RAPT_DEV void gc_construct(const FieldP &f, double t0, double x, double y, double z, double v, double pa_deg, bool use_pa,
                           double ppar_in, double mass, double &ppar, double &mu)
{
    double vc = v / RAPT_C_LIGHT;
    double gamma = 1 / sqrt(1 - vc * vc);
    double pp = ppar_in;
    if (use_pa) {
        double vpar = (pa_deg == 90) ? 0.0 : v * cos(pa_deg * RAPT_PI / 180);
        pp = gamma * mass * vpar;
    }
    mu = magnetic_moment<F>(f, t0, x, y, z, pp / (mass * gamma), v, mass);
    ppar = pp;
}

// utils.guidingcenter.gyrovector, utils.py:298-303
template <class F>
RAPT_DEV void gyrovector(const FieldP &f, double t, const double (&r)[3], const double (&v)[3], double mass, double q,
                         double (&o)[3])
{
    double vsq = dot3(v[0], v[1], v[2], v[0], v[1], v[2]);
    double gamma = 1 / sqrt(1 - vsq / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    double bx, by, bz; F::B(f, t, r[0], r[1], r[2], bx, by, bz);
    double Bsq = dot3(bx, by, bz, bx, by, bz);
    double cx = by * v[2] - bz * v[1], cy = bz * v[0] - bx * v[2], cz = bx * v[1] - by * v[0];
    double s = gamma * mass / (q * Bsq);
    o[0] = s * cx; o[1] = s * cy; o[2] = s * cz;
}

// GuidingCenter.init(Particle) :168-186 with utils.guidingcenter :251-326 (tol 1e-3, <= 20 iterations)
template <class F>
RAPT_DEV int switch_p2g(const FieldP &f, const double (&prow)[7], double mass, double q,
                        double (&grow)[5], double &mu, double &vout)
{
    double gm = sqrt(mass * mass + dot3(prow[4], prow[5], prow[6], prow[4], prow[5], prow[6]) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    double v[3] = {prow[4] / gm, prow[5] / gm, prow[6] / gm};
    double r[3] = {prow[1], prow[2], prow[3]}, g[3], old[3], gc[3];
    const double t = prow[0];
    gyrovector<F>(f, t, r, v, mass, q, g);
    old[0] = r[0] - g[0]; old[1] = r[1] - g[1]; old[2] = r[2] - g[2];
    bool ok = false;
    for (int it = 1; it <= 20; it++) {
        gyrovector<F>(f, t, old, v, mass, q, g);
        gc[0] = r[0] - g[0]; gc[1] = r[1] - g[1]; gc[2] = r[2] - g[2];
        double dx = gc[0] - old[0], dy = gc[1] - old[1], dz = gc[2] - old[2];
        if (sqrt(dot3(dx, dy, dz, dx, dy, dz)) / sqrt(dot3(gc[0], gc[1], gc[2], gc[0], gc[1], gc[2])) < 1e-3) { ok = true; break; }
        old[0] = gc[0]; old[1] = gc[1]; old[2] = gc[2];
    }
    if (!ok) return RAPT_ST_GCITER;
    double bx, by, bz; F::B(f, t, gc[0], gc[1], gc[2], bx, by, bz);
    double vp = dot3(v[0], v[1], v[2], bx, by, bz) / sqrt(dot3(bx, by, bz, bx, by, bz));
    double spd = sqrt(dot3(v[0], v[1], v[2], v[0], v[1], v[2]));
    double sc = spd / RAPT_C_LIGHT;
    double gamma = 1 / sqrt(1 - sc * sc);
    double pp;
    gc_construct<F>(f, t, gc[0], gc[1], gc[2], spd, 0, false, mass * gamma * vp, mass, pp, mu);
    grow[0] = t; grow[1] = gc[0]; grow[2] = gc[1]; grow[3] = gc[2]; grow[4] = pp;
    vout = spd;
    return 0;
}

// Particle.init(GuidingCenter) :149-164 with utils.GCtoFP :422-433, getperp :360-374, gyrophase 0;
// the field is evaluated at t_eval = the new Particle's tcur (quirk Q12)
template <class F>
RAPT_DEV void switch_g2p(const FieldP &f, const double (&grow)[5], double mu, double mass, double q, double t_eval,
                         double (&prow)[7])
{
    double B = F::magB(f, grow[0], grow[1], grow[2], grow[3]);
    double pm = grow[4] / mass / RAPT_C_LIGHT;
    double gammasq = 1 + 2 * mu * B / (mass * RAPT_C_LIGHT * RAPT_C_LIGHT) + pm * pm;
    double v;
    if (sqrt(gammasq) - 1 < 1e-6) { double pv = grow[4] / mass; v = sqrt(2 * mu * B / mass + pv * pv); }
    else v = RAPT_C_LIGHT * sqrt(1 - 1 / gammasq);
    double vpar = grow[4] / mass / sqrt(gammasq);
    // GCtoFP
    double bx, by, bz; F::B(f, t_eval, grow[1], grow[2], grow[3], bx, by, bz);
    double sB = sqrt(dot3(bx, by, bz, bx, by, bz));
    double b[3] = {bx / sB, by / sB, bz / sB};
    double pa = acos(vpar / v);
    double rc = cycrad2(sB, vpar, v, mass, q);
    double u[3];
    if (bx == 0) { u[0] = 1; u[1] = 0; u[2] = 0; }
    else if (by == 0) { u[0] = 0; u[1] = 1; u[2] = 0; }
    else if (bz == 0) { u[0] = 0; u[1] = 0; u[2] = 1; }
    else { double cc = -1.0 * (bx + by) / bz, nrm = sqrt(2 + cc * cc); u[0] = 1 / nrm; u[1] = 1 / nrm; u[2] = cc / nrm; }
    double un = sqrt(dot3(u[0], u[1], u[2], u[0], u[1], u[2]));
    u[0] /= un; u[1] /= un; u[2] /= un;
    double w[3] = {b[1] * u[2] - b[2] * u[1], b[2] * u[0] - b[0] * u[2], b[0] * u[1] - b[1] * u[0]};
    double s = sgn(q);
    double cg = 1.0, sg = 0.0, cpa = cos(pa), spa = sin(pa);
    double vel[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        prow[1 + i] = grow[1 + i] + rc * (cg * u[i] + sg * w[i]);
        vel[i] = v * ((cpa * b[i] + s * spa * sg * u[i]) - s * spa * cg * w[i]);
    }
    // Particle.__init__ :106-108
    double gamma = 1 / sqrt(1 - dot3(vel[0], vel[1], vel[2], vel[0], vel[1], vel[2]) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    prow[0] = grow[0];
#pragma unroll
    for (int i = 0; i < 3; i++) prow[4 + i] = mass * gamma * vel[i];
}

// ------------------------------------------------------------------------------------------------

template <class F>
__global__ void k_field_ops(const OpsArgs a)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double t = a.tpos[4 * i], x = a.tpos[4 * i + 1], y = a.tpos[4 * i + 2], z = a.tpos[4 * i + 3];
    double u, v, w;
    if (a.B) { F::B(a.f, t, x, y, z, u, v, w); a.B[3 * i] = u; a.B[3 * i + 1] = v; a.B[3 * i + 2] = w; }
    if (a.E) { F::E(a.f, t, x, y, z, u, v, w); a.E[3 * i] = u; a.E[3 * i + 1] = v; a.E[3 * i + 2] = w; }
    if (a.unitb) { F::unitb(a.f, t, x, y, z, u, v, w); a.unitb[3 * i] = u; a.unitb[3 * i + 1] = v; a.unitb[3 * i + 2] = w; }
    if (a.magB) a.magB[i] = F::magB(a.f, t, x, y, z);
    if (a.gradB) { F::gradB(a.f, t, x, y, z, u, v, w); a.gradB[3 * i] = u; a.gradB[3 * i + 1] = v; a.gradB[3 * i + 2] = w; }
    if (a.jac) { double J[9]; F::jacobianB(a.f, t, x, y, z, J); for (int k = 0; k < 9; k++) a.jac[9 * i + k] = J[k]; }
    if (a.curlb) { F::curlb(a.f, t, x, y, z, u, v, w); a.curlb[3 * i] = u; a.curlb[3 * i + 1] = v; a.curlb[3 * i + 2] = w; }
    if (a.curv) a.curv[i] = F::curvature(a.f, t, x, y, z);
    if (a.dBdt) a.dBdt[i] = a.f.is_static ? 0.0 : F::dBdt(a.f, t, x, y, z);
    if (a.dbdt) {
        if (a.f.is_static) { u = v = w = 0; } else F::dbdt(a.f, t, x, y, z, u, v, w);
        a.dbdt[3 * i] = u; a.dbdt[3 * i + 1] = v; a.dbdt[3 * i + 2] = w;
    }
    if (a.lscale) a.lscale[i] = F::lengthscale(a.f, t, x, y, z);
    if (a.tscale) a.tscale[i] = a.f.is_static ? nan("") : F::timescale(a.f, t, x, y, z);
}


template <class F>
__global__ void k_misc(const MiscArgs a)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (a.op == 0) {          // a0..a6 = t0,x,y,z,v,pa,mass -> o0 = ppar, o1 = mu
        double pp, mu;
        gc_construct<F>(a.f, a.a0[i], a.a1[i], a.a2[i], a.a3[i], a.a4[i], a.a5[i], true, 0.0, a.a6[i], pp, mu);
        a.o0[i] = pp; a.o1[i] = mu;
    } else if (a.op == 1) {   // a0 = prow7, a1 = mass, a2 = charge -> o0 = grow5, o1 = mu, o2 = v, io = status
        double pr[7], gr[5] = {0, 0, 0, 0, 0}, mu = 0, v = 0;
        for (int k = 0; k < 7; k++) pr[k] = a.a0[7 * i + k];
        int rc = switch_p2g<F>(a.f, pr, a.a1[i], a.a2[i], gr, mu, v);
        for (int k = 0; k < 5; k++) a.o0[5 * i + k] = gr[k];
        a.o1[i] = mu; a.o2[i] = v; a.io[i] = rc;
    } else if (a.op == 2) {   // a0 = grow5, a1 = mu, a2 = mass, a3 = charge -> o0 = prow7
        double gr[5], pr[7];
        for (int k = 0; k < 5; k++) gr[k] = a.a0[5 * i + k];
        switch_g2p<F>(a.f, gr, a.a1[i], a.a2[i], a.a3[i], a.t_eval, pr);
        for (int k = 0; k < 7; k++) a.o0[7 * i + k] = pr[k];
    } else if (a.op == 3) {   // a0 = rows (stride), a1 = mu, a2 = mass, a3 = charge -> io
        const double *r = a.a0 + a.stride * i;
        if (a.mode == 0) {
            double y[6] = {r[1], r[2], r[3], r[4], r[5], r[6]};
            a.io[i] = particle_isadiabatic<F>(a.f, a.p, r[0], y, a.a2[i], a.a3[i]) ? 1 : 0;
        } else {
            double y[4] = {r[1], r[2], r[3], r[4]};
            a.io[i] = gc_isadiabatic<F>(a.f, a.p, r[0], y, a.a1[i], a.a2[i], a.a3[i]) ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bounce-period set-up.  One thread per guiding centre: (Bm, v) from (mu, p_par)
// (GuidingCenter.py:595-605), ds = 1/(curvature * fieldlineresolution) (fieldline.py:31-35), then the
// RKF45 trace d(s,x,y,z)/ds = +-(1, b) in chunks of ds, forward with tol 1e-3, backward with 1e-4
// (fieldline.py:66,88; quirk Q8), until |B| > Bm after a chunk.  rkf.py's control flow is kept verbatim
// (accepted 4th-order solution, factor 0.84 (tol/r)^(1/4) in [0.1, 4], hmin 1e-6 -> chunk abandoned).
// ------------------------------------------------------------------------------------------------

template <class F>
RAPT_DEV void fl_deriv(const FieldP &f, double time, const double (&Y)[4], double sign, double (&o)[4])
{
    double ux, uy, uz; F::unitb(f, time, Y[1], Y[2], Y[3], ux, uy, uz);
    o[0] = sign * 1.0; o[1] = sign * ux; o[2] = sign * uy; o[3] = sign * uz;
}

// one rkf() call (rkf.py:13-143) over [0, b]; appends accepted points to out (stride 4).
// Returns points appended, or -1 if h < hmin (RuntimeError in the reference).
template <class F>
RAPT_DEV long long rkf_chunk(const FieldP &f, double time, double sign, const double (&x0)[4], double b, double tol,
                             double hmax, double hmin, double *out, long long cap, double (&xlast)[4])
{
    const double b21 = 2.500000000000000e-01, b31 = 9.375000000000000e-02, b32 = 2.812500000000000e-01,
                 b41 = 8.793809740555303e-01, b42 = -3.277196176604461e+00, b43 = 3.320892125625853e+00,
                 b51 = 2.032407407407407e+00, b52 = -8.000000000000000e+00, b53 = 7.173489278752436e+00,
                 b54 = -2.058966861598441e-01, b61 = -2.962962962962963e-01, b62 = 2.000000000000000e+00,
                 b63 = -1.381676413255361e+00, b64 = 4.529727095516569e-01, b65 = -2.750000000000000e-01;
    const double r1 = 2.777777777777778e-03, r3 = -2.994152046783626e-02, r4 = -2.919989367357789e-02,
                 r5 = 2.000000000000000e-02, r6 = 3.636363636363636e-02;
    const double c1 = 1.157407407407407e-01, c3 = 5.489278752436647e-01, c4 = 5.353313840155945e-01,
                 c5 = -2.000000000000000e-01;
    double t = 0, h = hmax, x[4], k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], yy[4], d[4];
    long long n = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = x0[i];
    while (t < b) {
        if (t + h > b) h = b - t;
        fl_deriv<F>(f, time, x, sign, d);
#pragma unroll
        for (int i = 0; i < 4; i++) { k1[i] = h * d[i]; yy[i] = x[i] + b21 * k1[i]; }
        fl_deriv<F>(f, time, yy, sign, d);
#pragma unroll
        for (int i = 0; i < 4; i++) { k2[i] = h * d[i]; yy[i] = x[i] + b31 * k1[i] + b32 * k2[i]; }
        fl_deriv<F>(f, time, yy, sign, d);
#pragma unroll
        for (int i = 0; i < 4; i++) { k3[i] = h * d[i]; yy[i] = x[i] + b41 * k1[i] + b42 * k2[i] + b43 * k3[i]; }
        fl_deriv<F>(f, time, yy, sign, d);
#pragma unroll
        for (int i = 0; i < 4; i++) { k4[i] = h * d[i]; yy[i] = x[i] + b51 * k1[i] + b52 * k2[i] + b53 * k3[i] + b54 * k4[i]; }
        fl_deriv<F>(f, time, yy, sign, d);
#pragma unroll
        for (int i = 0; i < 4; i++) { k5[i] = h * d[i]; yy[i] = x[i] + b61 * k1[i] + b62 * k2[i] + b63 * k3[i] + b64 * k4[i] + b65 * k5[i]; }
        fl_deriv<F>(f, time, yy, sign, d);
        double r = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            k6[i] = h * d[i];
            double ri = fabs(r1 * k1[i] + r3 * k3[i] + r4 * k4[i] + r5 * k5[i] + r6 * k6[i]) / h;
            r = fmax(r, ri);
        }
        if (r <= tol) {
            t = t + h;
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = x[i] + c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i];
            if (n < cap) {
#pragma unroll
                for (int i = 0; i < 4; i++) out[4 * n + i] = x[i];
            }
            n++;
        }
        h = h * fmin(fmax(0.84 * pow(tol / r, 0.25), 0.1), 4.0);      // r == 0 -> inf -> 4 (quirk Q15)
        if (h > hmax) h = hmax;
        else if (h < hmin) return -1;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) xlast[i] = x[i];
    return n;
}


// Fieldline(tpos, field, Bmax=Bm).trace() + getB() (fieldline.py:13-105, 123-127) for one thread: the curve
// (s, x, y, z, |B|) goes to cv (stride 5, room for cap points); bw is scratch for the backward half (stride 4,
// cap points).  Returns the number of points, or cap + 1 when the line does not fit.
template <class F>
RAPT_DEV long long fieldline_trace(const FieldP &f, double t, double x, double y, double z, double Bm, double flres,
                                   double *cv, double *bw, long long cap, double &ds_out)
{
    const double ds = 1 / F::curvature(f, t, x, y, z) / flres;          // fieldline.py:31-35
    ds_out = ds;
    double *fw = cv;                                   // forward half staged in the output buffer (stride 4)
    long long nf = 1, nb = 1;
    fw[0] = 0; fw[1] = x; fw[2] = y; fw[3] = z;
    bw[0] = 0; bw[1] = x; bw[2] = y; bw[3] = z;
    for (int dir = 0; dir < 2; dir++) {
        double *arr = dir ? bw : fw;
        long long np_ = 1;
        const double sign = dir ? -1.0 : 1.0, tol = dir ? 1e-4 : 1e-3;
        double cur[4] = {0, x, y, z}, nxt[4];
        for (;;) {
            if (np_ >= cap) { np_ = cap + 1; break; }
            long long m = rkf_chunk<F>(f, t, sign, cur, ds, tol, ds, 1e-6, arr + 4 * np_, cap - np_, nxt);
            if (m < 0) break;
            if (np_ + m > cap) { np_ = cap + 1; break; }
            np_ += m;
#pragma unroll
            for (int k = 0; k < 4; k++) cur[k] = nxt[k];
            if (!(F::magB(f, t, cur[1], cur[2], cur[3]) <= Bm)) break;      // |B| > Bm, or NaN outside a grid
        }
        if (dir) nb = np_; else nf = np_;
    }
    if (nf > cap || nb > cap || (nb - 1) + nf > cap) return cap + 1;
    const long long n = (nb - 1) + nf;
    // assemble curve = reversed(backward[1:]) + forward, in place: move the forward half up first
    for (long long k = nf - 1; k >= 0; k--) {
        double s = fw[4 * k], px = fw[4 * k + 1], py = fw[4 * k + 2], pz = fw[4 * k + 3];
        double *o = cv + 5 * ((nb - 1) + k);
        o[0] = s; o[1] = px; o[2] = py; o[3] = pz;
    }
    for (long long k = 0; k < nb - 1; k++) {
        const double *q = bw + 4 * (nb - 1 - k);
        double *o = cv + 5 * k;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
    }
    for (long long k = 0; k < n; k++) {
        double *o = cv + 5 * k;
        o[4] = F::magB(f, t, o[1], o[2], o[3]);        // Fieldline.getB, fieldline.py:123-127
    }
    return n;
}

template <class F>
__global__ void __launch_bounds__(128) k_bounce_setup(const BounceArgs a)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double t = a.t[i], x = a.x[i], y = a.y[i], z = a.z[i];
    double Bm, v = 0;
    if (a.mu) {
        // GuidingCenter.bounceperiod :595-605
        const double ppar = a.ppar[i], mu = a.mu[i], mass = a.mass[i];
        const double Bmag = F::magB(a.f, t, x, y, z);
        const double pmc = ppar / (mass * RAPT_C_LIGHT);
        const double gamma = sqrt(1 + 2 * mu * Bmag / (mass * RAPT_C_LIGHT * RAPT_C_LIGHT) + pmc * pmc);
        if (gamma - 1 < 1e-6) {
            double p = sqrt(2 * mass * mu * Bmag + ppar * ppar);
            v = p / mass; Bm = (p * p) / (2 * mass * mu);
        } else {
            double p = mass * RAPT_C_LIGHT * sqrt((gamma + 1) * (gamma - 1));
            Bm = p * p / ((p - ppar) * (p + ppar)) * Bmag;
            v = p / mass / gamma;
        }
    } else Bm = a.Bm[i];                                     // Fieldline(tpos, field, Bmax=Bm)
    const long long cap = a.max_pts;
    double *cv = a.curve + (size_t)i * cap * 5;
    double *bw = a.scratch + (size_t)i * cap * 4;
    double ds;
    const long long n = fieldline_trace<F>(a.f, t, x, y, z, Bm, a.flres, cv, bw, cap, ds);
    a.Bm[i] = Bm; a.v[i] = v; a.ds[i] = ds;
    a.npts[i] = (int)n;
    if (n > cap) { if (a.period) a.period[i] = nan(""); return; }
    // flutils.bounceperiod (flutils.py:252): tau_b = (2/v) S_b.  quadrature 1: brentq + QUADPACK as the reference;
    // 0: mirror points and integral in closed form on the same spline (rapt_quad.cuh)
    if (a.period) a.period[i] = (2 / v) * halfbouncepath_curve(cv, bw, n, Bm, a.quadrature);
}


// ------------------------------------------------------------------------------------------------
// Adaptive epochs.  k_adaptive_switch runs between the per-mode advance kernels:
//   first == 1 : Adaptive.__init__ (Adaptive.py:96-104) -- build the Particle, test isadiabatic(),
//                start as GuidingCenter if adiabatic;
//   first == 0 : consume the status the advance kernels left (Adiabatic / NonAdiabatic raised after a
//                completed row, Particle.py:308-309, GuidingCenter.py:457-458), apply the
//                Particle->GuidingCenter or GuidingCenter->Particle transform (Adaptive.py:208-221),
//                open the new segment (its first row is always stored) and update Adaptive.advance's
//                loop variable t = current.tcur (Adaptive.py:222).
// Then every tracer that still has t < delta is appended to the work list of its mode: warp ballot ->
// per-warp counts -> block-level scan in shared memory -> one atomicAdd per block and mode.  The next
// epoch's kernels therefore see tracers regrouped by mode, finished ones dropped.
// ------------------------------------------------------------------------------------------------
// the per-tracer part: returns the work list the tracer joins (-1 idle/finished, 0 particle list, 1 GC list)
template <class F>
RAPT_DEV int adaptive_switch_one(const AdaptArgs &a, const long long i)
{
    int want = -1;
    {
        const double mass = a.mass[i], q = a.charge[i];
        int mode, st, nseg;
        bool newseg = false;
        double prow[7], grow[5], mu = 0, v = 0;
        if (a.first) {
            // Particle.__init__ :106-108
            const double vx = a.vx0[i], vy = a.vy0[i], vz = a.vz0[i];
            const double gamma = 1 / sqrt(1 - dot3(vx, vy, vz, vx, vy, vz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
            prow[0] = a.t0[i]; prow[1] = a.x0[i]; prow[2] = a.y0[i]; prow[3] = a.z0[i];
            prow[4] = mass * gamma * vx; prow[5] = mass * gamma * vy; prow[6] = mass * gamma * vz;
            const double yy[6] = {prow[1], prow[2], prow[3], prow[4], prow[5], prow[6]};
            st = RAPT_ST_OK; nseg = 1; newseg = true;
            a.nstored[i] = 0; a.tvar[i] = 0; a.tcur[i] = prow[0];
            if (particle_isadiabatic<F>(a.f, a.p, prow[0], yy, mass, q)) {
                int rc = switch_p2g<F>(a.f, prow, mass, q, grow, mu, v);
                mode = 1;
                if (rc) st = rc;
            } else mode = 0;
        } else {
            mode = a.mode[i]; st = a.status[i]; nseg = a.nseg[i];
            if (st == RAPT_ST_ADIABATIC && mode == 0) {          // Adaptive.py:215-221
                prow[0] = a.pt[i]; prow[1] = a.px[i]; prow[2] = a.py[i]; prow[3] = a.pz[i];
                prow[4] = a.ppx[i]; prow[5] = a.ppy[i]; prow[6] = a.ppz[i];
                int rc = switch_p2g<F>(a.f, prow, mass, q, grow, mu, v);
                mode = 1; nseg++; newseg = true;
                st = rc ? rc : RAPT_ST_OK;
                a.tcur[i] = grow[0];                               // GuidingCenter.__init__: tcur = t0
            } else if (st == RAPT_ST_NONADIABATIC && mode == 1) {  // Adaptive.py:208-214
                grow[0] = a.gt[i]; grow[1] = a.gx[i]; grow[2] = a.gy[i]; grow[3] = a.gz[i]; grow[4] = a.gpp[i];
                switch_g2p<F>(a.f, grow, a.mu[i], mass, q, 0.0 /* Particle().tcur, quirk Q12 */, prow);
                mode = 0; nseg++; newseg = true;
                st = RAPT_ST_OK;
                a.tcur[i] = prow[0];                               // Particle.__init__: tcur = t0
            }
        }
        const bool sliced = (!a.first) && st == RAPT_ST_SLICE;          // interrupted at a slice boundary: resume as is
        if (sliced) st = RAPT_ST_OK;
        if (newseg && st == RAPT_ST_OK) {
            const int tag = 2 * (nseg - 1) + mode;
            a.segtag[i] = tag;
            int nst = a.nstored[i];
            if (mode == 0) {
                a.pt[i] = prow[0]; a.px[i] = prow[1]; a.py[i] = prow[2]; a.pz[i] = prow[3];
                a.ppx[i] = prow[4]; a.ppy[i] = prow[5]; a.ppz[i] = prow[6];
                if (a.rows && nst < a.max_rows) {
                    double *r = a.rows + ((size_t)i * a.max_rows + nst) * 8;
                    for (int k = 0; k < 7; k++) r[k] = prow[k];
                    r[7] = (double)tag;
                    a.nstored[i] = nst + 1;
                }
            } else {
                a.gt[i] = grow[0]; a.gx[i] = grow[1]; a.gy[i] = grow[2]; a.gz[i] = grow[3]; a.gpp[i] = grow[4];
                a.mu[i] = mu; a.v[i] = v;
                if (a.rows && nst < a.max_rows) {
                    double *r = a.rows + ((size_t)i * a.max_rows + nst) * 8;
                    for (int k = 0; k < 5; k++) r[k] = grow[k];
                    r[5] = mu; r[6] = 0; r[7] = (double)tag;
                    a.nstored[i] = nst + 1;
                }
            }
        }
        a.mode[i] = mode; a.nseg[i] = nseg; a.status[i] = st;
        if (sliced) want = mode;
        else {
            // Adaptive.py:205,222 : t = current.tcur ; loop while t < delta (absolute tcur vs duration, quirk Q14)
            double tv = a.first ? 0.0 : a.tcur[i];
            a.tvar[i] = tv;
            a.rem[i] = a.delta - tv;
            if (st == RAPT_ST_OK && tv < a.delta) {
                // next reference call: current.advance(delta - t) from the tracer's last row (Adaptive.py:207):
                // tstop = t0 + (delta - t) (Particle.py:304, GuidingCenter.py:452); dt is chosen when it starts
                want = mode;
                const double tstart = (mode == 0) ? a.pt[i] : a.gt[i];
                a.seg_tstop[i] = tstart + (a.delta - tv);
                a.seg_x[i] = tstart;
                a.seg_dt[i] = 0.0;
                if (newseg) a.seg_row[i] = 0;
            }
        }
    }
    return want;
}

template <class F>
__global__ void __launch_bounds__(256) k_adaptive_switch(const AdaptArgs a)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int want = (i < a.n) ? adaptive_switch_one<F>(a, i) : -1;
    // ---- regroup by mode: ballot + shared-memory scan + one atomic per block and mode
    __shared__ int wcount[2][8];
    __shared__ int wbase[2][8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mP = __ballot_sync(0xffffffffu, want == 0), mG = __ballot_sync(0xffffffffu, want == 1);
    if (lane == 0) { wcount[0][warp] = __popc(mP); wcount[1][warp] = __popc(mG); }
    __syncthreads();
    if (threadIdx.x < 2) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { wbase[threadIdx.x][w] = tot; tot += wcount[threadIdx.x][w]; }
        int base = tot ? atomicAdd(a.counts + threadIdx.x, tot) : 0;
        for (int w = 0; w < 8; w++) wbase[threadIdx.x][w] += base;
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1;
    if (want == 0) a.listP[wbase[0][warp] + __popc(mP & below)] = (int)i;
    if (want == 1) a.listG[wbase[1][warp] + __popc(mG & below)] = (int)i;
}

}  // namespace RAPT_NS
