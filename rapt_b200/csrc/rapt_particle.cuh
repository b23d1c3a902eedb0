// rapt_particle.cuh -- Particle.advance (Particle.py:230-309) for an ensemble: one thread per
// particle, fp64 state and all ten DOP853 stage vectors in registers, persistent lanes that pull the
// next particle from a global queue when theirs finishes.
//
// The integrator is scipy.integrate.ode "dop853" as the reference drives it (Particle.py:300-305):
// a FRESH solver call per output row (HINIT, facold = 1e-4, nsteps = 500), beta = 0.1,
// rtol/atol = params["solvertolerances"].  See SURVEY.md §3.5 for the control flow.
//
// Loop shape (bounds warp divergence): every lane runs one flat loop whose body is exactly one RK
// step attempt; "fetch next particle" and "start next output row (HINIT)" are short predicated
// prologues of the same iteration, so lanes that sit in different rows / different particles still
// execute the 11+1 stage evaluations together.
#pragma once
#include "rapt_fields.cuh"
#include "dop_const.cuh"

namespace RAPT_NS {

using rapt::ParamsP;
using rapt::AdvArgs;

#define ST_OK RAPT_ST_OK
#define ST_ADIABATIC RAPT_ST_ADIABATIC
#define ST_NONADIABATIC RAPT_ST_NONADIABATIC
#define ST_NMAX RAPT_ST_NMAX
#define ST_HSMALL RAPT_ST_HSMALL

// ---------------------------------------------------------------------------------------------
// Newton-Lorentz right-hand side, Particle.py:284-298.  gm = gamma*m is frozen at its entry value
// when the field is static (quirk Q3), recomputed from p otherwise.
// ---------------------------------------------------------------------------------------------
template <class F>
RAPT_DEV void particle_rhs(const FieldP &f, double q, double mass, double &gm, double &igm, int equatorial,
                           double t, const double (&Y)[6], double (&out)[6])
{
    double bx, by, bz;
    if (!f.is_static) {
        gm = sqrt(mass * mass + dot3(Y[3], Y[4], Y[5], Y[3], Y[4], Y[5]) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
#if !RAPT_STRICT
        igm = fast_rcp(gm);
#endif
    }
    F::B(f, t, Y[0], Y[1], Y[2], bx, by, bz);
    double cx = Y[4] * bz - Y[5] * by, cy = Y[5] * bx - Y[3] * bz, cz = Y[3] * by - Y[4] * bx;
#if RAPT_STRICT
    out[0] = Y[3] / gm; out[1] = Y[4] / gm; out[2] = Y[5] / gm;
    if (F::HAS_E) {
        double ex, ey, ez; F::E(f, t, Y[0], Y[1], Y[2], ex, ey, ez);
        out[3] = q * (ex + cx / gm); out[4] = q * (ey + cy / gm); out[5] = q * (ez + cz / gm);
    } else {
        out[3] = q * (0.0 + cx / gm); out[4] = q * (0.0 + cy / gm); out[5] = q * (0.0 + cz / gm);
    }
#else
    out[0] = Y[3] * igm; out[1] = Y[4] * igm; out[2] = Y[5] * igm;
    double qg = q * igm;
    if (F::HAS_E) {
        double ex, ey, ez; F::E(f, t, Y[0], Y[1], Y[2], ex, ey, ez);
        out[3] = fma(q, ex, qg * cx); out[4] = fma(q, ey, qg * cy); out[5] = fma(q, ez, qg * cz);
    } else {
        out[3] = qg * cx; out[4] = qg * cy; out[5] = qg * cz;
    }
#endif
    if (equatorial) { out[2] = 0; out[5] = 0; }
}

// utils.cyclotron_radius (utils.py:139-145) / lengthscale (fields.py:261) < epss
// [and cyclotron_period (utils.py:63-66) / timescale (fields.py:277-280) < epst when not static]:
// Particle.isadiabatic, Particle.py:380-384
template <class F>
RAPT_DEV bool particle_isadiabatic(const FieldP &f, const ParamsP &p, double t, const double (&y)[6], double mass, double q)
{
#if !RAPT_STRICT
    // fast flavour: the same inequality without a square root or a division.  gamma m v_perp = p_perp,
    // p_perp^2 B^2 = p^2 B^2 - (p.B)^2, L = |B| / max|J| and max|J| = maxdiff / (2 d), so
    //   rho / L < epss  <=>  p_perp maxdiff < epss |q| 2d B^2  <=>  (p^2 B^2 - (p.B)^2) maxdiff^2 < (epss |q| 2d)^2 (B^2)^3.
    // The reference's quotient form costs ~20 fp64 divisions / square roots per row, executed by the few lanes that
    // finish a row in the same iteration (profiles/r2_adaptive.md).
    {
        double bx, by, bz; F::B(f, t, y[0], y[1], y[2], bx, by, bz);
        const double B2 = dot3(bx, by, bz, bx, by, bz), pB = dot3(y[3], y[4], y[5], bx, by, bz);
        const double pp2B2 = fma(dot3(y[3], y[4], y[5], y[3], y[4], y[5]), B2, -pB * pB);
        const double md = F::max_central_difference(f, t, y[0], y[1], y[2]);
        const double lim = p.epss * fabs(q) * (2 * f.gradstep);
        const bool sp = pp2B2 * (md * md) < (lim * lim) * (B2 * B2 * B2);
        if (f.is_static || !sp) return sp;
    }
#endif
    double gm = sqrt(mass * mass + dot3(y[3], y[4], y[5], y[3], y[4], y[5]) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    double vx = y[3] / gm, vy = y[4] / gm, vz = y[5] / gm;
    double vsq = dot3(vx, vy, vz, vx, vy, vz);
    double gamma = 1.0 / sqrt(1 - vsq / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    double bx, by, bz; F::B(f, t, y[0], y[1], y[2], bx, by, bz);
    double Bmag = sqrt(dot3(bx, by, bz, bx, by, bz));
    double vpar = dot3(vx, vy, vz, bx, by, bz) / Bmag;
    double vperp = sqrt(vsq - vpar * vpar);
    double rho = gamma * mass * vperp / (fabs(q) * Bmag);
    bool sp = rho / F::lengthscale(f, t, y[0], y[1], y[2]) < p.epss;
    if (f.is_static || !sp) return sp;
    double per = 2 * RAPT_PI * gamma * mass / Bmag / fabs(q);
    return per / F::timescale(f, t, y[0], y[1], y[2]) < p.epst;
}

// ---------------------------------------------------------------------------------------------
// The kernel.  Flat per-lane loop whose body is exactly ONE step attempt, so lanes that sit in
// different rows / particles still execute the stage arithmetic and field evaluations together.
// Inside the body the 13 right-hand-side evaluations (HINIT's Euler probe for lanes that start a row,
// stages 2..12, the FSAL evaluation for lanes that accepted) go through ONE copy of the field code in
// a warp-uniform, non-unrolled stage loop: the hot loop is ~25 KB of SASS and stays inside the 32 KB
// L1.5 instruction cache.  (v1 inlined 13 copies, 53 KB, and stalled on instruction fetch; v2 ran one
// evaluation per iteration with a per-lane stage and serialised the stage arithmetic across lanes --
// profiles/r1_particle_history.md.)  All k-vectors are indexed statically inside their case so they
// live in registers.
// ---------------------------------------------------------------------------------------------
#if RAPT_STRICT
#define RAPT_POW(x, e) pow((x), (e))
#else
// x > 0; relative error ~1e-15, no special-case slow paths
#define RAPT_POW(x, e) exp((e) * log(x))
#endif

#define D8_PREP(CC, EXPR)                                                           \
    {                                                                               \
        _Pragma("unroll") for (int i = 0; i < 6; i++) yin[i] = y[i] + h * (EXPR);   \
        tin = x + (CC) * h;                                                         \
    }
#define D8_SAVE(K) { _Pragma("unroll") for (int i = 0; i < 6; i++) K[i] = kout[i]; }

template <class F>
__global__ void __launch_bounds__(128, 2) k_particle_dop853(const AdvArgs a)
{
    if (F::CAN_FAIL) grid_cache_reset();     // gridded field: per-thread cell cache (rapt_fields.cuh)
    const double rtol = a.p.rtol, atol = a.p.atol;
    const int eqf = a.p.enforce_equatorial;
    const double beta = 0.1, safe = 0.9, fac1 = 0.3, fac2 = 6.0, uround = 2.3e-16;
    const double expo1 = 1.0 / 8.0 - beta * 0.2, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const double pf0 = pow(1e-4, beta);          // facold^beta at the first step of every row

    double y[6], k1[6], k2[6], k3[6], k4[6], k5[6], k6[6], k7[6], k8[6], k9[6], k10[6], yin[6], kout[6];
    double x = 0, h = 0, xend = 0, label = 0, tstop = 0, tlim = 0, dt = 0, facold = 1e-4, hmax = 0, tin = 0, dnf = 0, hnew = 0;
    double mass = 0, q = 0, gm = 1, igm = 1;
    int pid = -1;
    int nstep = 0, naccpt = 0, nrejct = 0, ncalls = 0, nstep_row = 0, naccpt_row = 0;
    int rowidx = 0, nst = 0, st = ST_OK;
    bool last = false, reject = false, need_row = false, have = false;
    double *myrows = nullptr;
#if !RAPT_STRICT
    double isk[6];
#endif

    for (;;) {
        // ---- (A) particle finished?  write it back and fetch the next one
        if (have && need_row && !(st == ST_OK && x < tlim)) {
            if (a.seg_tstop) {                               // sliced adaptive epoch: keep the call's state
                a.seg_x[pid] = x; a.seg_dt[pid] = dt; a.seg_row[pid] = rowidx;
                if (st == ST_OK && x < tstop) st = RAPT_ST_SLICE;
            }
            a.t[pid] = label; a.s1[pid] = y[0]; a.s2[pid] = y[1]; a.s3[pid] = y[2];
            a.s4[pid] = y[3]; a.s5[pid] = y[4]; a.s6[pid] = y[5];
            int *c = a.counters + 4 * (long long)pid;
            int nf = 2 * ncalls + 11 * nstep + naccpt;       // as scipy counts: SURVEY.md §3.1
            if (a.append) { c[0] += nf; c[1] += nstep; c[2] += naccpt; c[3] += nrejct; }
            else { c[0] = nf; c[1] = nstep; c[2] = naccpt; c[3] = nrejct; }
            a.status[pid] = st;
            a.tcur[pid] = x + dt;                            // Particle.py:306
            if (a.nrows) a.nrows[pid] = rowidx + 1;
            a.nstored[pid] = nst;
            have = false;
        }
        if (!have) {
            int w = atomicAdd(a.queue, 1);
            if (w >= a.nwork) break;
            pid = a.order ? a.order[w] : w;
            x = a.t[pid];
            y[0] = a.s1[pid]; y[1] = a.s2[pid]; y[2] = a.s3[pid];
            y[3] = a.s4[pid]; y[4] = a.s5[pid]; y[5] = a.s6[pid];
            mass = a.mass[pid]; q = a.charge[pid];
            double delta = a.delta_arr ? a.delta_arr[pid] : a.delta;
            tstop = x + delta;                               // Particle.py:304  (t0 + delta)
            tlim = tstop;
            label = x;
            double dt_fixed = 0;
            int row0 = 0;
            if (a.seg_tstop) {                               // resume a sliced call
                tstop = a.seg_tstop[pid]; tlim = fmin(tstop, a.slice_end);
                x = a.seg_x[pid]; dt_fixed = a.seg_dt[pid]; row0 = a.seg_row[pid];
            }
            // Particle.py:274-275, 282: gm, vel, dt = cyclotron_period / cyclotronresolution
            gm = sqrt(mass * mass + dot3(y[3], y[4], y[5], y[3], y[4], y[5]) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
            igm = 1.0 / gm;
            {
                double vx = y[3] / gm, vy = y[4] / gm, vz = y[5] / gm;
                double gamma = 1.0 / sqrt(1 - dot3(vx, vy, vz, vx, vy, vz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
                double Bm = F::magB(a.f, x, y[0], y[1], y[2]);
                dt = 2 * RAPT_PI * gamma * mass / Bm / fabs(q) / a.p.cyclotronresolution;
            }
            if (dt_fixed > 0) dt = dt_fixed;
            if (a.dt_out) a.dt_out[pid] = dt;
            nstep = naccpt = nrejct = ncalls = 0; rowidx = row0; st = ST_OK;
            myrows = a.rows ? a.rows + (size_t)pid * (size_t)a.max_rows * 8 : nullptr;
            if (a.append) nst = a.nstored[pid];
            else {
                nst = 0;
                if (myrows && a.store_every > 0 && a.max_rows > 0) {
                    double2 *r = reinterpret_cast<double2 *>(myrows);
                    r[0] = make_double2(x, y[0]); r[1] = make_double2(y[1], y[2]);
                    r[2] = make_double2(y[3], y[4]); r[3] = make_double2(y[5], 0.0);
                    nst = 1;
                }
            }
            have = true; need_row = true;
            if (!(dt > 0.0) || dt > 1e300) { st = RAPT_ST_HSMALL; continue; }   // degenerate output step (B = 0 or inf, bad
                                                                             // resolution): the reference would never return
            if (!(x < tlim)) continue;                       // delta <= 0 (or beyond this slice): nothing to do
            particle_rhs<F>(a.f, q, mass, gm, igm, eqf, x, y, k1);           // k1 = f(x, y)
        }
        // ---- (B) one step attempt; stage 1 = HINIT for lanes that start an output row
        bool accepted = false, skip = false, rowdone = false;
#pragma unroll 1
        for (int s = 1; s <= 13; s++) {
            bool active = !skip;
            switch (s) {
            case 1:
                active = need_row;
                if (active) {
                    // new output row = new solver call: xend, HINIT part 1
                    xend = x + dt; label = xend;             // Particle.py:305
                    hmax = fabs(xend - x);
                    double dny = 0;
                    dnf = 0;
#pragma unroll
                    for (int i = 0; i < 6; i++) {
#if RAPT_STRICT
                        double sk = atol + rtol * fabs(y[i]);
                        dnf += (k1[i] / sk) * (k1[i] / sk);
                        dny += (y[i] / sk) * (y[i] / sk);
#else
                        isk[i] = fast_rcp(atol + rtol * fabs(y[i]));
                        double a_ = k1[i] * isk[i], b_ = y[i] * isk[i];
                        dnf += a_ * a_; dny += b_ * b_;
#endif
                    }
                    h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
                    h = fmin(h, hmax);
#pragma unroll
                    for (int i = 0; i < 6; i++) yin[i] = y[i] + h * k1[i];
                    tin = x + h;
                }
                break;
            case 2:
                if (skip) break;                 // HINIT left the grid
                // step prologue (every lane): failure checks, clip the step to the row end
                if (nstep_row > 500) st = ST_NMAX;
                else if (0.1 * fabs(h) <= fabs(x) * uround) st = ST_HSMALL;
                if (st != ST_OK) {
                    // solver failure: r.integrate() hands back the state it reached and the reference appends it as a row
                    // labelled with the row's end time before `while r.successful()` ends the loop (Particle.py:304-307)
                    rowdone = true; skip = true; active = false;
                } else {
                    if ((x + 1.01 * h - xend) > 0.0) { h = xend - x; last = true; }
                    nstep_row++; nstep++;
#pragma unroll
                    for (int i = 0; i < 6; i++) yin[i] = y[i] + h * T8(A2_1) * k1[i];   // Hairer's association (h*a21)*k1
                    tin = x + T8(C2) * h;
                }
                break;
            case 3: D8_PREP(T8(C3), T8(A3_1) * k1[i] + T8(A3_2) * k2[i]) break;
            case 4: D8_PREP(T8(C4), T8(A4_1) * k1[i] + T8(A4_3) * k3[i]) break;
            case 5: D8_PREP(T8(C5), T8(A5_1) * k1[i] + T8(A5_3) * k3[i] + T8(A5_4) * k4[i]) break;
            case 6: D8_PREP(T8(C6), T8(A6_1) * k1[i] + T8(A6_4) * k4[i] + T8(A6_5) * k5[i]) break;
            case 7: D8_PREP(T8(C7), T8(A7_1) * k1[i] + T8(A7_4) * k4[i] + T8(A7_5) * k5[i] + T8(A7_6) * k6[i]) break;
            case 8: D8_PREP(T8(C8), T8(A8_1) * k1[i] + T8(A8_4) * k4[i] + T8(A8_5) * k5[i] + T8(A8_6) * k6[i] + T8(A8_7) * k7[i]) break;
            case 9: D8_PREP(T8(C9), T8(A9_1) * k1[i] + T8(A9_4) * k4[i] + T8(A9_5) * k5[i] + T8(A9_6) * k6[i] + T8(A9_7) * k7[i] + T8(A9_8) * k8[i]) break;
            case 10: D8_PREP(T8(C10), T8(A10_1) * k1[i] + T8(A10_4) * k4[i] + T8(A10_5) * k5[i] + T8(A10_6) * k6[i] + T8(A10_7) * k7[i] + T8(A10_8) * k8[i] + T8(A10_9) * k9[i]) break;
            case 11: D8_PREP(T8(C11), T8(A11_1) * k1[i] + T8(A11_4) * k4[i] + T8(A11_5) * k5[i] + T8(A11_6) * k6[i] + T8(A11_7) * k7[i] + T8(A11_8) * k8[i] + T8(A11_9) * k9[i] + T8(A11_10) * k10[i]) break;
            case 12:
#pragma unroll
                for (int i = 0; i < 6; i++)
                    yin[i] = y[i] + h * (T8(A12_1) * k1[i] + T8(A12_4) * k4[i] + T8(A12_5) * k5[i] + T8(A12_6) * k6[i] + T8(A12_7) * k7[i] + T8(A12_8) * k8[i] + T8(A12_9) * k9[i] + T8(A12_10) * k10[i] + T8(A12_11) * k2[i]);
                tin = x + h;
                break;
            default:   // 13: FSAL evaluation f(x+h, ynew) for lanes whose step was accepted
                active = accepted;
#pragma unroll
                for (int i = 0; i < 6; i++) yin[i] = k5[i];
                tin = x + h;
                break;
            }

            if (active) particle_rhs<F>(a.f, q, mass, gm, igm, eqf, tin, yin, kout);

            switch (s) {
            case 1:
                if (active) {
                    double der2 = 0;
#pragma unroll
                    for (int i = 0; i < 6; i++) {
#if RAPT_STRICT
                        double sk = atol + rtol * fabs(y[i]);
                        der2 += ((kout[i] - k1[i]) / sk) * ((kout[i] - k1[i]) / sk);
#else
                        double d_ = (kout[i] - k1[i]) * isk[i];
                        der2 += d_ * d_;
#endif
                    }
                    // gridded field: the probe point may lie outside the grid, where the reference's interpolator raises
                    // from inside r.integrate() (no row for this call, the rows so far are kept)
                    if (F::CAN_FAIL && !(der2 == der2)) { st = RAPT_ST_FIELD; skip = true; break; }
                    der2 = sqrt(der2) / h;
                    double der12 = fmax(fabs(der2), sqrt(dnf));
                    // h1 = (0.01/der12)^(1/8) only matters when it is the smallest of the three candidates;
                    // h1 >= hmax  <=>  0.01/der12 >= hmax^8 (monotone), decided without the pow unless close.
                    double h1;
                    if (der12 <= 1e-15) h1 = fmax(1e-6, fabs(h) * 1e-3);
                    else {
                        double tq = 0.01 / der12, hm2 = hmax * hmax, hm4 = hm2 * hm2;
                        if (tq > hm4 * hm4 * 1.000001) h1 = hmax;
                        else h1 = pow(tq, 1.0 / 8.0);
                    }
                    h = fmin(fmin(100 * fabs(h), h1), hmax);
                    facold = 1e-4; last = false; reject = false; nstep_row = 0; naccpt_row = 0;
                    ncalls++;
                    need_row = false;
                }
                break;
            case 2: if (active) D8_SAVE(k2) break;
            case 3: if (active) D8_SAVE(k3) break;
            case 4: if (active) D8_SAVE(k4) break;
            case 5: if (active) D8_SAVE(k5) break;
            case 6: if (active) D8_SAVE(k6) break;
            case 7: if (active) D8_SAVE(k7) break;
            case 8: if (active) D8_SAVE(k8) break;
            case 9: if (active) D8_SAVE(k9) break;
            case 10: if (active) D8_SAVE(k10) break;
            case 11: if (active) D8_SAVE(k2) break;
            case 12:
                if (active) {
                    D8_SAVE(k3)
                    double err = 0, err2 = 0;
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        k4[i] = T8(B1) * k1[i] + T8(B6) * k6[i] + T8(B7) * k7[i] + T8(B8) * k8[i] + T8(B9) * k9[i] + T8(B10) * k10[i] + T8(B11) * k2[i] + T8(B12) * k3[i];
                        k5[i] = y[i] + h * k4[i];
                        double sk = atol + rtol * fmax(fabs(y[i]), fabs(k5[i]));
                        double e3 = k4[i] - T8(BHH1) * k1[i] - T8(BHH2) * k9[i] - T8(BHH3) * k3[i];
                        double e5 = T8(ER1) * k1[i] + T8(ER6) * k6[i] + T8(ER7) * k7[i] + T8(ER8) * k8[i] + T8(ER9) * k9[i] + T8(ER10) * k10[i] + T8(ER11) * k2[i] + T8(ER12) * k3[i];
#if RAPT_STRICT
                        err2 += (e3 / sk) * (e3 / sk);
                        err += (e5 / sk) * (e5 / sk);
#else
                        double is_ = fast_rcp(sk);
                        e3 *= is_; e5 *= is_;
                        err2 += e3 * e3; err += e5 * e5;
#endif
                    }
                    double deno = err + 0.01 * err2;
                    if (deno <= 0.0) deno = 1.0;
#if RAPT_STRICT
                    err = fabs(h) * err * sqrt(1.0 / (6 * deno));
#else
                    err = fabs(h) * err * fast_rsqrt(6 * deno);
#endif
                    if (err <= 1.0) {
                        // accepted.  The controller's new step is only consumed when the row continues.
                        if (!last) {
                            double fac11 = RAPT_POW(err, expo1);
                            double fac = fac11 / ((facold == 1e-4) ? pf0 : RAPT_POW(facold, beta));
                            fac = fmax(facc2, fmin(facc1, fac / safe));
                            hnew = h / fac;
                            if (fabs(hnew) > hmax) hnew = hmax;
                            if (reject) hnew = fmin(fabs(hnew), fabs(h));
                        }
                        facold = fmax(err, 1e-4);
                        naccpt++; naccpt_row++;
                        accepted = true;
                    } else {
                        // rejected: scipy 1.18.1 shrinks by 1/facc1 whatever err is; Hairer uses the controller
                        if (a.p.dop853_reject_rule == 1) h = h / fmin(facc1, RAPT_POW(err, expo1) / safe);
                        else h = h / facc1;
                        reject = true;
                        if (naccpt_row >= 1) nrejct++;
                        last = false;
                        if (F::CAN_FAIL && !(err == err)) { st = RAPT_ST_FIELD; need_row = true; }   // left the grid: keep the last row
                    }
                }
                break;
            default:
                if (active) {
                    D8_SAVE(k1)                              // k1 = f(x+h, ynew)
#pragma unroll
                    for (int i = 0; i < 6; i++) y[i] = k5[i];
                    x = x + h;
                    if (last) rowdone = true;
                    else {
                        h = hnew;
                        reject = false;
                    }
                }
                break;
            }
        }
        if (rowdone) {
            // ---- output row complete (Particle.py:305-309)
            rowidx++;
            if (myrows && a.store_every > 0 && (rowidx % a.store_every) == 0 && nst < a.max_rows) {
                double2 *r = reinterpret_cast<double2 *>(myrows + (size_t)nst * 8);
                double tag = a.segtag ? (double)a.segtag[pid] : (double)nstep;
                r[0] = make_double2(label, y[0]); r[1] = make_double2(y[1], y[2]);
                r[2] = make_double2(y[3], y[4]); r[3] = make_double2(y[5], tag);
                nst++;
            }
            if (a.p.check_adiabaticity) {
                if (particle_isadiabatic<F>(a.f, a.p, label, y, mass, q)) st = ST_ADIABATIC;
            }
            need_row = true;
        }
    }
}
#undef D8_PREP
#undef D8_SAVE

}  // namespace RAPT_NS
