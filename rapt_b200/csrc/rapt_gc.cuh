// rapt_gc.cuh -- GuidingCenter.advance (GuidingCenter.py:397-458) for an ensemble: one thread per
// guiding centre, DOPRI5 (scipy "dopri5" as driven at GuidingCenter.py:450-453: fresh call per
// output row, beta -> 0.04, nsteps 500) with the three selectable equations of motion
// (GuidingCenter.py:329-395) on the finite-difference operators of rapt/fields.py.
//
// Loop shape (same as the particle kernel): a flat per-lane loop whose body is ONE step attempt; the
// right-hand side (7 field evaluations + normalisations) has a single call site inside a warp-uniform,
// non-unrolled stage loop (HINIT's Euler probe for lanes that start an output row, then stages 2..7).
// The first version ran one RHS per iteration with a per-lane stage: the RHS stayed converged (29 of
// 32 lanes) but the per-stage bookkeeping then executed once per distinct stage in the warp at ~5
// lanes and cost as many issue slots as the RHS itself (profiles/r1_gc_v1_ncu_summary.txt).
//
// The reference evaluates curlb (6 x unitb) and gradB (6 x magB) at the SAME six shifted points
// (fields.py:125-131, 191-200); unitb and magB of one point share B and sqrt(B.B), so evaluating each
// point once (7 field evaluations per RHS instead of 13) is bit-identical.
#pragma once
#include "rapt_fields.cuh"
#include "dop_const.cuh"

namespace RAPT_NS {

using rapt::ParamsP;
using rapt::AdvArgs;

// |B| and b at one point.  tf: time factor (times the field's constant) of a separable field, evaluated once by the caller.
template <class F>
RAPT_DEV void mag_and_unit(const FieldP &f, double t, double tf, double x, double y, double z,
                           double &m, double &ux, double &uy, double &uz)
{
    double bx, by, bz;
    if (F::SEPARABLE) F::Bspace(f, tf, x, y, z, bx, by, bz);
    else F::B(f, t, x, y, z, bx, by, bz);
#if RAPT_STRICT
    m = sqrt(dot3(bx, by, bz, bx, by, bz));
    ux = bx / m; uy = by / m; uz = bz / m;
#else
    const double bb = dot3(bx, by, bz, bx, by, bz), im = fast_rsqrt(bb);
    m = bb * im;
    ux = bx * im; uy = by * im; uz = bz * im;
#endif
}

// gradB (fields.py:125-131) and curlb (fields.py:191-200) from one pass over the six shifted points
template <class F>
RAPT_DEV void grad_and_curl(const FieldP &f, double t, double tf, double x, double y, double z,
                            double (&g)[3], double (&c)[3])
{
    if (F::UNIFORM) { g[0] = g[1] = g[2] = 0; c[0] = c[1] = c[2] = 0; return; }
    const double d = f.gradstep;
    double mp, mm, pxx, pxy, pxz, mxx, mxy, mxz, pyx, pyy, pyz, myx, myy, myz, pzx, pzy, pzz, mzx, mzy, mzz;
#if RAPT_STRICT
    const double den = 2 * d;
#define RAPT_FD(a, b) (((a) - (b)) / den)
#else
    const double i2d = 1.0 / (2 * d);
#define RAPT_FD(a, b) (((a) - (b)) * i2d)
#endif
    mag_and_unit<F>(f, t, tf, x + d, y, z, mp, pxx, pxy, pxz); mag_and_unit<F>(f, t, tf, x - d, y, z, mm, mxx, mxy, mxz);
    g[0] = RAPT_FD(mp, mm);
    mag_and_unit<F>(f, t, tf, x, y + d, z, mp, pyx, pyy, pyz); mag_and_unit<F>(f, t, tf, x, y - d, z, mm, myx, myy, myz);
    g[1] = RAPT_FD(mp, mm);
    mag_and_unit<F>(f, t, tf, x, y, z + d, mp, pzx, pzy, pzz); mag_and_unit<F>(f, t, tf, x, y, z - d, mm, mzx, mzy, mzz);
    g[2] = RAPT_FD(mp, mm);
    (void)pxx; (void)mxx; (void)pyy; (void)myy; (void)pzz; (void)mzz;
#if RAPT_STRICT
    // summation order of np.dot(_M1, beta) in the reference's BLAS ('oracle')
    c[0] = ((pyz + (-myz + -pzy)) + mzy) / den;
    c[1] = ((-pxz + pzx) + (mxz + -mzx)) / den;
    c[2] = ((pxy + myx) + (-mxy + -pyx)) / den;
#else
    c[0] = ((pyz - myz) - (pzy - mzy)) * i2d;
    c[1] = ((pzx - mzx) - (pxz - mxz)) * i2d;
    c[2] = ((pxy - mxy) - (pyx - myx)) * i2d;
#endif
#undef RAPT_FD
}

struct GcConst { double mass, q, mu, v; };

// GuidingCenter._TaoChanBrizardEOM :329-355, _BrizardChanEOM :357-379, _NorthropTellerEOM :381-395
template <class F>
RAPT_DEV void gc_rhs(const FieldP &f, const GcConst &c, int eom, int equatorial,
                     double t, const double (&Y)[4], double (&out)[4])
{
    const double m = c.mass, q = c.q, mu = c.mu, ppar = Y[3];
    double bx, by, bz, gB[3], cb[3];
    double tf = 1.0;
    if (F::SEPARABLE) { tf = F::tfactor(f, t); F::Bspace(f, tf, Y[0], Y[1], Y[2], bx, by, bz); }
    else F::B(f, t, Y[0], Y[1], Y[2], bx, by, bz);
#if !RAPT_STRICT
    // fast flavour: one rsqrt for |B| and b, reciprocals instead of divisions
    const double bb = dot3(bx, by, bz, bx, by, bz), ib = fast_rsqrt(bb);
    const double Bmag = bb * ib;
    const double ux = bx * ib, uy = by * ib, uz = bz * ib;
    // divisions by the tracer's constants (m, q, c) as multiplications by branch-free reciprocals: a true fp64
    // division is ~12 FP64-pipe instructions plus a slow-path call site, three of them per right-hand side
    const double iq = fast_rcp(q), im = fast_rcp(m);
    const double ic = 1.0 / RAPT_C_LIGHT;                          // folded at compile time
    grad_and_curl<F>(f, t, tf, Y[0], Y[1], Y[2], gB, cb);
    if (eom == 0) {
        const double imc = im * ic;
        const double pm = ppar * imc;
        const double g2 = fma(pm, pm, fma(2 * mu * Bmag, imc * ic, 1.0));
        const double ig = fast_rsqrt(g2);                          // 1/gamma
        const double pq = ppar * iq;
        const double Bsx = fma(pq, cb[0], bx), Bsy = fma(pq, cb[1], by), Bsz = fma(pq, cb[2], bz);
        const double iBsp = fast_rcp(dot3(Bsx, Bsy, Bsz, ux, uy, uz));
        double ex = 0, ey = 0, ez = 0, dbx = 0, dby = 0, dbz = 0;
        if (F::HAS_E) F::E(f, t, Y[0], Y[1], Y[2], ex, ey, ez);
        // a separable field changes its magnitude, not its direction: db/dt = 0 (the reference's central
        // difference of unit vectors returns pure round-off, ~1e-13, there)
        if (F::TIME_DEP && !F::SEPARABLE) { if (!f.is_static) F::dbdt(f, t, Y[0], Y[1], Y[2], dbx, dby, dbz); }
        const double mg = mu * ig;
        const double Esx = ex - (ppar * dbx + mg * gB[0]) * iq;
        const double Esy = ey - (ppar * dby + mg * gB[1]) * iq;
        const double Esz = ez - (ppar * dbz + mg * gB[2]) * iq;
        const double cx = Esy * uz - Esz * uy, cy = Esz * ux - Esx * uz, cz = Esx * uy - Esy * ux;
        const double pgm = ppar * ig * im;
        out[0] = fma(pgm, Bsx, cx) * iBsp;
        out[1] = fma(pgm, Bsy, cy) * iBsp;
        out[2] = fma(pgm, Bsz, cz) * iBsp;
        out[3] = q * dot3(Esx, Esy, Esz, Bsx, Bsy, Bsz) * iBsp;
    } else if (eom == 1) {
        const double vc = c.v * ic;
        const double ig = sqrt(1 - vc * vc);                       // 1/gamma
        const double pq = ppar * iq;
        const double Bsx = fma(pq, cb[0], bx), Bsy = fma(pq, cb[1], by), Bsz = fma(pq, cb[2], bz);
        const double iBsp = fast_rcp(dot3(Bsx, Bsy, Bsz, ux, uy, uz));
        const double cx = uy * gB[2] - uz * gB[1], cy = uz * gB[0] - ux * gB[2], cz = ux * gB[1] - uy * gB[0];
        const double pgm = ppar * ig * im, mqg = mu * iq * ig;
        out[0] = fma(pgm, Bsx, mqg * cx) * iBsp;
        out[1] = fma(pgm, Bsy, mqg * cy) * iBsp;
        out[2] = fma(pgm, Bsz, mqg * cz) * iBsp;
        out[3] = -mu * dot3(Bsx, Bsy, Bsz, gB[0], gB[1], gB[2]) * ig * iBsp;
    } else {
        const double vc = c.v * ic;
        const double ig = sqrt(1 - vc * vc);
        const double igm = ig * im, gm = m / ig;
        const double cx = uy * gB[2] - uz * gB[1], cy = uz * gB[0] - ux * gB[2], cz = ux * gB[1] - uy * gB[0];
        const double s = (gm * (c.v * c.v) + ppar * ppar * igm) * 0.5 * iq * (ib * ib);
        const double pg = ppar * igm;
        out[0] = fma(s, cx, pg * ux);
        out[1] = fma(s, cy, pg * uy);
        out[2] = fma(s, cz, pg * uz);
        out[3] = -mu * dot3(ux, uy, uz, gB[0], gB[1], gB[2]) * ig;
    }
#else
    const double Bmag = sqrt(dot3(bx, by, bz, bx, by, bz));
    const double ux = bx / Bmag, uy = by / Bmag, uz = bz / Bmag;
    grad_and_curl<F>(f, t, tf, Y[0], Y[1], Y[2], gB, cb);
    if (eom == 0) {
        const double pm = ppar / (m * RAPT_C_LIGHT);
        const double gamma = sqrt(1 + 2 * mu * Bmag / (m * RAPT_C_LIGHT * RAPT_C_LIGHT) + RAPT_SQ(pm));
        const double Bsx = bx + ppar * cb[0] / q, Bsy = by + ppar * cb[1] / q, Bsz = bz + ppar * cb[2] / q;
        const double Bsp = dot3(Bsx, Bsy, Bsz, ux, uy, uz);
        double ex = 0, ey = 0, ez = 0, dbx = 0, dby = 0, dbz = 0;
        if (F::HAS_E) F::E(f, t, Y[0], Y[1], Y[2], ex, ey, ez);
        if (F::TIME_DEP) { if (!f.is_static) F::dbdt(f, t, Y[0], Y[1], Y[2], dbx, dby, dbz); }
        const double Esx = ex - (ppar * dbx + mu * gB[0] / gamma) / q;
        const double Esy = ey - (ppar * dby + mu * gB[1] / gamma) / q;
        const double Esz = ez - (ppar * dbz + mu * gB[2] / gamma) / q;
        const double cx = Esy * uz - Esz * uy, cy = Esz * ux - Esx * uz, cz = Esx * uy - Esy * ux;
        const double gmm = gamma * m;
        out[0] = (ppar * Bsx / gmm + cx) / Bsp;
        out[1] = (ppar * Bsy / gmm + cy) / Bsp;
        out[2] = (ppar * Bsz / gmm + cz) / Bsp;
        out[3] = q * dot3(Esx, Esy, Esz, Bsx, Bsy, Bsz) / Bsp;
    } else if (eom == 1) {
        const double vc = c.v / RAPT_C_LIGHT;
        const double gamma = 1.0 / sqrt(1 - RAPT_SQ(vc));
        const double Bsx = bx + ppar * cb[0] / q, Bsy = by + ppar * cb[1] / q, Bsz = bz + ppar * cb[2] / q;
        const double Bsp = dot3(Bsx, Bsy, Bsz, ux, uy, uz);
        const double cx = uy * gB[2] - uz * gB[1], cy = uz * gB[0] - ux * gB[2], cz = ux * gB[1] - uy * gB[0];
        const double gmm = gamma * m, qg = q * gamma;
        out[0] = (ppar * Bsx / gmm + mu * cx / qg) / Bsp;
        out[1] = (ppar * Bsy / gmm + mu * cy / qg) / Bsp;
        out[2] = (ppar * Bsz / gmm + mu * cz / qg) / Bsp;
        out[3] = -mu * dot3(Bsx, Bsy, Bsz, gB[0], gB[1], gB[2]) / (gamma * Bsp);
    } else {
        const double vc = c.v / RAPT_C_LIGHT;
        const double gamma = 1.0 / sqrt(1 - RAPT_SQ(vc));
        const double gm = gamma * m;
        const double cx = uy * gB[2] - uz * gB[1], cy = uz * gB[0] - ux * gB[2], cz = ux * gB[1] - uy * gB[0];
        const double s = (gm * RAPT_SQ(c.v) + RAPT_SQ(ppar) / gm) / (2 * q * RAPT_SQ(Bmag));
        out[0] = s * cx + ppar * ux / gm;
        out[1] = s * cy + ppar * uy / gm;
        out[2] = s * cz + ppar * uz / gm;
        out[3] = -mu * dot3(ux, uy, uz, gB[0], gB[1], gB[2]) / gamma;
    }
#endif
    if (equatorial) { out[2] = 0; out[3] = 0; }
}

// utils.cyclotron_radius2 (utils.py:183-187) given |B|
RAPT_DEV double cycrad2(double Bmag, double vpar, double v, double mass, double q)
{
    double vc = v / RAPT_C_LIGHT;
    double gamma = 1.0 / sqrt(1 - vc * vc);
    double vperp = sqrt((v - vpar) * (v + vpar));
    return gamma * mass * vperp / (fabs(q) * Bmag);
}

// GuidingCenter.isadiabatic :323-327 with cycrad :517-529 and cycper :531-541 (quirk Q10 kept)
template <class F>
RAPT_DEV bool gc_isadiabatic(const FieldP &f, const ParamsP &p, double t, const double (&y)[4],
                             double mu, double mass, double q)
{
    const double pp = y[3];
#if !RAPT_STRICT
    // fast flavour: rho / L < epss without the reference's chain of quotients.  (gamma m)^2 (v^2 - v_par^2) = 2 m mu B in
    // both of its branches, times the 1 / (1 - v^2 / c^2) of cyclotron_radius2 in the non-relativistic one (quirk kept), so
    //   rho / L < epss  <=>  2 m mu G maxdiff^2 < (epss |q| 2d)^2 B^3.
    {
        double bx, by, bz; F::B(f, t, y[0], y[1], y[2], bx, by, bz);
        const double B2 = dot3(bx, by, bz, bx, by, bz), B1 = B2 * fast_rsqrt(B2);
        const double ic = 1.0 / RAPT_C_LIGHT, im = fast_rcp(mass);
        const double pmc = pp * im * ic, w = 2 * mu * B1 * im;                 // w = 2 mu B / m
        const double g2 = fma(pmc, pmc, fma(w, ic * ic, 1.0));
        double G = 1.0;
        if (g2 < 1.000002000001) G = fast_rcp(1.0 - fma(pp * im, pp * im, w) * (ic * ic));   // gamma - 1 < 1e-6 (:521-523)
        const double md = F::max_central_difference(f, t, y[0], y[1], y[2]);
        const double lim = p.epss * fabs(q) * (2 * f.gradstep);
        const bool sp = (2 * mass * mu * G) * (md * md) < (lim * lim) * (B2 * B1);
        if (f.is_static || !sp) return sp;
    }
#endif
    const double Bmag = F::magB(f, t, y[0], y[1], y[2]);
    const double pmc = pp / mass / RAPT_C_LIGHT;
    const double gamma = sqrt(1 + 2 * mu * Bmag / (mass * RAPT_C_LIGHT * RAPT_C_LIGHT) + pmc * pmc);
    double vp, v;
    if (gamma - 1 < 1e-6) { vp = pp / mass; v = sqrt(2 * mu * Bmag / mass + vp * vp); }
    else { vp = pp / mass / gamma; v = RAPT_C_LIGHT * sqrt(1 - 1 / (gamma * gamma)); }
    const double rho = cycrad2(Bmag, vp, v, mass, q);
    bool sp = rho / F::lengthscale(f, t, y[0], y[1], y[2]) < p.epss;
    if (f.is_static || !sp) return sp;
    const double g2 = sqrt(1 + 2 * mu * Bmag / (mass * RAPT_C_LIGHT * RAPT_C_LIGHT) + pp * pp);
    double v2;
    if (g2 - 1 < 1e-6) { double vq = pp / mass; v2 = sqrt(2 * mu * Bmag / mass + vq * vq); }
    else v2 = RAPT_C_LIGHT * sqrt(1 - 1 / (g2 * g2));
    const double vc = v2 / RAPT_C_LIGHT;
    const double per = 2 * RAPT_PI * (1.0 / sqrt(1 - vc * vc)) * mass / Bmag / fabs(q);
    return per / F::timescale(f, t, y[0], y[1], y[2]) < p.epst;
}

#if RAPT_STRICT
#define RAPT_GC_POW(x, e) pow((x), (e))
#else
#define RAPT_GC_POW(x, e) exp((e) * log(x))     // x > 0; ~1e-15 relative, no slow paths
#endif

#ifndef RAPT_GC_DEFER
#define RAPT_GC_DEFER 0      /* 1: two tracers per lane, HINIT probes batched across the warp (see k_gc_dopri5); measured, does not pay: profiles/r2_gc_batched_probes.md */
#endif
#ifndef RAPT_GC_DEFER_MIN
#define RAPT_GC_DEFER_MIN 16 /* lanes of a warp that must wait for a probe before the probe slot is run */
#endif
#define RAPT_GC_PARK_WORDS 27
#ifndef RAPT_GC_THREADS
#define RAPT_GC_THREADS 512  /* threads per block of a launch that fills the GPU (kernels_tu.cu:block_threads); MINB counts resident 128-thread
                                units per SM, so 512 is one block per SM at MINB = 4: 212.4 vs 215.6 ms on config 3, 383.5 vs 389.4 on config 5 */
#endif

// MINB = resident CTAs per SM the register allocation is tuned for (2: 255 regs, no spills; 3: 168 regs; 4: 128 regs, ~0.5 KB of spill traffic per step, fastest: profiles/r1_other_configs.md)
//
// Batched HINIT probes (RAPT_GC_DEFER = 1; fast flavour, analytic fields; OFF by default).  A row takes ~9 steps, so in
// every iteration ~11 % of the lanes of a warp start a row and need HINIT's Euler probe -- one more right-hand side, which
// the whole warp then sits through at ~3.5 of 32 lanes (ncu: the right-hand side ran 7.25 times per iteration at 26.8
// lanes, profiles/r2_01_gc_ncu_summary.txt).  With the switch on every lane owns TWO tracers: one in registers, one parked
// in shared memory.  A lane whose tracer reaches a row boundary parks it and steps its other tracer; the probe slot is only
// run when RAPT_GC_DEFER_MIN lanes of the warp have a tracer waiting (or lanes would idle otherwise).  Every tracer's
// arithmetic is unchanged -- only WHEN its probe runs (tests/test_kernel_host.py: same bits as the build without it).
// Measured in three A/B calls (profiles/r2_gc_batched_probes.md): the right-hand side runs 7.5 % less often at 28.9 lanes
// and the kernel executes 5 % fewer instructions, but the 27 KB of parking space per block come out of the L1 that this
// kernel's 0.4 KB/thread of register spills live in, and the exchange code pushes the hot loop past the instruction cache:
// FP64 pipe 80.5 % -> 69-71 % active, 1.5-3.7 % SLOWER on config 3, 0.2-4.6 % on config 5.  Kept as a switch.
template <class F, int MINB>
__global__ void __launch_bounds__(RAPT_GC_THREADS, (MINB * 128 >= RAPT_GC_THREADS) ? (MINB * 128) / RAPT_GC_THREADS : 1) k_gc_dopri5(const AdvArgs a)
{
    constexpr bool DEFER = RAPT_GC_DEFER && !RAPT_STRICT && !F::CAN_FAIL;
    if (F::CAN_FAIL) grid_cache_reset();     // gridded field: per-thread cell cache (rapt_fields.cuh)
    const double rtol = a.p.rtol, atol = a.p.atol;
    const int eqf = a.p.enforce_equatorial, eom = a.eom;
    const double beta = 0.04, safe = 0.9, fac1 = 0.2, fac2 = 10.0, uround = 2.3e-16;
    const double expo1 = 0.2 - beta * 0.75, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    const double pf0 = pow(1e-4, beta);          // facold^beta at the first step of every row

    double y[4], k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], y1[4], yin[4], kout[4];
    double x = 0, h = 0, xend = 0, tstop = 0, tlim = 0, dt = 0, facold = 1e-4, hmax = 0, tin = 0, dnf = 0;
    GcConst gc = {0, 0, 0, 0};
    int pid = -1;
    int nstep = 0, naccpt = 0, nrejct = 0, ncalls = 0, nstep_row = 0, naccpt_row = 0;
    int rowidx = 0, nst = 0, st = RAPT_ST_OK;
    bool last = false, reject = false, need_row = false, have = false;
    double *myrows = nullptr;
    // the parked tracer of this lane: word w of thread i at park[w * blockDim + i] (conflict-free)
    __shared__ double park[DEFER ? RAPT_GC_PARK_WORDS * RAPT_GC_THREADS : 1];
    bool phave = false, pneed = false;           // a tracer is parked / it waits for its HINIT probe (else it is mid-row)
    (void)pf0;

#define GC_SWD(v, w) { double *s_ = &park[(w) * RAPT_GC_THREADS + threadIdx.x]; const double t_ = *s_; *s_ = (v); (v) = t_; }
#define GC_SWI(i0, i1, w) { double *s_ = &park[(w) * RAPT_GC_THREADS + threadIdx.x]; const double t_ = *s_;                      \
                            *s_ = __hiloint2double((i0), (i1)); (i0) = __double2hiint(t_); (i1) = __double2loint(t_); }
    // exchange the register tracer with the parked one (either may be empty)
#define GC_SWAP_SLOTS()                                                                                                       \
    {                                                                                                                         \
        GC_SWD(y[0], 0) GC_SWD(y[1], 1) GC_SWD(y[2], 2) GC_SWD(y[3], 3)                                                       \
        GC_SWD(k1[0], 4) GC_SWD(k1[1], 5) GC_SWD(k1[2], 6) GC_SWD(k1[3], 7)                                                   \
        GC_SWD(x, 8) GC_SWD(h, 9) GC_SWD(xend, 10) GC_SWD(tstop, 11) GC_SWD(tlim, 12) GC_SWD(dt, 13) GC_SWD(facold, 14)       \
        GC_SWD(hmax, 15) GC_SWD(gc.mass, 16) GC_SWD(gc.q, 17) GC_SWD(gc.mu, 18) GC_SWD(gc.v, 19)                              \
        { double mr_ = __longlong_as_double((long long)(size_t)myrows); GC_SWD(mr_, 20)                                       \
          myrows = (double *)(size_t)__double_as_longlong(mr_); }                                                             \
        GC_SWI(pid, nstep, 21) GC_SWI(naccpt, nrejct, 22) GC_SWI(ncalls, nstep_row, 23) GC_SWI(naccpt_row, rowidx, 24)        \
        GC_SWI(nst, st, 25)                                                                                                   \
        { int fl_ = (last ? 1 : 0) | (reject ? 2 : 0), zero_ = 0; GC_SWI(fl_, zero_, 26)                                      \
          last = (fl_ & 1) != 0; reject = (fl_ & 2) != 0; }                                                                   \
        { const bool h_ = have, n_ = need_row; have = phave; need_row = pneed; phave = h_; pneed = n_; }                      \
    }

    bool done = false, more_work = true;
    for (;;) {
        // ---- (A) guiding centre finished?  write it back
        if (have && need_row && !(st == RAPT_ST_OK && x < tlim)) {
            if (a.seg_tstop) {                                   // sliced adaptive epoch
                a.seg_row[pid] = rowidx;
                if (st == RAPT_ST_OK && x < tstop) st = RAPT_ST_SLICE;
            }
            a.t[pid] = x; a.s1[pid] = y[0]; a.s2[pid] = y[1]; a.s3[pid] = y[2]; a.s4[pid] = y[3];
            int *c = a.counters + 4 * (long long)pid;
            int nf = 2 * ncalls + 6 * nstep;                     // as scipy counts: SURVEY.md §3.1
            if (a.append) { c[0] += nf; c[1] += nstep; c[2] += naccpt; c[3] += nrejct; }
            else { c[0] = nf; c[1] = nstep; c[2] = naccpt; c[3] = nrejct; }
            a.status[pid] = st;
            a.tcur[pid] = x;                                     // GuidingCenter.py:456
            if (a.nrows) a.nrows[pid] = rowidx + 1;
            a.nstored[pid] = nst;
            have = false;
        }
        // ---- (A') two tracers per lane.  ONE exchange per iteration in front of the fetch (each expansion of the exchange
        // is ~100 instructions and the hot loop sits at the instruction-cache size):
        //  * the register tracer waits for its probe and the park slot is free while more than one tracer per lane is
        //    still queued: park it, the fetch below brings another one;
        //  * the register tracer waits for its probe and the parked one is mid-row: step that one;
        //  * the register slot is empty: continue with the parked tracer, unless it waits for a probe and another tracer
        //    can be fetched first.
        if (DEFER) {
            // the queue counter is only looked at by a lane that has to decide (a read of that contended line every
            // iteration cost 0.8 long-scoreboard stalls per issue), and never again once little work is left
            if (more_work && ((have && need_row && !phave) || (!have && phave && pneed)))
                more_work = (a.nwork - *(volatile int *)a.queue) > (int)(gridDim.x * blockDim.x);
            if ((have && need_row && (phave ? !pneed : more_work)) || (!have && phave && (!pneed || !more_work))) GC_SWAP_SLOTS()
        }
        if (!have && !done) {
            bool got = false;
            if (!DEFER || !phave || more_work) {
                int w = atomicAdd(a.queue, 1);
                if (w < a.nwork) {
                    got = true;
                    pid = a.order ? a.order[w] : w;
                    x = a.t[pid];
                    y[0] = a.s1[pid]; y[1] = a.s2[pid]; y[2] = a.s3[pid]; y[3] = a.s4[pid];
                    gc.mass = a.mass[pid]; gc.q = a.charge[pid]; gc.mu = a.mu[pid]; gc.v = a.v[pid];
                    dt = a.dtin[pid];
                    double delta = a.delta_arr ? a.delta_arr[pid] : a.delta;
                    tstop = x + delta;                                   // GuidingCenter.py:452
                    tlim = tstop;
                    int row0 = 0;
                    if (a.seg_tstop) { tstop = a.seg_tstop[pid]; tlim = fmin(tstop, a.slice_end); row0 = a.seg_row[pid]; }
                    nstep = naccpt = nrejct = ncalls = 0; rowidx = row0; st = RAPT_ST_OK;
                    myrows = a.rows ? a.rows + (size_t)pid * (size_t)a.max_rows * 8 : nullptr;
                    if (a.append) nst = a.nstored[pid];
                    else {
                        nst = 0;
                        if (myrows && a.store_every > 0 && a.max_rows > 0) {
                            double2 *r = reinterpret_cast<double2 *>(myrows);
                            r[0] = make_double2(x, y[0]); r[1] = make_double2(y[1], y[2]);
                            r[2] = make_double2(y[3], gc.mu); r[3] = make_double2(0.0, 0.0);
                            nst = 1;
                        }
                    }
                    have = true; need_row = true;
                    // degenerate output step (B = 0 or inf, bad resolution): the reference would never return;
                    // delta <= 0 (or beyond this slice): nothing to do.  Both are retired at the top of the next iteration.
                    if (!(dt > 0.0) || dt > 1e300) st = RAPT_ST_HSMALL;
                    else if (x < tlim) gc_rhs<F>(a.f, gc, eom, eqf, x, y, k1);              // k1 = f(x, y)
                }
            }
            // nothing fetched: with a parked tracer (it waits for a probe: the probe slot below brings it in) the lane goes on
            if (!got && !(DEFER && phave)) done = true;
        }
        if (DEFER) {
            if (__all_sync(0xffffffffu, done)) break;            // lanes leave together (the votes below are warp-wide)
        } else if (done) break;
        // a tracer that was just fetched but cannot run (bad dt, nothing to do) waits for its retirement
        const bool dead = have && need_row && !(st == RAPT_ST_OK && x < tlim);
        // ---- (V) does this iteration run the probe slot?
        bool do_probe = true;
        if (DEFER) {
            const bool a_np = have && need_row && !dead;
            const unsigned m_np = __ballot_sync(0xffffffffu, a_np || (phave && pneed));
            const unsigned m_rd = __ballot_sync(0xffffffffu, have && !need_row);
            const unsigned m_any = __ballot_sync(0xffffffffu, (have && !dead) || phave);
            // lanes that cannot step this iteration unless the probe runs (no second tracer to turn to)
            const unsigned m_idle = __ballot_sync(0xffffffffu, (a_np || (phave && pneed)) && !(have && !need_row));
            do_probe = m_np != 0 && (__popc(m_np) >= RAPT_GC_DEFER_MIN || __popc(m_idle) >= 4 || 2 * __popc(m_rd) < __popc(m_any));
            if (do_probe && !a_np && !dead && phave && pneed) GC_SWAP_SLOTS()      // this lane's parked tracer takes the probe
        }
        // ---- (B) one step attempt; stage 1 = HINIT's Euler probe for lanes that start an output row
        bool skip = !have || dead || (need_row && !do_probe), hin = false, rowdone = false;
#pragma unroll 1
        for (int s = (DEFER && !do_probe) ? 2 : 1; s <= 7; s++) {
            bool active = !skip;
            switch (s) {
            case 1:
                active = !skip && need_row;
                if (active) {
                    // new output row = new solver call: HINIT part 1
                    xend = x + dt;                               // GuidingCenter.py:453
                    hmax = fabs(xend - x);
                    double dny = 0;
                    dnf = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
#if RAPT_STRICT
                        double sk = atol + rtol * fabs(y[i]);
                        dnf += (k1[i] / sk) * (k1[i] / sk);
                        dny += (y[i] / sk) * (y[i] / sk);
#else
                        const double is_ = fast_rcp(atol + rtol * fabs(y[i]));
                        const double a_ = k1[i] * is_, b_ = y[i] * is_;
                        dnf += a_ * a_; dny += b_ * b_;
#endif
                    }
#if RAPT_STRICT
                    h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
#else
                    h = 1e-6;
                    if (!(dnf <= 1e-10 || dny <= 1e-10)) { const double qq = dny * fast_rcp(dnf); h = (qq * fast_rsqrt(qq)) * 0.01; }
#endif
                    h = fmin(h, hmax);
#pragma unroll
                    for (int i = 0; i < 4; i++) yin[i] = y[i] + h * k1[i];
                    tin = x + h;
                    hin = true;
                }
                break;
            case 2:
                if (skip) break;
                if (hin && F::CAN_FAIL && !(kout[0] == kout[0] && kout[3] == kout[3])) {
                    // gridded field: the probe point lies outside the grid, where the reference's interpolator raises from
                    // inside r.integrate() (no row for this call, the rows so far are kept)
                    st = RAPT_ST_FIELD; skip = true; active = false;
                    break;
                }
                if (hin) {
                    double der2 = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
#if RAPT_STRICT
                        double sk = atol + rtol * fabs(y[i]);
                        der2 += ((kout[i] - k1[i]) / sk) * ((kout[i] - k1[i]) / sk);
#else
                        const double d_ = (kout[i] - k1[i]) * fast_rcp(atol + rtol * fabs(y[i]));
                        der2 += d_ * d_;
#endif
                    }
#if RAPT_STRICT
                    der2 = sqrt(der2) / h;
                    double der12 = fmax(fabs(der2), sqrt(dnf));
                    double h1;
                    if (der12 <= 1e-15) h1 = fmax(1e-6, fabs(h) * 1e-3);
                    else {
                        // (0.01/der12)^(1/5) only matters when it is the smallest candidate
                        double tq = 0.01 / der12, hm2 = hmax * hmax;
                        if (tq > hm2 * hm2 * hmax * 1.000001) h1 = hmax;
                        else h1 = pow(tq, 1.0 / 5.0);
                    }
#else
                    // der12^2 = max(der2/h^2, dnf); h1 = (0.01/der12)^(1/5) = exp((log 0.01 - log(der12^2)/2)/5):
                    // no square root or division, branch-free log/exp
                    const double ih = fast_rcp(h);
                    const double d12 = fmax(der2 * ih * ih, dnf);
                    double h1;
                    if (d12 <= 1e-30) h1 = fmax(1e-6, fabs(h) * 1e-3);
                    else h1 = fast_exp(0.2 * fma(-0.5, fast_log(d12), -4.605170185988091));
#endif
                    h = fmin(fmin(100 * fabs(h), h1), hmax);
                    facold = RAPT_STRICT ? 1e-4 : beta * -9.210340371976182; last = false; reject = false; nstep_row = 0; naccpt_row = 0;
                    ncalls++;
                    need_row = false;
                }
                if (nstep_row > 500) st = RAPT_ST_NMAX;
                else if (0.1 * fabs(h) <= fabs(x) * uround) st = RAPT_ST_HSMALL;
                if (st != RAPT_ST_OK) {
                    // solver failure: the state reached is still appended as a row, labelled with the time reached, before
                    // `while r.successful()` ends the loop (GuidingCenter.py:452-456)
                    rowdone = true; skip = true; active = false;
                } else {
                    if ((x + 1.01 * h - xend) > 0.0) { h = xend - x; last = true; }
                    nstep_row++; nstep++;
#pragma unroll
                    for (int i = 0; i < 4; i++) yin[i] = y[i] + h * T5(A2_1) * k1[i];
                    tin = x + T5(C2) * h;
                }
                break;
            case 3:
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { k2[i] = kout[i]; yin[i] = y[i] + h * (T5(A3_1) * k1[i] + T5(A3_2) * k2[i]); }
                    tin = x + T5(C3) * h;
                }
                break;
            case 4:
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { k3[i] = kout[i]; yin[i] = y[i] + h * (T5(A4_1) * k1[i] + T5(A4_2) * k2[i] + T5(A4_3) * k3[i]); }
                    tin = x + T5(C4) * h;
                }
                break;
            case 5:
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { k4[i] = kout[i]; yin[i] = y[i] + h * (T5(A5_1) * k1[i] + T5(A5_2) * k2[i] + T5(A5_3) * k3[i] + T5(A5_4) * k4[i]); }
                    tin = x + T5(C5) * h;
                }
                break;
            case 6:
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) { k5[i] = kout[i]; yin[i] = y[i] + h * (T5(A6_1) * k1[i] + T5(A6_2) * k2[i] + T5(A6_3) * k3[i] + T5(A6_4) * k4[i] + T5(A6_5) * k5[i]); }
                    tin = x + h;
                }
                break;
            default:  // 7: the new state; its right-hand side is both the 7th stage and the next k1 (FSAL)
                if (active) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        k6[i] = kout[i];
                        y1[i] = y[i] + h * (T5(A7_1) * k1[i] + T5(A7_3) * k3[i] + T5(A7_4) * k4[i] + T5(A7_5) * k5[i] + T5(A7_6) * k6[i]);
                        yin[i] = y1[i];
                    }
                    tin = x + h;
                }
                break;
            }
            if (active) gc_rhs<F>(a.f, gc, eom, eqf, tin, yin, kout);
        }
        if (!skip) {
            // ---- error estimate and step control
            double err = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double e = (T5(E1) * k1[i] + T5(E3) * k3[i] + T5(E4) * k4[i] + T5(E5) * k5[i] + T5(E6) * k6[i] + T5(E7) * kout[i]) * h;
#if RAPT_STRICT
                double sk = atol + rtol * fmax(fabs(y[i]), fabs(y1[i]));
                err += (e / sk) * (e / sk);
#else
                e *= fast_rcp(atol + rtol * fmax(fabs(y[i]), fabs(y1[i])));
                err += e * e;
#endif
            }
#if RAPT_STRICT
            err = sqrt(err / 4);
#else
            err *= 0.25;
            err = (err > 0.0) ? err * fast_rsqrt(err) : 0.0;
#endif
            if (err <= 1.0) {
                double hnew = h;
                if (!last) {
#if RAPT_STRICT
                    double fac11 = RAPT_GC_POW(err, expo1);
                    double fac = fac11 / ((facold == 1e-4) ? pf0 : RAPT_GC_POW(facold, beta));
#else
                    // err^expo1 / facold^beta as one branch-free log and one exp; facold is carried as
                    // beta * log(facold) (only read when the row continues, so only updated here)
                    const double lg = fast_log(fmax(err, 1e-300));
                    double fac = fast_exp(fma(expo1, lg, -facold));
                    facold = beta * fmax(lg, -9.210340371976182);
#endif
                    fac = fmax(facc2, fmin(facc1, fac / safe));
                    hnew = h / fac;
                    if (fabs(hnew) > hmax) hnew = hmax;
                    if (reject) hnew = fmin(fabs(hnew), fabs(h));
                }
#if RAPT_STRICT
                facold = fmax(err, 1e-4);
#endif
                naccpt++; naccpt_row++;
#pragma unroll
                for (int i = 0; i < 4; i++) { k1[i] = kout[i]; y[i] = y1[i]; }
                x = x + h;
                if (last) rowdone = true;
                else {
                    h = hnew;
                    reject = false;
                }
            } else {
#if RAPT_STRICT
                double fac11 = RAPT_GC_POW(err, expo1);
#else
                double fac11 = fast_exp(expo1 * fast_log(fmin(err, 1e300)));
#endif
                h = h / fmin(facc1, fac11 / safe);
                reject = true;
                if (naccpt_row >= 1) nrejct++;
                last = false;
                if (F::CAN_FAIL && !(err == err)) { st = RAPT_ST_FIELD; need_row = true; }   // left the grid: keep the last row
            }
        }
        if (rowdone) {
            // ---- output row complete (GuidingCenter.py:453-458)
            rowidx++;
            if (myrows && a.store_every > 0 && (rowidx % a.store_every) == 0 && nst < a.max_rows) {
                double2 *r = reinterpret_cast<double2 *>(myrows + (size_t)nst * 8);
                double tag = a.segtag ? (double)a.segtag[pid] : (double)nstep;
                r[0] = make_double2(x, y[0]); r[1] = make_double2(y[1], y[2]);
                r[2] = make_double2(y[3], gc.mu); r[3] = make_double2(0.0, tag);
                nst++;
            }
            if (a.p.check_adiabaticity) {
                if (!gc_isadiabatic<F>(a.f, a.p, x, y, gc.mu, gc.mass, gc.q)) st = RAPT_ST_NONADIABATIC;
            }
            need_row = true;
        }
    }
#undef GC_SWD
#undef GC_SWI
#undef GC_SWAP_SLOTS
}

}  // namespace RAPT_NS
