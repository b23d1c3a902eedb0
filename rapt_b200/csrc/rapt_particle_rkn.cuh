// rapt_particle_rkn.cuh -- Particle.advance for STATIC fields in Nystrom form (fast flavour only).
//
// In a static field the reference freezes gamma*m at its entry value (Particle.py:286-291, quirk Q3),
// so dx/dt = p/(gamma m) is linear in the momentum and the position part of every DOP853 stage follows
// from the stage *momentum derivatives* K_l alone:
//     P_i = p + h sum_j a_ij K_j                    X_i = x + (h/gm) (rs_i p + h sum_l (A.A)_il K_l)
//     p'  = p + h sum_j b_j K_j                     x'  = x + (h/gm) (sb p + h sum_l (b.A)_l K_l)
// and likewise for the two error estimators (tools/gen_coeffs.py forms A.A, b.A, er.A, w.A exactly from the
// double tableau and rounds once).  Only 3 of the 6 components of each k-vector are stored: 33 doubles
// instead of 60, 128 registers instead of 255, so 16 warps/SM instead of 8 hide the FP64 dependency
// latency of the field evaluation (profiles/r1_particle_history.md).  Same step-control decisions as
// k_particle_dop853: the estimators differ from the 6-component form only by its own cancellation
// round-off (~1e-10 relative of err), positions by ~1 ulp per step.
// Not used when params["enforce equatorial"] is set or the field is not static (gamma m then varies).
//
// The step is STRAIGHT-LINE code: with 3-component vectors and the branch-free rsqrt the 12 stages, the new
// state and both error estimators are ~1.1k SASS instructions, the 13 copies of the field evaluation write
// straight into their K_l registers (no copies, no switch/branch overhead) and the loop still streams from
// the instruction cache (profiles/r1_particle_history.md: rolled 482 ms, unrolled 394 ms, this form 269 ms).
#pragma once
#include "rapt_particle.cuh"

// ONE 512-thread block per SM (16 warps, 128 registers): measured 222.5 ms against 245.9 ms for four 128-thread blocks on
// config 2 (profiles/r2_tail.md).  Per-tracer fetch/retire times showed why: with four blocks per SM the warps of one SM
// advanced at 3.1 ... 13 us per step for the whole kernel (the hardware scheduler does not share issue slots evenly between
// warps of different blocks), and the orbits that sat in a starved warp were a 36 ms tail on < 500 lanes.
#ifndef RAPT_RKN_MINB
#define RAPT_RKN_MINB 1
#endif
#ifndef RAPT_RKN_THREADS
#define RAPT_RKN_THREADS 512
#endif
#ifndef RAPT_RKN_LOCKSTEP
#define RAPT_RKN_LOCKSTEP 0  /* k > 0: the warps of a block meet at a barrier every k iterations (see the end of the loop) */
#endif
#ifndef RAPT_RKN_CTRL
#define RAPT_RKN_CTRL 1      /* 1: the accepted-step controller multiplies by 1/fac (no fp64 division on the ~3-lane path) */
#endif
#ifndef RAPT_RKN_HK
#define RAPT_RKN_HK 1        /* 1: stage vectors scaled by h (see the step body); 0: the unscaled form measured in round 1 */
#endif
namespace RAPT_NS {

// K = dp/dt = q (E + P x B / (gamma m)), Particle.py:295
template <class F>
RAPT_DEV void lorentz_K(const FieldP &f, double q, double qg, double t, const double (&X)[3], const double (&P)[3],
                        double (&K)[3])
{
    double bx, by, bz;
    F::Bs(f, qg, t, X[0], X[1], X[2], bx, by, bz);                 // q/(gamma m) * B
    const double cx = P[1] * bz - P[2] * by, cy = P[2] * bx - P[0] * bz, cz = P[0] * by - P[1] * bx;
    if (F::HAS_E) {
        double ex, ey, ez; F::E(f, t, X[0], X[1], X[2], ex, ey, ez);
        K[0] = fma(q, ex, cx); K[1] = fma(q, ey, cy); K[2] = fma(q, ez, cz);
    } else {
        K[0] = cx; K[1] = cy; K[2] = cz;
    }
}

template <class F>
#ifdef RAPT_RKN_MAXNREG
__global__ void __maxnreg__(RAPT_RKN_MAXNREG) k_particle_rkn(const AdvArgs a)
#else
__global__ void __launch_bounds__(RAPT_RKN_THREADS, RAPT_RKN_MINB) k_particle_rkn(const AdvArgs a)
#endif

{
    if (F::CAN_FAIL) grid_cache_reset();     // gridded field: per-thread cell cache (rapt_fields.cuh)
    const double rtol = a.p.rtol, atol = a.p.atol;
    const double beta = 0.1, safe = 0.9, fac1 = 0.3, fac2 = 6.0, uround = 2.3e-16;
    const double expo1 = 1.0 / 8.0 - beta * 0.2, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    (void)facc2;
    const double lf0 = beta * -9.210340371976182;            // beta * log(1e-4): log of facold^beta at the start of a call

    double x[3], p[3], K1[3], X[3], P[3];
    double t = 0, h = 0, xend = 0, tstop = 0, tlim = 0, dt = 0, lfacold = 0, hmax = 0;
    double igm = 1, qg = 0, q = 0;
    int pid = -1;
    int nstep = 0, naccpt = 0, nrejct = 0, ncalls = 0, nstep_row = 0, naccpt_row = 0;
    int rowidx = 0, nst = 0, st = ST_OK;
    bool last = false, reject = false, need_row = false, have = false;
    double *myrows = nullptr;
    bool done = false;
    (void)done;
#if RAPT_RKN_LOCKSTEP
    unsigned iter = 0;
#endif
#ifdef RAPT_RKN_TRACE_TIMES
    double t_fetch = 0;
#endif

    for (;;) {
        // ---- (A) one step attempt: failure checks, clip the step to the row end.  The loop is ordered
        // step -> write back / fetch -> HINIT so that the divergent parts (B), (C) sit at the END of an
        // iteration: the warp reconverges at the loop's back edge and runs the step with all lanes.
        // (With HINIT in front of the step the lanes that skipped it ran ahead: 19 of 32 lanes per issue.)
        if (have && !need_row) {
        bool rowdone = false;
        if (nstep_row > 500) st = ST_NMAX;
        else if (0.1 * fabs(h) <= fabs(t) * uround) st = ST_HSMALL;
        // solver failure: r.integrate() hands back the state it reached and the reference appends it as a row labelled
        // with the row's end time before `while r.successful()` ends the loop (Particle.py:304-307)
        if (st != ST_OK) rowdone = true;
        else {
        if ((t + 1.01 * h - xend) > 0.0) { h = xend - t; last = true; }
        nstep_row++; nstep++;
        const double hg = h * igm;
        double K2[3], K3[3], K4[3], K5[3], K6[3], K7[3], K8[3], K9[3], K10[3], K11[3], K12[3];
#if RAPT_RKN_HK
        // Stage vectors scaled by the step: K_l holds h K_l.  The field comes back already multiplied by h q/(gamma m)
        // (one multiply per step instead of one per sum), every stage momentum is an FMA chain onto p and the |h| of
        // the error norm is absorbed: ~9 % fewer FP64 instructions per step than the unscaled form below.
        // K1 itself is scaled in place for the duration of the attempt (no second copy in registers): the FSAL
        // evaluation replaces it when the step is accepted, a rejected step divides the h out again.
        const double hqg = h * qg, hq = h * q;
#pragma unroll
        for (int i = 0; i < 3; i++) K1[i] *= h;
#define RKN_STAGE(PEXPR, XEXPR, CS, KOUT)                                                          \
        {                                                                                          \
            _Pragma("unroll") for (int i = 0; i < 3; i++) { P[i] = (PEXPR); X[i] = fma(hg, (XEXPR), x[i]); } \
            lorentz_K<F>(a.f, hq, hqg, t + (CS) * h, X, P, KOUT);                                   \
        }
        RKN_STAGE(fma(T8(A2_1), K1[i], p[i]),
                  T8(A2_1) * p[i], T8(C2), K2)
        RKN_STAGE(fma(T8(A3_2), K2[i], fma(T8(A3_1), K1[i], p[i])),
                  fma(TN(AA3_1), K1[i], TN(RS3) * p[i]), T8(C3), K3)
        RKN_STAGE(fma(T8(A4_3), K3[i], fma(T8(A4_1), K1[i], p[i])),
                  fma(TN(AA4_2), K2[i], fma(TN(AA4_1), K1[i], TN(RS4) * p[i])), T8(C4), K4)
        RKN_STAGE(fma(T8(A5_4), K4[i], fma(T8(A5_3), K3[i], fma(T8(A5_1), K1[i], p[i]))),
                  fma(TN(AA5_3), K3[i], fma(TN(AA5_2), K2[i], fma(TN(AA5_1), K1[i], TN(RS5) * p[i]))), T8(C5), K5)
        RKN_STAGE(fma(T8(A6_5), K5[i], fma(T8(A6_4), K4[i], fma(T8(A6_1), K1[i], p[i]))),
                  fma(TN(AA6_4), K4[i], fma(TN(AA6_3), K3[i], fma(TN(AA6_1), K1[i], TN(RS6) * p[i]))), T8(C6), K6)
        RKN_STAGE(fma(T8(A7_6), K6[i], fma(T8(A7_5), K5[i], fma(T8(A7_4), K4[i], fma(T8(A7_1), K1[i], p[i])))),
                  fma(TN(AA7_5), K5[i], fma(TN(AA7_4), K4[i], fma(TN(AA7_3), K3[i], fma(TN(AA7_1), K1[i], TN(RS7) * p[i])))), T8(C7), K7)
        RKN_STAGE(fma(T8(A8_7), K7[i], fma(T8(A8_6), K6[i], fma(T8(A8_5), K5[i], fma(T8(A8_4), K4[i], fma(T8(A8_1), K1[i], p[i]))))),
                  fma(TN(AA8_6), K6[i], fma(TN(AA8_5), K5[i], fma(TN(AA8_4), K4[i], fma(TN(AA8_3), K3[i], fma(TN(AA8_1), K1[i], TN(RS8) * p[i]))))), T8(C8), K8)
        RKN_STAGE(fma(T8(A9_8), K8[i], fma(T8(A9_7), K7[i], fma(T8(A9_6), K6[i], fma(T8(A9_5), K5[i], fma(T8(A9_4), K4[i], fma(T8(A9_1), K1[i], p[i])))))),
                  fma(TN(AA9_7), K7[i], fma(TN(AA9_6), K6[i], fma(TN(AA9_5), K5[i], fma(TN(AA9_4), K4[i], fma(TN(AA9_3), K3[i], fma(TN(AA9_1), K1[i], TN(RS9) * p[i])))))), T8(C9), K9)
        RKN_STAGE(fma(T8(A10_9), K9[i], fma(T8(A10_8), K8[i], fma(T8(A10_7), K7[i], fma(T8(A10_6), K6[i], fma(T8(A10_5), K5[i], fma(T8(A10_4), K4[i], fma(T8(A10_1), K1[i], p[i]))))))),
                  fma(TN(AA10_8), K8[i], fma(TN(AA10_7), K7[i], fma(TN(AA10_6), K6[i], fma(TN(AA10_5), K5[i], fma(TN(AA10_4), K4[i], fma(TN(AA10_3), K3[i], fma(TN(AA10_1), K1[i], TN(RS10) * p[i]))))))), T8(C10), K10)
        RKN_STAGE(fma(T8(A11_10), K10[i], fma(T8(A11_9), K9[i], fma(T8(A11_8), K8[i], fma(T8(A11_7), K7[i], fma(T8(A11_6), K6[i], fma(T8(A11_5), K5[i], fma(T8(A11_4), K4[i], fma(T8(A11_1), K1[i], p[i])))))))),
                  fma(TN(AA11_9), K9[i], fma(TN(AA11_8), K8[i], fma(TN(AA11_7), K7[i], fma(TN(AA11_6), K6[i], fma(TN(AA11_5), K5[i], fma(TN(AA11_4), K4[i], fma(TN(AA11_3), K3[i], fma(TN(AA11_1), K1[i], TN(RS11) * p[i])))))))), T8(C11), K11)
        RKN_STAGE(fma(T8(A12_11), K11[i], fma(T8(A12_10), K10[i], fma(T8(A12_9), K9[i], fma(T8(A12_8), K8[i], fma(T8(A12_7), K7[i], fma(T8(A12_6), K6[i], fma(T8(A12_5), K5[i], fma(T8(A12_4), K4[i], fma(T8(A12_1), K1[i], p[i]))))))))),
                  fma(TN(AA12_10), K10[i], fma(TN(AA12_9), K9[i], fma(TN(AA12_8), K8[i], fma(TN(AA12_7), K7[i], fma(TN(AA12_6), K6[i], fma(TN(AA12_5), K5[i], fma(TN(AA12_4), K4[i], fma(TN(AA12_3), K3[i], fma(TN(AA12_1), K1[i], TN(RS12) * p[i]))))))))), 1.0, K12)
#undef RKN_STAGE
        // new state (b-weights) in X, P and the two error estimators; everything carries one factor h, which is the |h|
        // of err = |h| err5 / sqrt(6 (err5 + 0.01 err3))
        double err = 0, err2 = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double sb = fma(T8(B12), K12[i], fma(T8(B11), K11[i], fma(T8(B10), K10[i], fma(T8(B9), K9[i], fma(T8(B8), K8[i], fma(T8(B7), K7[i], fma(T8(B6), K6[i], T8(B1) * K1[i])))))));
            P[i] = p[i] + sb;
            X[i] = fma(hg, fma(TN(BA11), K11[i], fma(TN(BA10), K10[i], fma(TN(BA9), K9[i], fma(TN(BA8), K8[i], fma(TN(BA7), K7[i], fma(TN(BA6), K6[i], fma(TN(BA1), K1[i], TN(SB) * p[i]))))))), x[i]);
            const double e5p = fma(T8(ER12), K12[i], fma(T8(ER11), K11[i], fma(T8(ER10), K10[i], fma(T8(ER9), K9[i], fma(T8(ER8), K8[i], fma(T8(ER7), K7[i], fma(T8(ER6), K6[i], T8(ER1) * K1[i])))))));
            const double e3p = fma(-T8(BHH3), K12[i], fma(-T8(BHH2), K9[i], fma(-T8(BHH1), K1[i], sb)));
            const double e5x = hg * (fma(TN(ERA11), K11[i], fma(TN(ERA10), K10[i], fma(TN(ERA9), K9[i], fma(TN(ERA8), K8[i], fma(TN(ERA7), K7[i], fma(TN(ERA6), K6[i], fma(TN(ERA5), K5[i], fma(TN(ERA4), K4[i], fma(TN(ERA1), K1[i], TN(SER) * p[i]))))))))));
            const double e3x = hg * (fma(TN(WA11), K11[i], fma(TN(WA10), K10[i], fma(TN(WA9), K9[i], fma(TN(WA8), K8[i], fma(TN(WA7), K7[i], fma(TN(WA6), K6[i], fma(TN(WA5), K5[i], fma(TN(WA4), K4[i], fma(TN(WA1), K1[i], TN(SW) * p[i]))))))))));
            const double iskx = fast_rcp1(atol + rtol * fmax(fabs(x[i]), fabs(X[i])));
            const double iskp = fast_rcp1(atol + rtol * fmax(fabs(p[i]), fabs(P[i])));
            const double a3 = e3x * iskx, b3 = e3p * iskp, a5 = e5x * iskx, b5 = e5p * iskp;
            err2 = fma(b3, b3, fma(a3, a3, err2)); err = fma(b5, b5, fma(a5, a5, err));
        }
        const double hnorm = 1.0;
#else
#define RKN_STAGE(PSUM, XSUM, CS, KOUT)                                                            \
        {                                                                                          \
            _Pragma("unroll") for (int i = 0; i < 3; i++) { P[i] = p[i] + h * (PSUM); X[i] = x[i] + hg * (XSUM); } \
            lorentz_K<F>(a.f, q, qg, t + (CS) * h, X, P, KOUT);                                     \
        }
        RKN_STAGE(T8(A2_1) * K1[i], T8(A2_1) * p[i], T8(C2), K2)
        RKN_STAGE(T8(A3_1) * K1[i] + T8(A3_2) * K2[i],
                  TN(RS3) * p[i] + h * (TN(AA3_1) * K1[i]), T8(C3), K3)
        RKN_STAGE(T8(A4_1) * K1[i] + T8(A4_3) * K3[i],
                  TN(RS4) * p[i] + h * (TN(AA4_1) * K1[i] + TN(AA4_2) * K2[i]), T8(C4), K4)
        RKN_STAGE(T8(A5_1) * K1[i] + T8(A5_3) * K3[i] + T8(A5_4) * K4[i],
                  TN(RS5) * p[i] + h * (TN(AA5_1) * K1[i] + TN(AA5_2) * K2[i] + TN(AA5_3) * K3[i]), T8(C5), K5)
        RKN_STAGE(T8(A6_1) * K1[i] + T8(A6_4) * K4[i] + T8(A6_5) * K5[i],
                  TN(RS6) * p[i] + h * (TN(AA6_1) * K1[i] + TN(AA6_3) * K3[i] + TN(AA6_4) * K4[i]), T8(C6), K6)
        RKN_STAGE(T8(A7_1) * K1[i] + T8(A7_4) * K4[i] + T8(A7_5) * K5[i] + T8(A7_6) * K6[i],
                  TN(RS7) * p[i] + h * (TN(AA7_1) * K1[i] + TN(AA7_3) * K3[i] + TN(AA7_4) * K4[i] + TN(AA7_5) * K5[i]), T8(C7), K7)
        RKN_STAGE(T8(A8_1) * K1[i] + T8(A8_4) * K4[i] + T8(A8_5) * K5[i] + T8(A8_6) * K6[i] + T8(A8_7) * K7[i],
                  TN(RS8) * p[i] + h * (TN(AA8_1) * K1[i] + TN(AA8_3) * K3[i] + TN(AA8_4) * K4[i] + TN(AA8_5) * K5[i] + TN(AA8_6) * K6[i]), T8(C8), K8)
        RKN_STAGE(T8(A9_1) * K1[i] + T8(A9_4) * K4[i] + T8(A9_5) * K5[i] + T8(A9_6) * K6[i] + T8(A9_7) * K7[i] + T8(A9_8) * K8[i],
                  TN(RS9) * p[i] + h * (TN(AA9_1) * K1[i] + TN(AA9_3) * K3[i] + TN(AA9_4) * K4[i] + TN(AA9_5) * K5[i] + TN(AA9_6) * K6[i] + TN(AA9_7) * K7[i]), T8(C9), K9)
        RKN_STAGE(T8(A10_1) * K1[i] + T8(A10_4) * K4[i] + T8(A10_5) * K5[i] + T8(A10_6) * K6[i] + T8(A10_7) * K7[i] + T8(A10_8) * K8[i] + T8(A10_9) * K9[i],
                  TN(RS10) * p[i] + h * (TN(AA10_1) * K1[i] + TN(AA10_3) * K3[i] + TN(AA10_4) * K4[i] + TN(AA10_5) * K5[i] + TN(AA10_6) * K6[i] + TN(AA10_7) * K7[i] + TN(AA10_8) * K8[i]), T8(C10), K10)
        RKN_STAGE(T8(A11_1) * K1[i] + T8(A11_4) * K4[i] + T8(A11_5) * K5[i] + T8(A11_6) * K6[i] + T8(A11_7) * K7[i] + T8(A11_8) * K8[i] + T8(A11_9) * K9[i] + T8(A11_10) * K10[i],
                  TN(RS11) * p[i] + h * (TN(AA11_1) * K1[i] + TN(AA11_3) * K3[i] + TN(AA11_4) * K4[i] + TN(AA11_5) * K5[i] + TN(AA11_6) * K6[i] + TN(AA11_7) * K7[i] + TN(AA11_8) * K8[i] + TN(AA11_9) * K9[i]), T8(C11), K11)
        RKN_STAGE(T8(A12_1) * K1[i] + T8(A12_4) * K4[i] + T8(A12_5) * K5[i] + T8(A12_6) * K6[i] + T8(A12_7) * K7[i] + T8(A12_8) * K8[i] + T8(A12_9) * K9[i] + T8(A12_10) * K10[i] + T8(A12_11) * K11[i],
                  TN(RS12) * p[i] + h * (TN(AA12_1) * K1[i] + TN(AA12_3) * K3[i] + TN(AA12_4) * K4[i] + TN(AA12_5) * K5[i] + TN(AA12_6) * K6[i] + TN(AA12_7) * K7[i] + TN(AA12_8) * K8[i] + TN(AA12_9) * K9[i] + TN(AA12_10) * K10[i]), 1.0, K12)
#undef RKN_STAGE
        // new state (b-weights) in X, P and the two error estimators
        double err = 0, err2 = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double sb = T8(B1) * K1[i] + T8(B6) * K6[i] + T8(B7) * K7[i] + T8(B8) * K8[i] + T8(B9) * K9[i] + T8(B10) * K10[i] + T8(B11) * K11[i] + T8(B12) * K12[i];
            P[i] = p[i] + h * sb;
            X[i] = x[i] + hg * (TN(SB) * p[i] + h * (TN(BA1) * K1[i] + TN(BA6) * K6[i] + TN(BA7) * K7[i] + TN(BA8) * K8[i] + TN(BA9) * K9[i] + TN(BA10) * K10[i] + TN(BA11) * K11[i]));
            const double e5p = T8(ER1) * K1[i] + T8(ER6) * K6[i] + T8(ER7) * K7[i] + T8(ER8) * K8[i] + T8(ER9) * K9[i] + T8(ER10) * K10[i] + T8(ER11) * K11[i] + T8(ER12) * K12[i];
            const double e3p = sb - T8(BHH1) * K1[i] - T8(BHH2) * K9[i] - T8(BHH3) * K12[i];
            const double e5x = igm * (TN(SER) * p[i] + h * (TN(ERA1) * K1[i] + TN(ERA4) * K4[i] + TN(ERA5) * K5[i] + TN(ERA6) * K6[i] + TN(ERA7) * K7[i] + TN(ERA8) * K8[i] + TN(ERA9) * K9[i] + TN(ERA10) * K10[i] + TN(ERA11) * K11[i]));
            const double e3x = igm * (TN(SW) * p[i] + h * (TN(WA1) * K1[i] + TN(WA4) * K4[i] + TN(WA5) * K5[i] + TN(WA6) * K6[i] + TN(WA7) * K7[i] + TN(WA8) * K8[i] + TN(WA9) * K9[i] + TN(WA10) * K10[i] + TN(WA11) * K11[i]));
            const double iskx = fast_rcp1(atol + rtol * fmax(fabs(x[i]), fabs(X[i])));
            const double iskp = fast_rcp1(atol + rtol * fmax(fabs(p[i]), fabs(P[i])));
            const double a3 = e3x * iskx, b3 = e3p * iskp, a5 = e5x * iskx, b5 = e5p * iskp;
            err2 += a3 * a3 + b3 * b3; err += a5 * a5 + b5 * b5;
        }
        const double hnorm = fabs(h);
#endif
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = hnorm * err * fast_rsqrt(6 * deno);
        if (err <= 1.0) {
            // accepted.  The controller's new step is only consumed when the row continues.
            if (!last) {
                // fac11 / facold^beta = exp(expo1 log(err) - beta log(facold)): one log and one exp, both
                // branch-free polynomial forms (~3e-14): this path runs at ~3 lanes in 9 of 10 iterations,
                // so its length matters
                const double lg = fast_log(fmax(err, 1e-300));
#if RAPT_RKN_CTRL
                // 1/fac directly: min(fac2, max(fac1, safe * facold^beta / err^expo1)) -- no division on this ~3-lane path
                const double rfac = fmin(fac2, fmax(fac1, safe * fast_exp(fma(-expo1, lg, lfacold))));
                double hnew = h * rfac;
#else
                double fac = fast_exp(fma(expo1, lg, -lfacold));
                fac = fmax(facc2, fmin(facc1, fac / safe));
                double hnew = h / fac;
#endif
                if (fabs(hnew) > hmax) hnew = hmax;
                if (reject) hnew = fmin(fabs(hnew), fabs(h));
                lfacold = beta * fmax(lg, -9.210340371976182);       // facold = max(err, 1e-4)
                t = t + h;
                h = hnew;
                reject = false;
            } else t = t + h;
            naccpt++; naccpt_row++;
            lorentz_K<F>(a.f, q, qg, t, X, P, K1);           // FSAL: k1 = f(t+h, ynew); the step becomes the state
#pragma unroll
            for (int i = 0; i < 3; i++) { x[i] = X[i]; p[i] = P[i]; }
            rowdone = last;
        } else {
#if RAPT_RKN_HK
            { const double ih = fast_rcp(h); _Pragma("unroll") for (int i = 0; i < 3; i++) K1[i] *= ih; }   // back to f(t, y)
#endif
            if (a.p.dop853_reject_rule == 1) h = h / fmin(facc1, RAPT_POW(err, expo1) / safe);
            else h = h / facc1;                              // scipy 1.18.1: 0.3 h whatever err is
            reject = true;
            if (naccpt_row >= 1) nrejct++;
            last = false;
            if (F::CAN_FAIL && !(err == err)) { st = RAPT_ST_FIELD; need_row = true; }   // left the grid: keep the last row
        }
        }   // st == ST_OK
        if (rowdone) {
            // ---- output row complete (Particle.py:305-309)
            rowidx++;
            if (myrows && a.store_every > 0 && (rowidx % a.store_every) == 0 && nst < a.max_rows) {
                double2 *r = reinterpret_cast<double2 *>(myrows + (size_t)nst * 8);
                double tag = a.segtag ? (double)a.segtag[pid] : (double)nstep;
                r[0] = make_double2(xend, x[0]); r[1] = make_double2(x[1], x[2]);
                r[2] = make_double2(p[0], p[1]); r[3] = make_double2(p[2], tag);
                nst++;
            }
            if (a.p.check_adiabaticity) {
                const double yy[6] = {x[0], x[1], x[2], p[0], p[1], p[2]};
                if (particle_isadiabatic<F>(a.f, a.p, xend, yy, a.mass[pid], a.charge[pid])) st = ST_ADIABATIC;
            }
            need_row = true;
        }
        }   // have && !need_row
        // ---- (B) particle finished?  write it back and fetch the next one
        if (have && need_row && !(st == ST_OK && t < tlim)) {
            if (a.seg_tstop) {                               // sliced adaptive epoch: keep the call's state
                a.seg_x[pid] = t; a.seg_dt[pid] = dt; a.seg_row[pid] = rowidx;
                if (st == ST_OK && t < tstop) st = RAPT_ST_SLICE;
            }
            a.t[pid] = xend; a.s1[pid] = x[0]; a.s2[pid] = x[1]; a.s3[pid] = x[2];
            a.s4[pid] = p[0]; a.s5[pid] = p[1]; a.s6[pid] = p[2];
            int *c = a.counters + 4 * (long long)pid;
            int nf = 2 * ncalls + 11 * nstep + naccpt;       // as scipy counts: SURVEY.md §3.1
            if (a.append) { c[0] += nf; c[1] += nstep; c[2] += naccpt; c[3] += nrejct; }
            else { c[0] = nf; c[1] = nstep; c[2] = naccpt; c[3] = nrejct; }
            a.status[pid] = st;
            a.tcur[pid] = t + dt;                            // Particle.py:306
            if (a.nrows) a.nrows[pid] = rowidx + 1;
            a.nstored[pid] = nst;
#ifdef RAPT_RKN_TRACE_TIMES      /* profiling build only: when was this tracer fetched and retired (global ns timer) */
            { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); a.tcur[pid] = (double)g_; a.dt_out[pid] = t_fetch; }
#endif
            have = false;
        }
        if (!have && !done) {
            int w = atomicAdd(a.queue, 1);
#if RAPT_RKN_LOCKSTEP
            if (w >= a.nwork) { done = true; w = 0; }
            else {
#else
            if (w >= a.nwork) break;
            {
#endif
            if (w < a.spread_first_wave) w = (w & 31) * (a.spread_first_wave >> 5) + (w >> 5);   // rapt_types.h: AdvArgs
            pid = a.order ? a.order[w] : w;
#ifdef RAPT_RKN_TRACE_TIMES
            { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); t_fetch = (double)g_; }
#endif
            t = a.t[pid];
            x[0] = a.s1[pid]; x[1] = a.s2[pid]; x[2] = a.s3[pid];
            p[0] = a.s4[pid]; p[1] = a.s5[pid]; p[2] = a.s6[pid];
            const double mass = a.mass[pid];
            q = a.charge[pid];
            double delta = a.delta_arr ? a.delta_arr[pid] : a.delta;
            tstop = t + delta;                               // Particle.py:304
            tlim = tstop;
            xend = t;                                        // row label of a tracer that takes no step
            double dt_fixed = 0;
            int row0 = 0;
            if (a.seg_tstop) {                               // resume a sliced call
                tstop = a.seg_tstop[pid]; tlim = fmin(tstop, a.slice_end);
                t = a.seg_x[pid]; dt_fixed = a.seg_dt[pid]; row0 = a.seg_row[pid];
            }
            // Particle.py:274-275, 282: gm (frozen: static field), vel, dt = cyclotron_period / cyclotronresolution
            double gm = sqrt(mass * mass + dot3(p[0], p[1], p[2], p[0], p[1], p[2]) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
            igm = 1.0 / gm; qg = q * igm;
            {
                double vx = p[0] / gm, vy = p[1] / gm, vz = p[2] / gm;
                double gamma = 1.0 / sqrt(1 - dot3(vx, vy, vz, vx, vy, vz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
                double Bm = F::magB(a.f, t, x[0], x[1], x[2]);
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
                    r[0] = make_double2(t, x[0]); r[1] = make_double2(x[1], x[2]);
                    r[2] = make_double2(p[0], p[1]); r[3] = make_double2(p[2], 0.0);
                    nst = 1;
                }
            }
            have = true; need_row = true;
            // degenerate output step (B = 0 or inf, bad resolution): the reference would never return.
            // delta <= 0 (or beyond this slice): nothing to do
            if (!(dt > 0.0) || dt > 1e300) st = RAPT_ST_HSMALL;
            else if (t < tlim) lorentz_K<F>(a.f, q, qg, t, x, p, K1);          // k1 = f(t, y)
            }
        }
        // ---- (C) a new output row is a new solver call: xend, hmax, HINIT
        if (have && need_row && st == ST_OK && t < tlim) {
            xend = t + dt;                                   // Particle.py:305 (also the row's time label)
            hmax = fabs(xend - t);
            double iskx[3], iskp[3];
            double dny = 0, dnf = 0, s1 = 0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                iskx[i] = fast_rcp1(atol + rtol * fabs(x[i])); iskp[i] = fast_rcp1(atol + rtol * fabs(p[i]));
                const double a_ = p[i] * igm * iskx[i], b_ = x[i] * iskx[i];
                const double c_ = K1[i] * iskp[i], d_ = p[i] * iskp[i], e_ = K1[i] * iskx[i];
                dnf += a_ * a_ + c_ * c_; dny += b_ * b_ + d_ * d_; s1 = fma(e_, e_, s1);
            }
            double h0 = 1e-6;
            if (!(dnf <= 1e-10 || dny <= 1e-10)) h0 = (dny * fast_rsqrt(dny * dnf)) * 0.01;      // 0.01 sqrt(dny / dnf)
            h0 = fmin(h0, hmax);
            // Euler probe f1 = f(t + h0, y + h0 f0); f1 - f0 = (h0 K1 / gm, K' - K1)
#pragma unroll
            for (int i = 0; i < 3; i++) { X[i] = x[i] + h0 * (p[i] * igm); P[i] = p[i] + h0 * K1[i]; }
            double Kp[3];
            lorentz_K<F>(a.f, q, qg, t + h0, X, P, Kp);
            double d2 = 0;
#pragma unroll
            for (int i = 0; i < 3; i++) { const double c_ = (Kp[i] - K1[i]) * iskp[i]; d2 = fma(c_, c_, d2); }
            const double hi = h0 * igm;
            d2 = fma(hi * hi, s1, d2);
            // gridded field: the probe point may lie outside the grid, where the reference's interpolator raises from
            // inside r.integrate() (no row for this call, the rows so far are kept)
            if (F::CAN_FAIL && !(d2 == d2)) st = RAPT_ST_FIELD;      // need_row stays set: retired at (B) of the next iteration
            else {
            // der2 = sqrt(d2)/h0, der12 = max(der2, sqrt(dnf)), h1 = (0.01/der12)^(1/8): compared in squares, the
            // root is only taken when h1 could be the minimum
            const double ih0 = fast_rcp(h0);
            const double d12 = fmax(d2 * ih0 * ih0, dnf);
            double h1;
            if (d12 <= 1e-30) h1 = fmax(1e-6, h0 * 1e-3);
            else {
                const double hm2 = hmax * hmax, hm4 = hm2 * hm2, hm8 = hm4 * hm4, hm16 = hm8 * hm8;
                if (hm16 > 1e-280 && d12 * hm16 * 1.000003 < 1e-4) h1 = hmax;      // h1 > hmax: not the minimum
                else h1 = fast_exp(0.125 * fma(-0.5, fast_log(d12), -4.605170185988091));   // (0.01 / sqrt(d12))^(1/8), branch-free
            }
            h = fmin(fmin(100 * h0, h1), hmax);
            lfacold = lf0; last = false; reject = false; nstep_row = 0; naccpt_row = 0;   // facold = 1e-4
            ncalls++;
            need_row = false;
            }
        }
#if RAPT_RKN_LOCKSTEP
        // The hardware warp scheduler is not fair: per-tracer fetch/retire times show warps of the same SM advancing at
        // 3.1 ... 13 us per step for the whole kernel (profiles/r2_tail.md), and the orbits that happen to sit in a starved
        // warp are the tail of the launch.  A barrier every k iterations makes the warps of a block advance together.
        if ((++iter % RAPT_RKN_LOCKSTEP) == 0 && __syncthreads_and(done)) break;
#endif
    }
}

}  // namespace RAPT_NS
