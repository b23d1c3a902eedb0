// TEST INFRASTRUCTURE ONLY.  The product's advance kernels (rapt_b200/csrc/rapt_particle.cuh, rapt_particle_rkn.cuh,
// rapt_gc.cuh: the FAST arithmetic flavour, the one bench.py times) compiled for the HOST through cuda_shim.h, behind
// entry points shaped like rapt_b200_particle_advance / rapt_b200_gc_advance (include/rapt_b200.h).  Every OpenMP thread
// runs the kernel body as one persistent lane pulling tracers from the shared work queue, exactly as a GPU thread does.
// tests/test_kernel_host.py compares the result with the oracle, so the kernels' step control, HINIT, FSAL reuse, row
// bookkeeping and Nystrom-form arithmetic are checked in the CPU-only test run as well.  Differences from the device
// build: MUFU seeds are 23-bit truncations of the exact value, FMA contraction is g++'s (-ffp-contract=fast -mfma).
// Nothing in rapt_b200/ loads this library and it is not part of librapt_b200.so; the product has no CPU path.
// Built twice (Makefile): libkernelhost.so = fast flavour (-ffp-contract=fast, as nvcc contracts that flavour);
// libkernelhost_strict.so = -DHC_STRICT, the strict flavour (-ffp-contract=off, as nvcc -fmad=false), whose operation
// order is the reference's: its trajectories are compared BIT FOR BIT with the goldens.
#include "cuda_shim.h"
#ifdef HC_STRICT
#define RAPT_STRICT 1
#else
#define RAPT_STRICT 0
#endif
#define RAPT_NS rapt_hostkernels
#include "../../include/rapt_b200.h"
#include "../../rapt_b200/csrc/rapt_particle.cuh"
#if !RAPT_STRICT
#include "../../rapt_b200/csrc/rapt_particle_rkn.cuh"
#endif
#include "../../rapt_b200/csrc/rapt_gc.cuh"
#include "../../rapt_b200/csrc/rapt_aux.cuh"
#include "../../rapt_b200/csrc/rapt_bc.cuh"
#include <omp.h>
#include <vector>

namespace rapt_hostkernels { double rapt_grid_cache[1]; }
using namespace rapt_hostkernels;

template <class Fn> static void run_lanes(int nthreads, Fn body)
{
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    body();
}

template <int KIND> static void go_particle(const rapt::AdvArgs &a, int rkn, int nthreads)
{
#if !RAPT_STRICT
    if (rkn) { run_lanes(nthreads, [&] { k_particle_rkn<Field<KIND>>(a); }); return; }
#endif
    run_lanes(nthreads, [&] { k_particle_dop853<Field<KIND>>(a); });
}
template <int KIND> static void go_gc(const rapt::AdvArgs &a, int nthreads)
{
    run_lanes(nthreads, [&] { k_gc_dopri5<Field<KIND>, 4>(a); });
}

static void fill_common(rapt::AdvArgs &a, const rapt_field_t *f, const rapt_params_t *p)
{
    static_assert(sizeof(rapt_field_t) == sizeof(rapt::FieldP), "FieldP layout");
    static_assert(sizeof(rapt_params_t) == sizeof(rapt::ParamsP), "ParamsP layout");
    memset(&a, 0, sizeof a);
    memcpy(&a.f, f, sizeof a.f);
    memcpy(&a.p, p, sizeof a.p);
}

extern "C" {

// rkn: 1 = k_particle_rkn (what the library launches for a static field without the equatorial constraint),
//      0 = k_particle_dop853 (the generic kernel)
int hc_particle_advance(const rapt_field_t *f, const rapt_params_t *p, long long n,
                        double *t, double *x, double *y, double *z, double *px, double *py, double *pz,
                        const double *mass, const double *charge, double delta,
                        long long store_every, long long max_rows, double *rows,
                        int *nrows, int *nstored, int *counters, int *status, double *tcur, double *dt_out,
                        int rkn, int nthreads)
{
    if (f->kind < 0 || f->kind > 5) return -1;
    if (rkn && (!f->is_static || p->enforce_equatorial || RAPT_STRICT)) return -2;
    rapt::AdvArgs a;
    fill_common(a, f, p);
    int queue = 0;
    a.nwork = n; a.order = nullptr; a.queue = &queue;
    a.t = t; a.s1 = x; a.s2 = y; a.s3 = z; a.s4 = px; a.s5 = py; a.s6 = pz;
    a.mass = mass; a.charge = charge; a.delta = delta;
    a.store_every = rows ? store_every : 0; a.max_rows = rows ? max_rows : 0; a.rows = rows;
    a.nstored = nstored; a.nrows = nrows; a.counters = counters; a.status = status; a.tcur = tcur; a.dt_out = dt_out;
    switch (f->kind) {
    case 0: go_particle<0>(a, rkn, nthreads); break;
    case 1: go_particle<1>(a, rkn, nthreads); break;
    case 2: go_particle<2>(a, rkn, nthreads); break;
    case 3: go_particle<3>(a, rkn, nthreads); break;
    case 4: go_particle<4>(a, rkn, nthreads); break;
    case 5: go_particle<5>(a, rkn, nthreads); break;
    }
    return 0;
}

int hc_gc_advance(const rapt_field_t *f, const rapt_params_t *p, int eom, long long n,
                  double *t, double *x, double *y, double *z, double *ppar,
                  const double *mu, const double *v, const double *mass, const double *charge,
                  const double *dt, double delta, long long store_every, long long max_rows, double *rows,
                  int *nrows, int *nstored, int *counters, int *status, double *tcur, int nthreads)
{
    if (f->kind < 0 || f->kind > 5) return -1;
    rapt::AdvArgs a;
    fill_common(a, f, p);
    int queue = 0;
    a.nwork = n; a.order = nullptr; a.queue = &queue;
    a.t = t; a.s1 = x; a.s2 = y; a.s3 = z; a.s4 = ppar;
    a.mass = mass; a.charge = charge; a.mu = mu; a.v = v; a.dtin = dt; a.delta = delta; a.eom = eom;
    a.store_every = rows ? store_every : 0; a.max_rows = rows ? max_rows : 0; a.rows = rows;
    a.nstored = nstored; a.nrows = nrows; a.counters = counters; a.status = status; a.tcur = tcur;
    switch (f->kind) {
    case 0: go_gc<0>(a, nthreads); break;
    case 1: go_gc<1>(a, nthreads); break;
    case 2: go_gc<2>(a, nthreads); break;
    case 3: go_gc<3>(a, nthreads); break;
    case 4: go_gc<4>(a, nthreads); break;
    case 5: go_gc<5>(a, nthreads); break;
    }
    return 0;
}

}  // extern "C"

// AdaptiveEnsemble: the epoch loop of rapt_b200_adaptive_advance (capi.cu) over the host builds of the same three
// kernels: [particle kernel over the particle-mode list, guiding-centre kernel over the GC-mode list] -> per-tracer
// switch (adaptive_switch_one, the body of k_adaptive_switch) -> regroup by mode.  One slice (the library's default).
template <int KIND> static int go_adaptive(const rapt_field_t *f, const rapt_params_t *p, long long n,
                                           const double *x, const double *y, const double *z,
                                           const double *vx, const double *vy, const double *vz,
                                           const double *t0, const double *mass, const double *charge,
                                           double gc_dt, double delta, long long store_every, long long max_rows, double *rows,
                                           int *nstored, int *nseg, int *mode_out, double *fin, int *counters, int *status,
                                           int *epochs_out, int nthreads)
{
    const bool want_rows = rows && max_rows > 0;
    std::vector<double> ps[7], gs[7], tvar(n), rem(n), tcur(n), dtg(n, gc_dt), dtp(n), sts(n), sx(n), sdt(n);
    for (int k = 0; k < 7; k++) { ps[k].assign(n, 0.0); gs[k].assign(n, 0.0); }
    std::vector<int> mode(n), segtag(n), srow(n), listP(n), listG(n);
    int counts[2] = {0, 0};
    for (long long i = 0; i < 4 * n; i++) counters[i] = 0;
    rapt::AdaptArgs sw;
    memset(&sw, 0, sizeof sw);
    memcpy(&sw.f, f, sizeof sw.f); memcpy(&sw.p, p, sizeof sw.p);
    sw.n = n; sw.delta = delta;
    sw.x0 = x; sw.y0 = y; sw.z0 = z; sw.vx0 = vx; sw.vy0 = vy; sw.vz0 = vz; sw.t0 = t0; sw.mass = mass; sw.charge = charge;
    sw.pt = ps[0].data(); sw.px = ps[1].data(); sw.py = ps[2].data(); sw.pz = ps[3].data();
    sw.ppx = ps[4].data(); sw.ppy = ps[5].data(); sw.ppz = ps[6].data();
    sw.gt = gs[0].data(); sw.gx = gs[1].data(); sw.gy = gs[2].data(); sw.gz = gs[3].data();
    sw.gpp = gs[4].data(); sw.mu = gs[5].data(); sw.v = gs[6].data();
    sw.mode = mode.data(); sw.status = status; sw.nseg = nseg; sw.segtag = segtag.data(); sw.nstored = nstored;
    sw.tvar = tvar.data(); sw.rem = rem.data(); sw.tcur = tcur.data();
    sw.max_rows = want_rows ? max_rows : 0; sw.rows = want_rows ? rows : nullptr;
    sw.listP = listP.data(); sw.listG = listG.data(); sw.counts = counts;
    sw.seg_tstop = sts.data(); sw.seg_x = sx.data(); sw.seg_dt = sdt.data(); sw.seg_row = srow.data();
    auto do_switch = [&](int first) {
        sw.first = first;
        counts[0] = counts[1] = 0;
        for (long long i = 0; i < n; i++) {
            const int want = adaptive_switch_one<Field<KIND>>(sw, i);
            if (want == 0) listP[counts[0]++] = (int)i;
            if (want == 1) listG[counts[1]++] = (int)i;
        }
    };
    do_switch(1);
    rapt_params_t pc = *p;
    pc.check_adiabaticity = 1;
    double tmin = t0[0];
    for (long long i = 1; i < n; i++) tmin = std::min(tmin, t0[i]);
    const double slice_end = tmin + (delta > 0 ? delta : 1.0);
    int epochs = 0;
    for (;; epochs++) {
        if (counts[0] == 0 && counts[1] == 0) break;
        if (epochs > 100000) return -3;
        const int cnt[2] = {counts[0], counts[1]};
        const bool rkn = !RAPT_STRICT && f->is_static && !p->enforce_equatorial;
        if (cnt[0] > 0) {
            rapt::AdvArgs a;
            fill_common(a, f, &pc);
            int queue = 0;
            a.nwork = cnt[0]; a.order = sw.listP; a.queue = &queue;
            a.t = sw.pt; a.s1 = sw.px; a.s2 = sw.py; a.s3 = sw.pz; a.s4 = sw.ppx; a.s5 = sw.ppy; a.s6 = sw.ppz;
            a.mass = sw.mass; a.charge = sw.charge; a.delta = 0; a.delta_arr = sw.rem;
            a.store_every = want_rows ? std::max<long long>(store_every, 1) : 0; a.max_rows = sw.max_rows; a.rows = sw.rows;
            a.nstored = sw.nstored; a.nrows = nullptr; a.counters = counters; a.status = sw.status;
            a.tcur = sw.tcur; a.dt_out = dtp.data(); a.segtag = sw.segtag; a.append = 1;
            a.seg_tstop = sw.seg_tstop; a.seg_x = sw.seg_x; a.seg_dt = sw.seg_dt; a.seg_row = sw.seg_row; a.slice_end = slice_end;
            go_particle<KIND>(a, rkn, nthreads);
        }
        if (cnt[1] > 0) {
            rapt::AdvArgs a;
            fill_common(a, f, &pc);
            int queue = 0;
            a.nwork = cnt[1]; a.order = sw.listG; a.queue = &queue;
            a.t = sw.gt; a.s1 = sw.gx; a.s2 = sw.gy; a.s3 = sw.gz; a.s4 = sw.gpp;
            a.mass = sw.mass; a.charge = sw.charge; a.mu = sw.mu; a.v = sw.v; a.dtin = dtg.data();
            a.delta = 0; a.delta_arr = sw.rem; a.eom = RAPT_EOM_TAOCHANBRIZARD;
            a.store_every = want_rows ? std::max<long long>(store_every, 1) : 0; a.max_rows = sw.max_rows; a.rows = sw.rows;
            a.nstored = sw.nstored; a.nrows = nullptr; a.counters = counters; a.status = sw.status;
            a.tcur = sw.tcur; a.segtag = sw.segtag; a.append = 1;
            a.seg_tstop = sw.seg_tstop; a.seg_x = sw.seg_x; a.seg_dt = sw.seg_dt; a.seg_row = sw.seg_row; a.slice_end = slice_end;
            go_gc<KIND>(a, nthreads);
        }
        do_switch(0);
    }
    if (epochs_out) *epochs_out = epochs;
    for (long long i = 0; i < n; i++) {
        if (mode_out) mode_out[i] = mode[i];
        if (fin) {
            double *o = fin + 8 * i;
            if (mode[i] == 0) { for (int k = 0; k < 7; k++) o[k] = ps[k][i]; o[7] = 0; }
            else { for (int k = 0; k < 5; k++) o[k] = gs[k][i]; o[5] = gs[5][i]; o[6] = gs[6][i]; o[7] = 1; }
        }
    }
    return 0;
}

// k_field_ops (the _Field operator layer, fields.py:76-280), one "thread" per point
template <int KIND> static void go_field_ops(rapt::OpsArgs a)
{
    // the kernel reads blockIdx.x * blockDim.x + threadIdx.x: present one point per call through a one-element view
    const long long n = a.n;
    double *outs[12] = {a.B, a.E, a.unitb, a.magB, a.gradB, a.jac, a.curlb, a.curv, a.dBdt, a.dbdt, a.lscale, a.tscale};
    const int width[12] = {3, 3, 3, 1, 3, 9, 3, 1, 1, 3, 1, 1};
    for (long long i = 0; i < n; i++) {
        rapt::OpsArgs b = a;
        b.n = 1; b.tpos = a.tpos + 4 * i;
        double **dst[12] = {&b.B, &b.E, &b.unitb, &b.magB, &b.gradB, &b.jac, &b.curlb, &b.curv, &b.dBdt, &b.dbdt, &b.lscale, &b.tscale};
        for (int k = 0; k < 12; k++) *dst[k] = outs[k] ? outs[k] + (long long)width[k] * i : nullptr;
        k_field_ops<Field<KIND>>(b);
    }
}

// k_bounce_setup (GuidingCenter.bounceperiod: field-line RKF45 trace + half-bounce path), one "thread" per guiding centre
template <int KIND> static void go_bounce(const rapt::BounceArgs &a)
{
    for (long long i = 0; i < a.n; i++) {
        rapt::BounceArgs b = a;
        b.n = 1;
        b.t = a.t + i; b.x = a.x + i; b.y = a.y + i; b.z = a.z + i;
        if (a.ppar) b.ppar = a.ppar + i;
        if (a.mu) b.mu = a.mu + i;
        if (a.mass) b.mass = a.mass + i;
        b.Bm = a.Bm + i; b.v = a.v + i; b.ds = a.ds + i; b.npts = a.npts + i;
        b.curve = a.curve + i * a.max_pts * 5; b.scratch = a.scratch + i * a.max_pts * 4;
        if (a.period) b.period = a.period + i;
        k_bounce_setup<Field<KIND>>(b);
    }
}

extern "C" {

// k_bounce_center: op 0 = BounceCenter.advance, op 1 = (S_b, I, gradI, deriv) at given points.  One lane walks the
// tracers (the shim's grid is one block of one thread), scratch curves of max_pts points.
int hc_bounce_center(const rapt_field_t *f, int op, int quadrature, long long n,
                     double *t, double *x, double *y, double *z, const double *mu, const double *Bm,
                     const double *v, const double *mass, const double *charge, const double *dt_in,
                     double bctimestep, double delta, double rtol, double atol, double fieldlineresolution, double eyegradientstep,
                     long long store_every, long long max_rows, double *rows, int *nrows, int *nstored, int *counters,
                     int *status, double *dt_out, double *out, long long max_pts)
{
    std::vector<double> cv((size_t)max_pts * 5), bw((size_t)max_pts * 4);
    rapt::BCArgs a;
    memset(&a, 0, sizeof a);
    memcpy(&a.f, f, sizeof a.f);
    a.op = op; a.quadrature = quadrature; a.rtol = rtol; a.atol = atol; a.flres = fieldlineresolution;
    a.eyestep = eyegradientstep; a.bctimestep = bctimestep; a.delta = delta;
    a.n = n; a.max_pts = max_pts; a.store_every = rows ? store_every : 0; a.max_rows = rows ? max_rows : 0;
    a.t = t; a.x = x; a.y = y; a.z = z; a.mu = mu; a.Bm = Bm; a.v = v; a.mass = mass; a.charge = charge;
    a.dtin = dt_in; a.dt_out = dt_out; a.rows = rows;
    a.nrows = nrows; a.nstored = nstored; a.counters = counters; a.status = status; a.out = out;
    a.curve = cv.data(); a.scratch = bw.data();
    switch (f->kind) {
    case 0: k_bounce_center<Field<0>>(a); break;
    case 1: k_bounce_center<Field<1>>(a); break;
    default: return -1;
    }
    return 0;
}

// rapt_b200_bounce_setup / rapt_b200_bounce_period in one: curve and/or period may be NULL
int hc_bounce(const rapt_field_t *f, int quadrature, double fieldlineresolution, long long n,
              const double *t, const double *x, const double *y, const double *z, const double *ppar,
              const double *mu, const double *mass, double *Bm, double *v, double *ds, int *npts, long long max_pts,
              double *curve, double *period)
{
    std::vector<double> cv, scr((size_t)n * max_pts * 4);
    if (!curve) { cv.resize((size_t)n * max_pts * 5); curve = cv.data(); }
    rapt::BounceArgs a;
    memset(&a, 0, sizeof a);
    memcpy(&a.f, f, sizeof a.f);
    a.flres = fieldlineresolution; a.n = n; a.max_pts = max_pts;
    a.t = t; a.x = x; a.y = y; a.z = z; a.ppar = ppar; a.mu = mu; a.mass = mass;
    a.Bm = Bm; a.v = v; a.ds = ds; a.npts = npts; a.curve = curve; a.scratch = scr.data();
    a.period = period; a.quadrature = quadrature;
    switch (f->kind) {
    case 0: go_bounce<0>(a); break;
    case 1: go_bounce<1>(a); break;
    case 4: go_bounce<4>(a); break;
    case 5: go_bounce<5>(a); break;
    default: return -1;
    }
    return 0;
}

int hc_field_ops(const rapt_field_t *f, long long npt, const double *tpos,
                 double *B, double *E, double *unitb, double *magB, double *gradB, double *jacobianB,
                 double *curlb, double *curvature, double *dBdt, double *dbdt, double *lengthscale, double *timescale)
{
    rapt::OpsArgs a;
    memset(&a, 0, sizeof a);
    memcpy(&a.f, f, sizeof a.f);
    a.n = npt; a.tpos = tpos;
    a.B = B; a.E = E; a.unitb = unitb; a.magB = magB; a.gradB = gradB; a.jac = jacobianB; a.curlb = curlb; a.curv = curvature;
    a.dBdt = dBdt; a.dbdt = dbdt; a.lscale = lengthscale; a.tscale = timescale;
    switch (f->kind) {
    case 0: go_field_ops<0>(a); break;
    case 1: go_field_ops<1>(a); break;
    case 2: go_field_ops<2>(a); break;
    case 3: go_field_ops<3>(a); break;
    case 4: go_field_ops<4>(a); break;
    case 5: go_field_ops<5>(a); break;
    default: return -1;
    }
    return 0;
}

int hc_adaptive_advance(const rapt_field_t *f, const rapt_params_t *p, long long n,
                        const double *x, const double *y, const double *z, const double *vx, const double *vy, const double *vz,
                        const double *t0, const double *mass, const double *charge,
                        double gc_dt, double delta, long long store_every, long long max_rows, double *rows,
                        int *nstored, int *nseg, int *mode_out, double *fin, int *counters, int *status, int *epochs_out,
                        int nthreads)
{
#define HC_ADAPT(K) case K: return go_adaptive<K>(f, p, n, x, y, z, vx, vy, vz, t0, mass, charge, gc_dt, delta, store_every, \
                                                  max_rows, rows, nstored, nseg, mode_out, fin, counters, status, epochs_out, nthreads);
    switch (f->kind) { HC_ADAPT(0) HC_ADAPT(1) HC_ADAPT(2) HC_ADAPT(3) HC_ADAPT(4) HC_ADAPT(5) }
#undef HC_ADAPT
    return -1;
}

}  // extern "C"
