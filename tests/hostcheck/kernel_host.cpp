// TEST INFRASTRUCTURE ONLY.  The product's advance kernels (rapt_b200/csrc/rapt_particle.cuh, rapt_particle_rkn.cuh,
// rapt_gc.cuh: the FAST arithmetic flavour, the one bench.py times) compiled for the HOST through cuda_shim.h, behind
// entry points shaped like rapt_b200_particle_advance / rapt_b200_gc_advance (include/rapt_b200.h).  Every OpenMP thread
// runs the kernel body as one persistent lane pulling tracers from the shared work queue, exactly as a GPU thread does.
// tests/test_kernel_host.py compares the result with the oracle, so the kernels' step control, HINIT, FSAL reuse, row
// bookkeeping and Nystrom-form arithmetic are checked in the CPU-only test run as well.  Differences from the device
// build: MUFU seeds are 23-bit truncations of the exact value, FMA contraction is g++'s (-ffp-contract=fast -mfma).
// Nothing in rapt_b200/ loads this library and it is not part of librapt_b200.so; the product has no CPU path.
#include "cuda_shim.h"
#define RAPT_STRICT 0
#define RAPT_NS rapt_hostfast
#include "../../include/rapt_b200.h"
#include "../../rapt_b200/csrc/rapt_particle_rkn.cuh"
#include "../../rapt_b200/csrc/rapt_gc.cuh"
#include <omp.h>

namespace rapt_hostfast { double rapt_grid_cache[1]; }
using namespace rapt_hostfast;

template <class Fn> static void run_lanes(int nthreads, Fn body)
{
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    body();
}

template <int KIND> static void go_particle(const rapt::AdvArgs &a, int rkn, int nthreads)
{
    if (rkn) run_lanes(nthreads, [&] { k_particle_rkn<Field<KIND>>(a); });
    else run_lanes(nthreads, [&] { k_particle_dop853<Field<KIND>>(a); });
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
    if (rkn && (!f->is_static || p->enforce_equatorial)) return -2;
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
