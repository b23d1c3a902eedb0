// rapt_types.h -- POD argument blocks shared by the host C-ABI layer, both arithmetic flavours of
// the kernels and the NVRTC-compiled user-field kernels.  Layouts of FieldP / ParamsP mirror
// rapt_field_t / rapt_params_t in include/rapt_b200.h (static_asserts in capi.cu).
#pragma once

namespace rapt {

struct FieldP {
    int kind, is_static, user_id, nprm;
    double prm[16];
    double gradstep, tstep;
};

// fields.Grid (rapt/fields.py:513-814) on the device.  The host side of the C ABI resolves
// rapt_field_t.user_id (the handle of rapt_b200_grid_create) into the device address of this block
// (FieldP::prm[0]; prm[1] != 0: the grid has an electric-field table).
// Tables: B and E as [nt][nx][ny][nz] nodes of 4 doubles (x, y, z component + pad): the two z-neighbours
// of a cell edge are one aligned 64-byte segment.  E == nullptr: the electric field is identically zero.
struct GridP {
    const double *t, *x, *y, *z;       // node coordinates (t: nt >= 2 time points, unused when nt == 1)
    const double *B, *E;
    int nt, nx, ny, nz;
    double x0, xinv, y0, yinv, z0, zinv;   // uniform axis: first node and 1/spacing (inv == 0: not uniform)
};

struct ParamsP {
    double rtol, atol, cyclotronresolution, epss, epst;
    int enforce_equatorial, check_adiabaticity, dop853_reject_rule, arith, sort_by_work;
    int reserved[3];
};

// one advance launch (Particle or GuidingCenter)
struct AdvArgs {
    FieldP f;
    ParamsP p;
    long long nwork;              // number of work items
    const int *order;             // work item -> particle id (NULL: identity)
    int *queue;                   // global work counter (zeroed before launch)
    // state, structure of arrays (Particle: t,x,y,z,px,py,pz ; GuidingCenter: t,X,Y,Z,ppar)
    double *t, *s1, *s2, *s3, *s4, *s5, *s6;
    const double *mass, *charge;
    const double *mu, *v, *dtin;  // guiding centre only
    double delta;
    const double *delta_arr;      // per-particle duration (adaptive), or NULL
    long long store_every, max_rows;
    double *rows;                 // [n][max_rows][8] or NULL
    int *nstored, *nrows, *counters, *status;
    double *tcur, *dt_out;
    const int *segtag;            // adaptive: column-7 tag per particle, or NULL
    int append;                   // adaptive: rows appended after nstored[pid]; counters accumulated
    int eom;                      // guiding centre only
    // adaptive, time-sliced epochs: a segment (= one reference advance() call) may be interrupted at a row
    // boundary once the solver time passes slice_end and resumed by a later launch.  Everything the call
    // fixed at its start is kept here so that the resumed run is bit-identical to an uninterrupted one.
    double *seg_tstop;            // absolute end time of the call (t0 + delta), or NULL: not sliced
    double *seg_x;                // Particle: solver time (can differ from the row label by an ulp)
    double *seg_dt;               // Particle: output step chosen at the start of the call (0: not chosen yet)
    int *seg_row;                 // output-row index within the call (decimation phase)
    double slice_end;
    // work order of the first wave: with `order` sorted longest-first the 32 lanes of a warp would start on 32 NEIGHBOURS of
    // the sorted list, i.e. whole warps of the very longest tracers.  spread_first_wave = L (lanes of the launch, a multiple
    // of 32) hands lane j of the k-th warp the item j * (L / 32) + k instead: every warp starts with the same mix.
    int spread_first_wave;
};

// state columns handed to the final-diagnostics kernel (capi.cu:k_final_diagnostics)
struct DiagCols { const double *col[8]; };

// batched _Field operators
struct OpsArgs {
    FieldP f;
    long long n;
    const double *tpos;
    double *B, *E, *unitb, *magB, *gradB, *jac, *curlb, *curv, *dBdt, *dbdt, *lscale, *tscale;
};

// small per-call kernels: 0 gc_construct, 1 p2g, 2 g2p, 3 isadiabatic
struct MiscArgs {
    FieldP f;
    ParamsP p;
    long long n;
    int op;                   // 0 gc_construct, 1 p2g, 2 g2p, 3 isadiabatic
    int mode;
    long long stride;
    double t_eval;
    const double *a0, *a1, *a2, *a3, *a4, *a5, *a6;
    double *o0, *o1, *o2;
    int *io;
};

// bounce-period set-up
struct BounceArgs {
    FieldP f;
    double flres;
    long long n, max_pts;
    const double *t, *x, *y, *z, *ppar, *mu, *mass;   // mu == NULL: Bm[] is an INPUT (plain Fieldline trace)
    double *Bm, *v, *ds;
    int *npts;
    double *curve;            // [n][max_pts][5] : s, x, y, z, |B|
    double *scratch;          // [n][max_pts][4] : backward half before reversal; then spline work arrays
    double *period;           // optional: bounce period (flutils.bounceperiod) from the traced curve
    int quadrature;           // 0: closed form on the quadratic spline, 1: brentq + QUADPACK QAGS as the reference
};

// BounceCenter.advance (rapt/BounceCenter.py:206-251) and its pieces (flutils.halfbouncepath / eye / gradI)
struct BCArgs {
    FieldP f;
    int op;                   // 0: advance; 1: out[i] = (S_b, I, gradI[3], deriv[3]) at the given points; 2: out[i] = I
    int quadrature;           // halfbouncepath: 1 brentq + QAGS as the reference, 0 closed form
    double rtol, atol, flres, eyestep, bctimestep, delta;
    long long n, max_pts, store_every, max_rows;
    double *t, *x, *y, *z;    // last row, in/out (t = row label, BounceCenter.py:248-250)
    const double *mu, *Bm;    // mirror field from mu (BounceCenter.py:227) unless Bm is given
    const double *v, *mass, *charge;
    const double *dtin;       // output step per tracer, or NULL: BCtimestep * bounce period
    double *dt_out, *tsolver; // optional
    double *rows;             // [n][max_rows][4] or NULL
    int *nrows, *nstored, *counters, *status;
    double *out;              // op 1: [n][8]; op 2: [n]
    double *curve, *scratch;  // per LANE: [lanes][max_pts][5], [lanes][max_pts][4]
    const int *order;         // work item -> tracer (sorted by field-line length, so that a warp's lanes trace lines of similar length), or NULL
};

// Adaptive: per-tracer state of both modes + the epoch bookkeeping (rapt/Adaptive.py:70-104, 187-222)
struct AdaptArgs {
    FieldP f;
    ParamsP p;
    long long n;
    int first;                         // 1: Adaptive.__init__ (initial mode choice), 0: after an epoch
    double delta;
    const double *x0, *y0, *z0, *vx0, *vy0, *vz0, *t0;   // constructor arguments (first == 1)
    const double *mass, *charge;
    double *pt, *px, *py, *pz, *ppx, *ppy, *ppz;        // Particle-mode state (last row)
    double *gt, *gx, *gy, *gz, *gpp, *mu, *v;           // GuidingCenter-mode state (last row), mu, speed
    int *mode, *status, *nseg, *segtag, *nstored;
    double *tvar, *rem, *tcur;                          // Adaptive.advance's `t`, delta - t, current.tcur
    double *seg_tstop, *seg_x, *seg_dt;                 // per-call resume state (see AdvArgs)
    int *seg_row;
    long long max_rows;
    double *rows;
    int *listP, *listG, *counts;                        // compacted work lists by mode; counts[0..1]
};

}  // namespace rapt

#ifndef RAPT_ST_OK   /* same values as include/rapt_b200.h (not visible to NVRTC) */
#define RAPT_ST_OK 1
#define RAPT_ST_ADIABATIC 2
#define RAPT_ST_NONADIABATIC 3
#ifndef RAPT_ST_SLICE
#define RAPT_ST_SLICE 4          /* adaptive: segment interrupted at a row boundary, to be resumed */
#endif
#define RAPT_ST_NMAX (-2)
#define RAPT_ST_HSMALL (-3)
#define RAPT_ST_GCITER (-5)
#define RAPT_ST_FIELD (-6)         /* Grid field evaluated outside its bounds (the reference raises ValueError) */
#define RAPT_ST_ROWCAP (-10)
#endif
