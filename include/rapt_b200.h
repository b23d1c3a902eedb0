/* rapt_b200.h -- C ABI of librapt_b200.so: the B200 (sm_100a) engine behind RAPT's
 * Particle / GuidingCenter / Adaptive `.advance()` hot path.
 *
 * The reference (mkozturk/rapt) is pure Python and has no FFI of its own; its boundary for this
 * path is the Python object level.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).  INTEGRATION.md shows the
 * ctypes stubs a RAPT maintainer would add to call them.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every function returns 0 on success or a
 *     negative RAPT_E_* code (never throws); rapt_b200_last_error() gives the text.
 *   - the caller owns every buffer; the library never frees caller memory.
 *   - arrays are contiguous little-endian float64 / int32.  Ensemble state is structure-of-arrays.
 *   - `_dev` variants take DEVICE pointers and a cudaStream_t (as void*) and do no host<->device
 *     copies and no synchronisation; the plain variants take HOST pointers, copy in, run, copy out.
 *   - one host thread per call; `_dev` calls share library-owned scratch (work queue counters, the
 *     longest-first ordering buffers), so issue them on one stream or otherwise order them yourself.
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns RAPT_E_NODEVICE.
 *
 * Trajectory rows: 8 doubles (64 B) per stored row, particle-major:
 *     rows[(i*max_rows + r)*8 + c]
 *   Particle     : c = 0..6 -> t, x, y, z, px, py, pz           (Particle.trajectory, Particle.py:99-100)
 *   GuidingCenter: c = 0..4 -> t, X, Y, Z, p_par ; c = 5 -> mu   (GuidingCenter.trajectory, GuidingCenter.py:116)
 *   c = 7 : plain advance -> cumulative attempted RK steps (as double) when the row was written;
 *           adaptive advance -> segment tag = 2*segment_index + mode (mode 0 particle, 1 guiding centre).
 *   Row 0 of each particle is its state at entry.  Output row k (k >= 1) is stored iff
 *   k % store_every == 0 and fewer than max_rows rows are stored; store_every = 0 stores nothing.
 *   The state arrays always return the LAST row (what the reference would have appended last).
 */
#ifndef RAPT_B200_H
#define RAPT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes */
#define RAPT_OK            0
#define RAPT_E_NODEVICE   -1   /* no CUDA device / driver */
#define RAPT_E_CUDA       -2   /* a CUDA call failed (see rapt_b200_last_error) */
#define RAPT_E_ARG        -3   /* bad argument */
#define RAPT_E_NVRTC      -4   /* user field snippet failed to compile (log returned) */
#define RAPT_E_UNSUPPORTED -5

/* ---- built-in analytic field models: rapt/fields.py */
#define RAPT_FIELD_EARTHDIPOLE   0   /* fields.py:282-317  prm = {_coeff = -3*B0*Re^3}            */
#define RAPT_FIELD_DOUBLEDIPOLE  1   /* fields.py:319-362  prm = {_coeff = -B0*Re^3, _dd, _k}     */
#define RAPT_FIELD_UNIFORMBZ     2   /* fields.py:364-390  prm = {Bz}                             */
#define RAPT_FIELD_CROSSEDEB     3   /* fields.py:392-427  prm = {Bz, Ey}                         */
#define RAPT_FIELD_VARDIPOLE     4   /* fields.py:429-470  prm = {amp, period}                    */
#define RAPT_FIELD_PARABOLIC     5   /* fields.py:472-511  prm = {B0, Bn, d}                      */
#define RAPT_FIELD_GRID          6   /* fields.py:513-814  gridded E/B; user_id = handle of rapt_b200_grid_create, prm unused */
#define RAPT_FIELD_USER        100   /* NVRTC-compiled snippet (rapt_b200_field_nvrtc)            */

/* ---- guiding-centre equations of motion: GuidingCenter.py:329-395, selected by advance(eom=...) :449 */
#define RAPT_EOM_TAOCHANBRIZARD  0
#define RAPT_EOM_BRIZARDCHAN     1
#define RAPT_EOM_NORTHROPTELLER  2

/* ---- per-particle status (scipy `idid` analogue + mode-switch signalling) */
#define RAPT_ST_OK            1   /* advance() ran to t0+delta                                     */
#define RAPT_ST_ADIABATIC     2   /* Particle: `raise Adiabatic`   (Particle.py:308-309)           */
#define RAPT_ST_NONADIABATIC  3   /* GuidingCenter: `raise NonAdiabatic` (GuidingCenter.py:457-458)*/
#define RAPT_ST_SLICE         4   /* internal to adaptive epochs: interrupted at a row boundary, resumed later */
#define RAPT_ST_NMAX         -2   /* dop: more than nsteps=500 steps in one output interval        */
#define RAPT_ST_HSMALL       -3   /* dop: step size underflow                                      */
#define RAPT_ST_GCITER       -5   /* utils.guidingcenter did not converge (utils.py:326)           */
#define RAPT_ST_FIELD        -6   /* fields.Grid evaluated outside the grid: scipy's ValueError (fields.py:735-740); rows so far are kept */
#define RAPT_ST_TRACE        -7   /* bounce centre: a field line could not be traced between its mirror points (the reference raises: flutils.py:117,311) */
#define RAPT_ST_ROWCAP      -10   /* adaptive: row buffer full before t0+delta                     */

#define RAPT_MODE_PARTICLE 0
#define RAPT_MODE_GC       1

/* The field-plugin object of rapt/fields.py (_Field, fields.py:5-280) as a POD snapshot. */
typedef struct rapt_field {
    int32_t kind;          /* RAPT_FIELD_*                                                       */
    int32_t is_static;     /* _Field.static            fields.py:41                               */
    int32_t user_id;       /* handle returned by rapt_b200_field_nvrtc (RAPT_FIELD_USER) or rapt_b200_grid_create (RAPT_FIELD_GRID) */
    int32_t nprm;
    double  prm[16];       /* model parameters (constructor arguments, see RAPT_FIELD_*)          */
    double  gradstep;      /* _Field.gradientstepsize  fields.py:39                               */
    double  tstep;         /* _Field.timederivstepsize fields.py:40                               */
} rapt_field_t;

/* By-value snapshot of rapt.params (rapt/__init__.py:21-34), taken at every advance() call. */
typedef struct rapt_params {
    double  rtol, atol;              /* "solvertolerances"                                       */
    double  cyclotronresolution;     /* Particle output step = cyclotron period / this (Particle.py:282) */
    double  epss, epst;              /* adiabaticity thresholds (Particle.py:374-375)             */
    int32_t enforce_equatorial;      /* "enforce equatorial" (Particle.py:296-297, GuidingCenter.py:353-354) */
    int32_t check_adiabaticity;      /* tracer attribute check_adiabaticity (Particle.py:102)     */
    int32_t dop853_reject_rule;      /* 0 = scipy 1.18.1 `_dop` (rejected step -> h/facc1); 1 = Hairer's Fortran */
    int32_t arith;                   /* 0 = fast (FMA contraction, reciprocal multiplies); 1 = strict
                                        (unfused, mirrors the CPU reference's operation order)    */
    int32_t sort_by_work;            /* 1 = schedule tracers longest-first (predicted steps from the initial state);
                                        2 = _dev entry points only: the `counters` buffer still holds this tracer's counts
                                        from the previous call and its step count orders this call (0 there: as 1)     */
    int32_t reserved[3];
} rapt_params_t;

/* ---- library / device management */
int         rapt_b200_init(int device);              /* select device, create context; returns RAPT_OK */
int         rapt_b200_device_count(void);             /* number of CUDA devices, 0 if none            */
const char *rapt_b200_last_error(void);
const char *rapt_b200_version(void);
/* measured FP64 peak of the current device (register-resident DFMA chains on all SMs), TFLOP/s */
int         rapt_b200_fp64_peak(int iters, double *tflops, double *sm_clock_mhz);

/* ---- field-plugin interface (rapt/fields.py)
 * rapt_b200_field_nvrtc: replaces "subclass _Field and override B/E" (examples/Creating new fields.ipynb
 * cells 7,10) for device execution.  `cuda_src` must define
 *     __device__ void rapt_user_B(double t, double x, double y, double z, const double *prm, double *B);
 * and, if has_E != 0,
 *     __device__ void rapt_user_E(double t, double x, double y, double z, const double *prm, double *E);
 * It is JIT-compiled with NVRTC for sm_100a into the same kernel templates as the built-ins and cached
 * by source hash.  On success *user_id receives the handle to put in rapt_field_t.user_id. */
int rapt_b200_field_nvrtc(const char *cuda_src, int has_E, int *user_id, char *log, int loglen);

/* rapt_b200_grid_create: replaces fields.Grid.__init__ / _set_interpolator / _update_interpolator
 * (rapt/fields.py:566-705) -- the parsed data files of a Grid (the dictionaries Grid.parsefile returns,
 * fields.py:553-562) become device-resident interpolation tables.
 *   t[nt]                       time of every data file, ascending (nt == 1: time-independent, 3-D interpolation)
 *   x[nx], y[ny], z[nz]         node coordinates, ascending, uniform spacing not required
 *   Bx,By,Bz,Ex,Ey,Ez           [nt][nx][ny][nz] C-ordered, SI units; Ex,Ey,Ez may all be NULL (E == 0)
 * All nt time points stay resident (the reference's rolling window of three, fields.py:697-705, exists to
 * save host memory; the interpolated value does not depend on it).  Interpolation = scipy
 * RegularGridInterpolator(method="linear", bounds_error=True) as Grid.Bgrid/Egrid call it (fields.py:707-772);
 * a tracer that leaves the grid ends with RAPT_ST_FIELD where the reference raises ValueError.
 * *grid_id receives the handle for rapt_field_t.user_id (kind = RAPT_FIELD_GRID, gradstep = 1e-3 Re as
 * fields.py:578). rapt_b200_grid_destroy frees the tables. */
int rapt_b200_grid_create(int64_t nt, int64_t nx, int64_t ny, int64_t nz,
                          const double *t, const double *x, const double *y, const double *z,
                          const double *Bx, const double *By, const double *Bz,
                          const double *Ex, const double *Ey, const double *Ez, int32_t *grid_id);
int rapt_b200_grid_destroy(int32_t grid_id);

/* Field operators at npt points (tpos = npt x 4: t,x,y,z).  Any output pointer may be NULL.
 * Replaces _Field.B/E/unitb/magB/gradB/jacobianB/curlb/curvature/dBdt/dbdt/lengthscale/timescale
 * (fields.py:43-280) for batched evaluation.  HOST pointers. */
int rapt_b200_field_ops(const rapt_field_t *f, int arith, int64_t npt, const double *tpos,
                        double *B, double *E, double *unitb, double *magB, double *gradB, double *jacobianB,
                        double *curlb, double *curvature, double *dBdt, double *dbdt,
                        double *lengthscale, double *timescale);

/* ---- Particle.advance (Particle.py:230-309) for an ensemble of n independent particles.
 * State in/out: t, x,y,z, px,py,pz = the last trajectory row of each particle.
 * dt_out = output step chosen at entry (cyclotron period / cyclotronresolution, Particle.py:282);
 * tcur = Particle.tcur after the call (one dt past the last row, Particle.py:306);
 * counters = n x 4 int32 (nfcn, nstep, naccpt, nrejct) summed over all solver calls, as scipy
 * would report them (nfcn = 2*rows + 11*nstep + naccpt). */
int rapt_b200_particle_advance(const rapt_field_t *f, const rapt_params_t *p, int64_t n,
                               double *t, double *x, double *y, double *z, double *px, double *py, double *pz,
                               const double *mass, const double *charge, double delta,
                               int64_t store_every, int64_t max_rows, double *rows,
                               int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status,
                               double *tcur, double *dt_out);
int rapt_b200_particle_advance_dev(const rapt_field_t *f, const rapt_params_t *p, int64_t n,
                               double *t, double *x, double *y, double *z, double *px, double *py, double *pz,
                               const double *mass, const double *charge, double delta,
                               int64_t store_every, int64_t max_rows, double *rows,
                               int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status,
                               double *tcur, double *dt_out, void *stream);

/* ---- GuidingCenter.__init__ (GuidingCenter.py:123-133): p_par and mu from speed v and pitch angle pa
 * (degrees; pa == 90 gives p_par = 0 exactly).  HOST pointers. */
int rapt_b200_gc_construct(const rapt_field_t *f, int arith, int64_t n, const double *t0, const double *x,
                           const double *y, const double *z, const double *v, const double *pa_deg,
                           const double *mass, double *ppar, double *mu);

/* ---- GuidingCenter.advance (GuidingCenter.py:397-458).  State in/out: t, X,Y,Z, p_par.
 * mu, v (speed, used by the BrizardChan / NorthropTeller EOMs), mass, charge, dt per particle.
 * dt = params["GCtimestep"], or bounceperiod()/bounceresolution from rapt_b200_bounce_setup. */
int rapt_b200_gc_advance(const rapt_field_t *f, const rapt_params_t *p, int eom, int64_t n,
                         double *t, double *x, double *y, double *z, double *ppar,
                         const double *mu, const double *v, const double *mass, const double *charge,
                         const double *dt, double delta,
                         int64_t store_every, int64_t max_rows, double *rows,
                         int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status, double *tcur);
int rapt_b200_gc_advance_dev(const rapt_field_t *f, const rapt_params_t *p, int eom, int64_t n,
                         double *t, double *x, double *y, double *z, double *ppar,
                         const double *mu, const double *v, const double *mass, const double *charge,
                         const double *dt, double delta,
                         int64_t store_every, int64_t max_rows, double *rows,
                         int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status, double *tcur,
                         void *stream);

/* ---- GuidingCenter.bounceperiod set-up (GuidingCenter.py:593-606 -> flutils.py:254-283 ->
 * fieldline.py:13-105 -> rkf.py:13-143): for each guiding centre compute the mirror field Bm and speed v,
 * the field-line step ds = 1/(curvature*fieldlineresolution), and trace the field line both ways with
 * RKF45 until |B| > Bm.  Outputs per particle: Bm, v, ds, npts and the curve
 * curve[(i*max_pts + k)*5 + c], c = s, x, y, z, |B| ordered as Fieldline.curve.  npts > max_pts means the
 * buffer was too small (retry).  The quadrature over the curve (scipy interp1d/brentq/quad in the
 * reference, flutils.py:295-314) is done by the caller.  With mu == NULL the call is a plain
 * Fieldline(tpos, field, Bmax=Bm).trace(): Bm[] is then an INPUT and ppar, mass, v may be NULL.  HOST pointers. */
int rapt_b200_bounce_setup(const rapt_field_t *f, int arith, double fieldlineresolution, int64_t n,
                           const double *t, const double *x, const double *y, const double *z, const double *ppar,
                           const double *mu, const double *mass,
                           double *Bm, double *v, double *ds, int32_t *npts, int64_t max_pts, double *curve);

/* ---- GuidingCenter.bounceperiod entirely on the device: the trace above followed by
 * flutils.halfbouncepath (flutils.py:274-316) with scipy's quadratic interpolating spline rebuilt per thread.
 *   quadrature = RAPT_QUAD_QUADPACK: the reference's own route -- brentq for the two mirror points and QUADPACK
 *     QAGS (epsabs 1.49e-8, epsrel 1e-4, limit 50) on 1/sqrt(1 - B(s)/Bm), both restated per thread; equals
 *     the reference's value to ~1e-9 (the integrand is evaluated within 1e-10 of its singularities);
 *   quadrature = RAPT_QUAD_CLOSED: mirror points and integral in closed form span by span; no quadrature error,
 *     so it differs from the reference by QUADPACK's own error (1e-7 typical, up to 2e-5 observed).
 * period[i] = NaN if the trace did not bracket both mirror points.  npts may be NULL.  HOST pointers. */
#define RAPT_QUAD_CLOSED   0
#define RAPT_QUAD_QUADPACK 1
int rapt_b200_bounce_period(const rapt_field_t *f, int arith, int quadrature, double fieldlineresolution, int64_t n,
                            const double *t, const double *x, const double *y, const double *z, const double *ppar,
                            const double *mu, const double *mass, double *period, int32_t *npts);

/* ---- BounceCenter.advance (rapt/BounceCenter.py:206-251) for n bounce centres.
 * State in/out: t (the row label = BounceCenter.tcur), x, y, z.  mu, v (speed), mass, charge per tracer.
 * Each tracer: dt = bctimestep * bounceperiod(last row) (params["BCtimestep"], BounceCenter.py:228-229; or
 * dt_in[i] when dt_in != NULL), then len(np.arange(tcur, tcur+delta, dt)) rows, each one scipy "dopri5" call over
 * dt (rtol, atol = params["solvertolerances"]) on dR/dt = gamma m v^2/(q S_b B^2) gradI x B, where S_b =
 * flutils.halfbouncepath (:254-316) and gradI = flutils.gradI (:153-229, step eyegradientstep =
 * params["eyegradientstep"]) trace five field lines per evaluation (fieldlineresolution =
 * params["fieldlineresolution"]).  As in the reference the label of a row is the START time of its step.
 * rows[(i*max_rows + k)*4 + c], c = t, x, y, z (every store_every-th row, row 0 = first computed row);
 * counters = n x 4 int32 (nfcn, nstep, naccpt, nrejct of dopri5, summed over rows); status: RAPT_ST_OK,
 * RAPT_ST_NMAX / RAPT_ST_HSMALL (scipy warns and the reference carries on with the unfinished row; here the
 * tracer stops), RAPT_ST_TRACE.  dt_out (optional) receives the step.  Only static fields (the reference's
 * constructor raises otherwise, BounceCenter.py:104-105); gridded fields are RAPT_E_UNSUPPORTED.  HOST pointers. */
int rapt_b200_bounce_center_advance(const rapt_field_t *f, int arith, int quadrature, int64_t n,
                                    double *t, double *x, double *y, double *z,
                                    const double *mu, const double *v, const double *mass, const double *charge,
                                    const double *dt_in, double bctimestep, double delta,
                                    double rtol, double atol, double fieldlineresolution, double eyegradientstep,
                                    int64_t store_every, int64_t max_rows, double *rows,
                                    int32_t *nrows, int32_t *nstored, int32_t *counters, int32_t *status, double *dt_out);

/* ---- the pieces of that right-hand side at n points, for callers of flutils.halfbouncepath / eye / gradI
 * (rapt/flutils.py:65-316; rapt/__init__.py:42 exports them): out[i*8 + c], c = S_b, I, gradI_x, gradI_y, gradI_z,
 * and BounceCenter.advance's deriv_x, deriv_y, deriv_z (needs v, mass, charge; pass NULL to get NaN there).
 * Bm per point (mirror field).  status[i] as above.  HOST pointers. */
int rapt_b200_bounce_center_terms(const rapt_field_t *f, int arith, int quadrature, int64_t n,
                                  const double *t, const double *x, const double *y, const double *z, const double *Bm,
                                  const double *v, const double *mass, const double *charge,
                                  double fieldlineresolution, double eyegradientstep, double *out, int32_t *status);

/* ---- flutils.eye (rapt/flutils.py:65-151) alone, as GuidingCenter.geteye calls it once per trajectory row
 * (GuidingCenter.py:608-624): I[i] = second invariant of the field line through (t, x, y, z)[i] with mirror field Bm[i].
 * HOST pointers. */
int rapt_b200_second_invariant(const rapt_field_t *f, int arith, int64_t n,
                               const double *t, const double *x, const double *y, const double *z, const double *Bm,
                               double fieldlineresolution, double *I, int32_t *status);

/* ---- Adaptive.__init__ + Adaptive.advance (Adaptive.py:70-104, 187-222) for an ensemble.
 * In: particle position/velocity (as the Adaptive constructor takes them), t0, mass, charge.
 * gc_dt: guiding-centre output step (params["GCtimestep"], must be != 0 for ensembles).
 * Runs epochs of (particle kernel | guiding-centre kernel | switch + compaction kernel) until every
 * tracer reaches `delta` (Adaptive.py:205 compares the absolute tcur with delta -- reproduced).
 * rows as above with the segment tag in column 7 (store_every applies within segments; the first row
 * of every segment is always stored).  Per particle: mode_out (final mode), nseg (number of segments),
 * final state in fin[i*8 + c] (layout of a row), counters n x 4, status.  HOST pointers. */
int rapt_b200_adaptive_advance(const rapt_field_t *f, const rapt_params_t *p, int64_t n,
                               const double *x, const double *y, const double *z,
                               const double *vx, const double *vy, const double *vz,
                               const double *t0, const double *mass, const double *charge,
                               double gc_dt, double delta, int64_t store_every, int64_t max_rows, double *rows,
                               int32_t *nstored, int32_t *nseg, int32_t *mode_out, double *fin,
                               int32_t *counters, int32_t *status, int32_t *epochs_out);

/* What the last rapt_b200_adaptive_advance of the calling thread did, for the roofline of the mix it executed
 * (bench.py --workload adaptive): out[0..14] = epochs, particle-kernel launches, guiding-centre-kernel launches,
 * tracers handed to particle launches (sum over epochs), same for guiding-centre launches, particle-mode attempted
 * steps, accepted steps, solver calls (= rows), guiding-centre attempted steps, solver calls, and the device time in ms
 * of the particle kernels, the guiding-centre kernels (the two run concurrently on two streams), the switch/regroup
 * kernels, and of the whole epoch loop; out[14] reserved; then, while room (n > 15), four numbers per epoch: tracers in
 * particle mode, tracers in guiding-centre mode, particle-kernel ms, guiding-centre-kernel ms. */
int rapt_b200_adaptive_last_stats(double *out, int n);

/* ---- mode-switch transforms, exposed for callers that drive segments themselves:
 * GuidingCenter.init(Particle) (GuidingCenter.py:168-186, utils.py:251-326) and
 * Particle.init(GuidingCenter) (Particle.py:149-164, utils.py:376-433; t_eval = the new Particle's tcur). */
int rapt_b200_switch_p2g(const rapt_field_t *f, int arith, int64_t n, const double *prow7, const double *mass,
                         const double *charge, double *grow5, double *mu, double *v, int32_t *status);
int rapt_b200_switch_g2p(const rapt_field_t *f, int arith, int64_t n, const double *grow5, const double *mu,
                         const double *mass, const double *charge, double t_eval, double *prow7);
/* isadiabatic predicates (Particle.py:345-384, GuidingCenter.py:287-327): out[i] = 0/1 */
int rapt_b200_isadiabatic(const rapt_field_t *f, const rapt_params_t *p, int mode, int64_t n, const double *rows,
                          int64_t row_stride, const double *mu, const double *mass, const double *charge, int32_t *out);

/* ---- final-state diagnostics of one shard (multi-GPU runs; SURVEY.md section 8b `rapt_b200_allgather_final` /
 * `rapt_b200_histogram`, BASELINE.json north_star "NCCL ... only to all-gather final states and diagnostics (histograms,
 * invariants)").  The reference has no counterpart beyond its per-object getters: kind 0 bins log10 of
 * Particle.getke() in eV (Particle.py:442-454), kind 1 bins the radial distance in Earth radii (getr, GuidingCenter.py:
 * 470-473).  ONE kernel pass over the DEVICE state columns cols[0..ncol) (t,x,y,z,px,py,pz or t,X,Y,Z,ppar):
 *   packed  (optional) receives the [n][ncol] rows that the caller's all-gather sends (it may point INTO the gather
 *           buffer at this rank's slot, so no second copy is made);
 *   hist    int64[nbins], ACCUMULATED (zero it first): counts of q in [lo, hi], bins as numpy.histogram;
 *   stats   double[4], ACCUMULATED: tracers with status 1, sum q, sum q^2 over those, tracers outside [lo, hi].
 * The collective itself (NCCL all-gather of `packed`, sum-all-reduce of hist/stats) is issued by the host layer on the
 * same stream through torch.distributed (rapt_b200/dist.py).  `cols` is a HOST array of device pointers. */
int rapt_b200_final_diagnostics_dev(int kind, int64_t n, int ncol, const double *const *cols, const double *mass,
                                    const int32_t *status, double *packed, int nbins, double lo, double hi,
                                    int64_t *hist, double *stats, void *stream);

/* The all-gathered rows, gathered[world][n_max][ncol], rearranged into global member order out[n_total][ncol].  Shards are
 * periodic: of every `period` consecutive members those at positions [offsets[r], offsets[r+1]) belong to rank r, in order
 * (offsets: HOST array of world + 1 ints from 0 to period).  offsets == NULL: round-robin (rank r holds members r,
 * r + world, ...).  rapt_b200/dist.py:ShardPlan builds the table; unequal runs give faster GPUs more tracers.
 * gathered / out: DEVICE pointers. */
int rapt_b200_unshard_dev(int world, int period, const int32_t *offsets, int64_t n_max, int ncol, int64_t n_total,
                          const double *gathered, double *out, void *stream);

/* kernel launches performed by this library since load (for bench.py's gpu_launches) */
int64_t rapt_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
