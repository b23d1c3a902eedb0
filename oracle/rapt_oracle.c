/* rapt_oracle.c -- CPU restatement of RAPT's particle-advance hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker / reported baseline.
 * The product (rapt_b200/, librapt_b200.so) never links, imports or calls it.
 *
 * PARITY PIN: the reference (mkozturk/rapt) has no tests or golden vectors of its own
 * (SURVEY.md §4).  This restatement is pinned against (i) trajectories + scipy solver counters
 * produced by running the UNMODIFIED reference in the build container (oracle/gen_golden.py ->
 * tests/golden/ (npz), numpy 2.3.5 / scipy 1.18.1) and (ii) the notebook outputs stored in the
 * reference (Speiser switch times, DoubleDipole magB values).  tests/test_oracle_golden.py holds
 * those checks.
 *
 * The adaptive Runge-Kutta arithmetic is NOT in the reference tree: it is scipy.integrate.ode
 * "dop853"/"dopri5" (third-party; reference pins scipy==1.3.1 in requirements.txt:2, the container
 * runs scipy 1.18.1 whose _dop is a C translation of Hairer's Fortran).  dop853()/dopri5() below
 * restate the published Hairer-Norsett-Wanner algorithm with scipy's call-site settings
 * (Particle.py:300, GuidingCenter.py:450); the one behavioural difference of scipy 1.18.1
 * (a rejected DOP853 step shrinks by h/facc1 = 0.3 h regardless of err; measured, and bit-exact
 * against the goldens in that form) is selectable against Hairer's h/min(facc1,fac11/safe).
 *
 * Plain C99, fp64, compiled with -ffp-contract=off so the operation order below is what runs.
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "dop_coeffs.h"

#define C_LIGHT 299792458.0            /* rapt/__init__.py:8 */
#define EARTH_B0 3.07e-5               /* rapt/__init__.py:9 */
#define EARTH_RE 6378137.0             /* rapt/__init__.py:10 */

enum { F_EARTHDIPOLE = 0, F_DOUBLEDIPOLE = 1, F_UNIFORMBZ = 2, F_CROSSEDEB = 3, F_VARDIPOLE = 4,
       F_PARABOLIC = 5, F_GRID = 6 /* fields.py:513-814 */, F_CHARGEDDIPOLE = 100 /* examples/Creating new fields.ipynb cell 10 */ };
enum { EOM_TAOCHANBRIZARD = 0, EOM_BRIZARDCHAN = 1, EOM_NORTHROPTELLER = 2 };
enum { ST_OK = 1, ST_NMAX = -2, ST_HSMALL = -3, ST_GCITER = -5, ST_FIELD = -6 /* Grid: ValueError, point outside the grid */ };
enum { MODE_PARTICLE = 0, MODE_GC = 1 };

/* fields.Grid (fields.py:513-814): six components sampled on a rectilinear grid at nt = 1..n time
 * points, each component a C-ordered [nt][nx][ny][nz] array exactly as _set_interpolator builds them */
typedef struct {
    int nt, nx, ny, nz;
    const double *t, *x, *y, *z;
    const double *B[3], *E[3];
} ogrid_t;

typedef struct {
    int kind;
    int is_static;
    double prm[8];
    double gradstep;      /* _Field.gradientstepsize  fields.py:39 */
    double tstep;         /* _Field.timederivstepsize fields.py:40 */
    const ogrid_t *grid;  /* F_GRID only */
} ofield_t;

/* set when a Grid field is evaluated outside its bounds: the reference raises ValueError there
 * (scipy RegularGridInterpolator, bounds_error=True) and the advance in progress is abandoned */
static __thread int g_field_err = 0;
/* diagnostics for the parity report: the smallest |err - 1| over all step attempts of the tracer being integrated
 * (how close its nearest accept/reject decision was), written to g_errgap_out[i] when that is set */
static __thread double tl_errgap = 1e300;
static double *g_errgap_out = 0;
void oracle_set_errgap_out(double *p) { g_errgap_out = p; }

typedef struct {
    double rtol, atol;              /* params["solvertolerances"] __init__.py:27 */
    double cyclotronresolution;     /* __init__.py:22 */
    double gctimestep;              /* __init__.py:25 (0 -> caller supplies bounce-period dt) */
    double epss, epst;              /* __init__.py:31-32 */
    int enforce_equatorial;         /* __init__.py:33 */
    int dop853_reject_rule;         /* 0: scipy 1.18.1 (h/facc1), 1: Hairer (h/min(facc1,fac11/safe)) */
} oparams_t;

/* --------------------------------------------------------------------------------------------
 * Field models: fields.py:301-317, 344-362, 376-390, 413-427, 455-470, 506-511
 * ------------------------------------------------------------------------------------------ */
/* scipy 1.18.1 interpolate/_poly_common.pxi find_interval_ascending: the interval i with
 * x[i] <= v < x[i+1], closed on the right at the last node; -1 outside [x[0], x[n-1]] (or NaN). */
static int grid_interval(const double *x, int n, double v)
{
    if (!(x[0] <= v && v <= x[n - 1])) return -1;
    if (v == x[n - 1]) return n - 2;
    int low = 0, high = n - 2;
    if (v < x[low + 1]) high = low;
    while (low < high) {
        int mid = (high + low) / 2;
        if (v < x[mid]) high = mid;
        else if (v >= x[mid + 1]) low = mid + 1;
        else { low = mid; break; }
    }
    return low;
}

/* Grid.Bgrid / Grid.Egrid (fields.py:707-772) = three scipy RegularGridInterpolator(method="linear") calls:
 * _rgi.py find_indices (index + normalised distance per dimension) and _evaluate_linear (the 2^d vertices
 * of the cell in itertools.product order, first dimension slowest; weight = ((1*w0)*w1)*...; value += v*weight).
 * Dimensions are (t, x, y, z) for two or more time points and (x, y, z) for one. */
static void grid_eval(const ogrid_t *g, const double *const comp[3], const double tp[4], double out[3])
{
    const double *ax[4]; int n[4], nd = 0, idx[4]; double yd[4], v[4];
    if (g->nt >= 2) { ax[nd] = g->t; n[nd] = g->nt; v[nd] = tp[0]; nd++; }
    ax[nd] = g->x; n[nd] = g->nx; v[nd] = tp[1]; nd++;
    ax[nd] = g->y; n[nd] = g->ny; v[nd] = tp[2]; nd++;
    ax[nd] = g->z; n[nd] = g->nz; v[nd] = tp[3]; nd++;
    for (int d = 0; d < nd; d++) {
        idx[d] = grid_interval(ax[d], n[d], v[d]);
        if (idx[d] < 0) { g_field_err = 1; out[0] = out[1] = out[2] = NAN; return; }
        yd[d] = (v[d] - ax[d][idx[d]]) / (ax[d][idx[d] + 1] - ax[d][idx[d]]);
    }
    for (int c = 0; c < 3; c++) {
        double value = 0.0;
        for (int corner = 0; corner < (1 << nd); corner++) {
            double weight = 1.0; long off = 0;
            for (int d = 0; d < nd; d++) {
                int up = (corner >> (nd - 1 - d)) & 1;
                weight = weight * (up ? yd[d] : 1 - yd[d]);
                off = off * n[d] + idx[d] + up;
            }
            double term = comp[c][off] * weight;
            value = value + term;
        }
        out[c] = value;
    }
}

static void field_B(const ofield_t *f, const double tp[4], double B[3])
{
    double t = tp[0], x = tp[1], y = tp[2], z = tp[3];
    switch (f->kind) {
    case F_EARTHDIPOLE: {           /* fields.py:315-317; prm[0] = _coeff = -3*B0*Re**3 */
        double r2 = x * x + y * y + z * z;
        double s = f->prm[0] / pow(r2, 2.5);
        B[0] = s * (x * z); B[1] = s * (y * z); B[2] = s * (z * z - r2 / 3);
        break; }
    case F_DOUBLEDIPOLE: {          /* fields.py:358-362; prm = {_coeff, _dd, _k} */
        double p1 = pow(x * x + y * y + z * z, 5.0 / 2.0);
        double a0 = 3 * x * z / p1, a1 = 3 * y * z / p1, a2 = (2 * z * z - x * x - y * y) / p1;
        x -= f->prm[1];
        double p2 = pow(x * x + y * y + z * z, 5.0 / 2.0);
        double k = f->prm[2];
        double b0 = k * (3 * x * z) / p2, b1 = k * (3 * y * z) / p2, b2 = k * (2 * z * z - x * x - y * y) / p2;
        B[0] = f->prm[0] * (a0 + b0); B[1] = f->prm[0] * (a1 + b1); B[2] = f->prm[0] * (a2 + b2);
        break; }
    case F_UNIFORMBZ: case F_CROSSEDEB:  /* fields.py:390; prm[0] = Bz */
        B[0] = 0; B[1] = 0; B[2] = f->prm[0];
        break;
    case F_VARDIPOLE: {             /* fields.py:469-470; prm = {amp, period} */
        double s = -EARTH_B0 * (EARTH_RE * EARTH_RE * EARTH_RE) * (1 + f->prm[0] * sin(2 * M_PI * t / f->prm[1]));
        double p = pow(x * x + y * y + z * z, 5.0 / 2.0);
        B[0] = s * (3 * x * z) / p; B[1] = s * (3 * y * z) / p; B[2] = s * (2 * z * z - x * x - y * y) / p;
        break; }
    case F_PARABOLIC:               /* fields.py:506-511; prm = {B0, Bn, d}; quirk Q6: module B0 outside */
        if (fabs(z) <= 1.0) B[0] = f->prm[0] * z / f->prm[2];
        else B[0] = (z > 0 ? 1.0 : (z < 0 ? -1.0 : 0.0)) * EARTH_B0;
        B[1] = 0; B[2] = f->prm[1];
        break;
    case F_GRID: grid_eval(f->grid, f->grid->B, tp, B); break;      /* fields.py:774-794 */
    case F_CHARGEDDIPOLE: {         /* notebook field; prm = {B0, Q, k} */
        double p = pow(x * x + y * y + z * z, 5.0 / 2.0);
        B[0] = f->prm[0] * (3 * x * z) / p; B[1] = f->prm[0] * (3 * y * z) / p;
        B[2] = f->prm[0] * (2 * z * z - x * x - y * y) / p;
        break; }
    default: B[0] = B[1] = B[2] = 0;
    }
}

static int field_has_E(const ofield_t *f) { return f->kind == F_CROSSEDEB || f->kind == F_CHARGEDDIPOLE || f->kind == F_GRID; }

static void field_E(const ofield_t *f, const double tp[4], double E[3])
{
    E[0] = E[1] = E[2] = 0;         /* fields.py:59-74 */
    if (f->kind == F_GRID) { grid_eval(f->grid, f->grid->E, tp, E); return; }      /* fields.py:796-814 */
    if (f->kind == F_CROSSEDEB) E[1] = f->prm[1];      /* fields.py:427; prm = {Bz, Ey} */
    else if (f->kind == F_CHARGEDDIPOLE) {
        double x = tp[1], y = tp[2], z = tp[3];
        double p = pow(x * x + y * y + z * z, 3.0 / 2.0), kq = f->prm[2] * f->prm[1];
        E[0] = kq * x / p; E[1] = kq * y / p; E[2] = kq * z / p;
    }
}

/* np.dot on 3-vectors as executed by the container's numpy (OpenBLAS 0.3.30 ddot): a fused chain
 * fma(a2,b2, fma(a1,b1, a0*b0)) -- determined by exhaustive comparison, see DESIGN.md. */
/* Python/numpy scalar `x**2` is libm pow(x, 2.0) (0.52 ulp, not always == x*x): mirrored for bit parity */
static double sq(double x) { return pow(x, 2.0); }
static double dot3(const double a[3], const double b[3]) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* fields.py:90-91, 108-109 */
static double field_magB(const ofield_t *f, const double tp[4])
{
    double B[3]; field_B(f, tp, B); return sqrt(dot3(B, B));
}
static void field_unitb(const ofield_t *f, const double tp[4], double b[3])
{
    double B[3]; field_B(f, tp, B);
    double m = sqrt(dot3(B, B));
    b[0] = B[0] / m; b[1] = B[1] / m; b[2] = B[2] / m;
}
static void shifted(const double tp[4], int axis, double d, double out[4])
{
    out[0] = tp[0]; out[1] = tp[1]; out[2] = tp[2]; out[3] = tp[3];
    out[axis] = tp[axis] + d;
}
/* fields.py:125-131 */
static void field_gradB(const ofield_t *f, const double tp[4], double g[3])
{
    double d = f->gradstep, a[4], b[4];
    for (int i = 0; i < 3; i++) {
        shifted(tp, i + 1, d, a); shifted(tp, i + 1, -d, b);
        g[i] = (field_magB(f, a) - field_magB(f, b)) / (2 * d);
    }
}
/* fields.py:148-153: J[i][j] = dB_i/dx_j */
static void field_jacobianB(const ofield_t *f, const double tp[4], double J[3][3])
{
    double d = f->gradstep, a[4], b[4], Bp[3], Bm[3];
    for (int j = 0; j < 3; j++) {
        shifted(tp, j + 1, d, a); shifted(tp, j + 1, -d, b);
        field_B(f, a, Bp); field_B(f, b, Bm);
        for (int i = 0; i < 3; i++) J[i][j] = (Bp[i] - Bm[i]) / (2 * d);
    }
}
/* fields.py:170-174 (quirk Q7: np.dot(gB, B) with the SCALAR B is element-wise) */
static double field_curvature(const ofield_t *f, const double tp[4])
{
    double Bv[3], gB[3], gp[3];
    field_B(f, tp, Bv);
    double B = sqrt(dot3(Bv, Bv));
    field_gradB(f, tp, gB);
    for (int i = 0; i < 3; i++) gp[i] = gB[i] - ((gB[i] * B) / sq(B)) * Bv[i];
    return sqrt(dot3(gp, gp)) / B;
}
/* fields.py:191-200 with _M1 (fields.py:33-36): beta = [b(+x) b(-x) b(+y) b(-y) b(+z) b(-z)] */
static void field_curlb(const ofield_t *f, const double tp[4], double cb[3])
{
    double d = f->gradstep, beta[18], q[4];
    for (int j = 0; j < 3; j++) {
        shifted(tp, j + 1, d, q);  field_unitb(f, q, beta + 6 * j);
        shifted(tp, j + 1, -d, q); field_unitb(f, q, beta + 6 * j + 3);
    }
    /* row0: +beta[8] -beta[11] -beta[13] +beta[16]; row1: -beta[2] +beta[5] +beta[12] -beta[15];
       row2: +beta[1] -beta[4] -beta[6] +beta[9] */
    /* summation order of np.dot(_M1, beta) as executed by the container's OpenBLAS dgemv
       (found by exhaustive search over orders; zeros do not contribute) */
    cb[0] = ((beta[8] + (-beta[11] + -beta[13])) + beta[16]) / (2 * d);
    cb[1] = ((-beta[2] + beta[12]) + (beta[5] + -beta[15])) / (2 * d);
    cb[2] = ((beta[1] + beta[9]) + (-beta[4] + -beta[6])) / (2 * d);
}
/* fields.py:217-223 */
static double field_dBdt(const ofield_t *f, const double tp[4])
{
    if (f->is_static) return 0;
    double d = f->tstep, a[4], b[4];
    shifted(tp, 0, -d, a); shifted(tp, 0, d, b);
    return (field_magB(f, b) - field_magB(f, a)) / d / 2;
}
/* fields.py:239-245 */
static void field_dbdt(const ofield_t *f, const double tp[4], double o[3])
{
    if (f->is_static) { o[0] = o[1] = o[2] = 0; return; }
    double d = f->tstep, a[4], b[4], b1[3], b2[3];
    shifted(tp, 0, -d, a); shifted(tp, 0, d, b);
    field_unitb(f, a, b1); field_unitb(f, b, b2);
    for (int i = 0; i < 3; i++) o[i] = (b2[i] - b1[i]) / d / 2;
}
/* fields.py:261 */
static double field_lengthscale(const ofield_t *f, const double tp[4])
{
    double J[3][3], m = 0;
    field_jacobianB(f, tp, J);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = fabs(J[i][j]); if (a > m) m = a; }
    return field_magB(f, tp) / m;
}
/* fields.py:277-280 (static -> None; callers never ask in that case) */
static double field_timescale(const ofield_t *f, const double tp[4])
{
    return field_magB(f, tp) / fabs(field_dBdt(f, tp));
}

/* --------------------------------------------------------------------------------------------
 * scipy.integrate.ode "dop853" / "dopri5" (Hairer, Norsett & Wanner), one call = one output row.
 * Settings from the call sites: nsteps=500, safety=0.9, max_step=0 (hmax=|xend-x|), first_step=0
 * (HINIT), uround=2.3e-16; dop853: dfactor .3, ifactor 6, beta 0.1 (Particle.py:300);
 * dopri5: dfactor .2, ifactor 10, beta 0 -> the solver's default 0.04 (GuidingCenter.py:450).
 * ------------------------------------------------------------------------------------------ */
typedef void (*rhs_fn)(double t, const double *y, double *dy, void *ctx);
typedef struct { long nfcn, nstep, naccpt, nrejct; } ocount_t;
#define NMAXD 6

static double hinit(int n, rhs_fn f, void *ctx, double x, const double *y, double posneg,
                    const double *f0, int iord, double hmax, double atol, double rtol)
{
    double dnf = 0, dny = 0, y1[NMAXD], f1[NMAXD];
    for (int i = 0; i < n; i++) {
        double sk = atol + rtol * fabs(y[i]);
        dnf += (f0[i] / sk) * (f0[i] / sk);
        dny += (y[i] / sk) * (y[i] / sk);
    }
    double h = (dnf <= 1e-10 || dny <= 1e-10) ? 1e-6 : sqrt(dny / dnf) * 0.01;
    h = fmin(h, hmax);
    h = copysign(h, posneg);
    for (int i = 0; i < n; i++) y1[i] = y[i] + h * f0[i];
    f(x + h, y1, f1, ctx);
    double der2 = 0;
    for (int i = 0; i < n; i++) {
        double sk = atol + rtol * fabs(y[i]);
        der2 += ((f1[i] - f0[i]) / sk) * ((f1[i] - f0[i]) / sk);
    }
    der2 = sqrt(der2) / h;
    double der12 = fmax(fabs(der2), sqrt(dnf));
    double h1 = (der12 <= 1e-15) ? fmax(1e-6, fabs(h) * 1e-3) : pow(0.01 / der12, 1.0 / iord);
    h = fmin(fmin(100 * fabs(h), h1), hmax);
    return copysign(h, posneg);
}

/* returns idid (1 ok, -2 nmax, -3 step too small); *xio, y updated in place */
static int dop853(int n, rhs_fn f, void *ctx, double *xio, double *y, double xend,
                  double rtol, double atol, int reject_rule, ocount_t *cnt)
{
    const double beta = 0.1, safe = 0.9, fac1 = 0.3, fac2 = 6.0, uround = 2.3e-16;
    const int nmax = 500;
    double x = *xio;
    double facold = 1e-4, expo1 = 1.0 / 8.0 - beta * 0.2, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    double posneg = copysign(1.0, xend - x), hmax = fabs(xend - x);
    double k1[NMAXD], k2[NMAXD], k3[NMAXD], k4[NMAXD], k5[NMAXD], k6[NMAXD], k7[NMAXD], k8[NMAXD],
           k9[NMAXD], k10[NMAXD], y1[NMAXD];
    int last = 0, reject = 0;
    long nstep = 0, naccpt = 0, nrejct = 0, nfcn = 0;
    f(x, y, k1, ctx);
    double h = hinit(n, f, ctx, x, y, posneg, k1, 8, hmax, atol, rtol);
    nfcn += 2;
    int idid;
    for (;;) {
        if (g_field_err) { idid = ST_FIELD; break; }
        if (nstep > nmax) { idid = ST_NMAX; break; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { idid = ST_HSMALL; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        nstep++;
        int i;
        for (i = 0; i < n; i++) y1[i] = y[i] + h * D8_A2_1 * k1[i];
        f(x + D8_C2 * h, y1, k2, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A3_1 * k1[i] + D8_A3_2 * k2[i]);
        f(x + D8_C3 * h, y1, k3, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A4_1 * k1[i] + D8_A4_3 * k3[i]);
        f(x + D8_C4 * h, y1, k4, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A5_1 * k1[i] + D8_A5_3 * k3[i] + D8_A5_4 * k4[i]);
        f(x + D8_C5 * h, y1, k5, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A6_1 * k1[i] + D8_A6_4 * k4[i] + D8_A6_5 * k5[i]);
        f(x + D8_C6 * h, y1, k6, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A7_1 * k1[i] + D8_A7_4 * k4[i] + D8_A7_5 * k5[i] + D8_A7_6 * k6[i]);
        f(x + D8_C7 * h, y1, k7, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A8_1 * k1[i] + D8_A8_4 * k4[i] + D8_A8_5 * k5[i] + D8_A8_6 * k6[i] + D8_A8_7 * k7[i]);
        f(x + D8_C8 * h, y1, k8, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A9_1 * k1[i] + D8_A9_4 * k4[i] + D8_A9_5 * k5[i] + D8_A9_6 * k6[i] + D8_A9_7 * k7[i] + D8_A9_8 * k8[i]);
        f(x + D8_C9 * h, y1, k9, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A10_1 * k1[i] + D8_A10_4 * k4[i] + D8_A10_5 * k5[i] + D8_A10_6 * k6[i] + D8_A10_7 * k7[i] + D8_A10_8 * k8[i] + D8_A10_9 * k9[i]);
        f(x + D8_C10 * h, y1, k10, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A11_1 * k1[i] + D8_A11_4 * k4[i] + D8_A11_5 * k5[i] + D8_A11_6 * k6[i] + D8_A11_7 * k7[i] + D8_A11_8 * k8[i] + D8_A11_9 * k9[i] + D8_A11_10 * k10[i]);
        f(x + D8_C11 * h, y1, k2, ctx);
        double xph = x + h;
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D8_A12_1 * k1[i] + D8_A12_4 * k4[i] + D8_A12_5 * k5[i] + D8_A12_6 * k6[i] + D8_A12_7 * k7[i] + D8_A12_8 * k8[i] + D8_A12_9 * k9[i] + D8_A12_10 * k10[i] + D8_A12_11 * k2[i]);
        f(xph, y1, k3, ctx);
        nfcn += 11;
        for (i = 0; i < n; i++) {
            k4[i] = D8_B1 * k1[i] + D8_B6 * k6[i] + D8_B7 * k7[i] + D8_B8 * k8[i] + D8_B9 * k9[i] + D8_B10 * k10[i] + D8_B11 * k2[i] + D8_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }
        double err = 0, err2 = 0;
        for (i = 0; i < n; i++) {
            double sk = atol + rtol * fmax(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - D8_BHH1 * k1[i] - D8_BHH2 * k9[i] - D8_BHH3 * k3[i];
            err2 += (erri / sk) * (erri / sk);
            erri = D8_ER1 * k1[i] + D8_ER6 * k6[i] + D8_ER7 * k7[i] + D8_ER8 * k8[i] + D8_ER9 * k9[i] + D8_ER10 * k10[i] + D8_ER11 * k2[i] + D8_ER12 * k3[i];
            err += (erri / sk) * (erri / sk);
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (n * deno));
        if (fabs(err - 1.0) < tl_errgap) tl_errgap = fabs(err - 1.0);
        double fac11 = pow(err, expo1);
        double fac = fac11 / pow(facold, beta);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0) {
            facold = fmax(err, 1e-4);
            naccpt++;
            f(xph, k5, k4, ctx);
            nfcn++;
            for (i = 0; i < n; i++) { k1[i] = k4[i]; y[i] = k5[i]; }
            x = xph;
            if (last) { idid = ST_OK; break; }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            if (reject_rule == 1) hnew = h / fmin(facc1, fac11 / safe);
            else hnew = h / facc1;
            reject = 1;
            if (naccpt >= 1) nrejct++;
            last = 0;
        }
        h = hnew;
    }
    *xio = x;
    if (cnt) { cnt->nfcn += nfcn; cnt->nstep += nstep; cnt->naccpt += naccpt; cnt->nrejct += nrejct; }
    return idid;
}

static int dopri5(int n, rhs_fn f, void *ctx, double *xio, double *y, double xend,
                  double rtol, double atol, ocount_t *cnt)
{
    const double beta = 0.04, safe = 0.9, fac1 = 0.2, fac2 = 10.0, uround = 2.3e-16;
    const int nmax = 500;
    double x = *xio;
    double facold = 1e-4, expo1 = 0.2 - beta * 0.75, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    double posneg = copysign(1.0, xend - x), hmax = fabs(xend - x);
    double k1[NMAXD], k2[NMAXD], k3[NMAXD], k4[NMAXD], k5[NMAXD], k6[NMAXD], y1[NMAXD], ysti[NMAXD];
    int last = 0, reject = 0;
    long nstep = 0, naccpt = 0, nrejct = 0, nfcn = 0;
    f(x, y, k1, ctx);
    double h = hinit(n, f, ctx, x, y, posneg, k1, 5, hmax, atol, rtol);
    nfcn += 2;
    int idid;
    for (;;) {
        if (g_field_err) { idid = ST_FIELD; break; }
        if (nstep > nmax) { idid = ST_NMAX; break; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { idid = ST_HSMALL; break; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        nstep++;
        int i;
        for (i = 0; i < n; i++) y1[i] = y[i] + h * D5_A2_1 * k1[i];
        f(x + D5_C2 * h, y1, k2, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D5_A3_1 * k1[i] + D5_A3_2 * k2[i]);
        f(x + D5_C3 * h, y1, k3, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D5_A4_1 * k1[i] + D5_A4_2 * k2[i] + D5_A4_3 * k3[i]);
        f(x + D5_C4 * h, y1, k4, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D5_A5_1 * k1[i] + D5_A5_2 * k2[i] + D5_A5_3 * k3[i] + D5_A5_4 * k4[i]);
        f(x + D5_C5 * h, y1, k5, ctx);
        for (i = 0; i < n; i++) ysti[i] = y[i] + h * (D5_A6_1 * k1[i] + D5_A6_2 * k2[i] + D5_A6_3 * k3[i] + D5_A6_4 * k4[i] + D5_A6_5 * k5[i]);
        double xph = x + h;
        f(xph, ysti, k6, ctx);
        for (i = 0; i < n; i++) y1[i] = y[i] + h * (D5_A7_1 * k1[i] + D5_A7_3 * k3[i] + D5_A7_4 * k4[i] + D5_A7_5 * k5[i] + D5_A7_6 * k6[i]);
        f(xph, y1, k2, ctx);
        for (i = 0; i < n; i++)
            k4[i] = (D5_E1 * k1[i] + D5_E3 * k3[i] + D5_E4 * k4[i] + D5_E5 * k5[i] + D5_E6 * k6[i] + D5_E7 * k2[i]) * h;
        nfcn += 6;
        double err = 0;
        for (i = 0; i < n; i++) {
            double sk = atol + rtol * fmax(fabs(y[i]), fabs(y1[i]));
            err += (k4[i] / sk) * (k4[i] / sk);
        }
        err = sqrt(err / n);
        if (fabs(err - 1.0) < tl_errgap) tl_errgap = fabs(err - 1.0);
        double fac11 = pow(err, expo1);
        double fac = fac11 / pow(facold, beta);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0) {
            facold = fmax(err, 1e-4);
            naccpt++;
            for (i = 0; i < n; i++) { k1[i] = k2[i]; y[i] = y1[i]; }
            x = xph;
            if (last) { idid = ST_OK; break; }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            hnew = h / fmin(facc1, fac11 / safe);
            reject = 1;
            if (naccpt >= 1) nrejct++;
            last = 0;
        }
        h = hnew;
    }
    *xio = x;
    if (cnt) { cnt->nfcn += nfcn; cnt->nstep += nstep; cnt->naccpt += naccpt; cnt->nrejct += nrejct; }
    return idid;
}

/* --------------------------------------------------------------------------------------------
 * utils.py helpers
 * ------------------------------------------------------------------------------------------ */
/* utils.py:63-66 */
static double cyclotron_period(const ofield_t *f, double t, const double pos[3], const double vel[3],
                               double mass, double charge)
{
    double gamma = 1.0 / sqrt(1 - dot3(vel, vel) / (C_LIGHT * C_LIGHT));
    double tp[4] = { t, pos[0], pos[1], pos[2] };
    double B = field_magB(f, tp);
    return 2 * M_PI * gamma * mass / B / fabs(charge);
}
/* utils.py:102-104 */
static double cyclotron_period2(const ofield_t *f, double t, const double pos[3], double speed,
                                double mass, double charge)
{
    double gamma = 1.0 / sqrt(1 - sq(speed / C_LIGHT));
    double tp[4] = { t, pos[0], pos[1], pos[2] };
    double B = field_magB(f, tp);
    return 2 * M_PI * gamma * mass / B / fabs(charge);
}
/* utils.py:139-145 */
static double cyclotron_radius(const ofield_t *f, double t, const double pos[3], const double vel[3],
                               double mass, double charge)
{
    double vsq = dot3(vel, vel);
    double gamma = 1.0 / sqrt(1 - vsq / (C_LIGHT * C_LIGHT));
    double tp[4] = { t, pos[0], pos[1], pos[2] }, B[3];
    field_B(f, tp, B);
    double Bmag = sqrt(dot3(B, B));
    double vpar = dot3(vel, B) / Bmag;
    double vperp = sqrt(vsq - sq(vpar));
    return gamma * mass * vperp / (fabs(charge) * Bmag);
}
/* utils.py:183-187 */
static double cyclotron_radius2(const ofield_t *f, double t, const double pos[3], double vpar, double v,
                                double mass, double charge)
{
    double gamma = 1.0 / sqrt(1 - sq(v / C_LIGHT));
    double tp[4] = { t, pos[0], pos[1], pos[2] }, B[3];
    field_B(f, tp, B);
    double Bmag = sqrt(dot3(B, B));
    double vperp = sqrt((v - vpar) * (v + vpar));
    return gamma * mass * vperp / (fabs(charge) * Bmag);
}
/* utils.py:214-216 */
static double magnetic_moment(const ofield_t *f, double t, const double pos[3], double vpar, double v, double mass)
{
    double gamma = 1.0 / sqrt(1 - sq(v / C_LIGHT));
    double tp[4] = { t, pos[0], pos[1], pos[2] };
    double Bmag = field_magB(f, tp);
    return sq(gamma) * mass * (v - vpar) * (v + vpar) / (2 * Bmag);
}
/* utils.py:298-303 */
static void gyrovector(const ofield_t *f, double t, const double r[3], const double v[3], double mass,
                       double charge, double out[3])
{
    double vsq = dot3(v, v);
    double gamma = 1 / sqrt(1 - vsq / (C_LIGHT * C_LIGHT));
    double tp[4] = { t, r[0], r[1], r[2] }, B[3], cr[3];
    field_B(f, tp, B);
    double Bsq = dot3(B, B);
    cross3(B, v, cr);
    double s = gamma * mass / (charge * Bsq);
    out[0] = s * cr[0]; out[1] = s * cr[1]; out[2] = s * cr[2];
}
/* utils.py:251-326; returns 0 on convergence, -1 otherwise (reference prints and returns None) */
static int guidingcenter(const ofield_t *f, double t, const double r[3], const double v[3], double mass,
                         double charge, double R[3], double *vp, double *spd)
{
    const double tol = 1e-3; const int maxiter = 20;
    double g[3], old[3], gc[3], d[3];
    gyrovector(f, t, r, v, mass, charge, g);
    for (int i = 0; i < 3; i++) old[i] = r[i] - g[i];
    for (int it = 1; it <= maxiter; it++) {
        gyrovector(f, t, old, v, mass, charge, g);
        for (int i = 0; i < 3; i++) { gc[i] = r[i] - g[i]; d[i] = gc[i] - old[i]; }
        if (sqrt(dot3(d, d)) / sqrt(dot3(gc, gc)) < tol) {
            double tp[4] = { t, gc[0], gc[1], gc[2] }, B[3];
            field_B(f, tp, B);
            *vp = dot3(v, B) / sqrt(dot3(B, B));
            *spd = sqrt(dot3(v, v));
            R[0] = gc[0]; R[1] = gc[1]; R[2] = gc[2];
            return 0;
        }
        old[0] = gc[0]; old[1] = gc[1]; old[2] = gc[2];
    }
    return -1;
}
/* utils.py:360-374 (zero vector raises in the reference; here returns x-hat) */
static void getperp(const double v[3], double o[3])
{
    if (v[0] == 0) { o[0] = 1; o[1] = 0; o[2] = 0; return; }
    if (v[1] == 0) { o[0] = 0; o[1] = 1; o[2] = 0; return; }
    if (v[2] == 0) { o[0] = 0; o[1] = 0; o[2] = 1; return; }
    double cc = -1.0 * (v[0] + v[1]) / v[2];
    double nrm = sqrt(2 + sq(cc));
    o[0] = 1 / nrm; o[1] = 1 / nrm; o[2] = cc / nrm;
}
/* utils.py:422-433 */
static void GCtoFP(const ofield_t *f, double t, const double R[3], double vp, double speed, double mass,
                   double charge, double gyrophase, double pos[3], double vel[3])
{
    double tp[4] = { t, R[0], R[1], R[2] }, B[3], b[3], u[3], w[3];
    field_B(f, tp, B);
    double Bsq = dot3(B, B), sB = sqrt(Bsq);
    for (int i = 0; i < 3; i++) b[i] = B[i] / sB;
    double pa = acos(vp / speed);
    double rc = cyclotron_radius2(f, t, R, vp, speed, mass, charge);
    getperp(B, u);
    double un = sqrt(dot3(u, u));
    for (int i = 0; i < 3; i++) u[i] = u[i] / un;
    cross3(b, u, w);
    double s = (charge > 0) ? 1.0 : (charge < 0 ? -1.0 : 0.0);
    double cg = cos(gyrophase), sg = sin(gyrophase), cpa = cos(pa), spa = sin(pa);
    for (int i = 0; i < 3; i++) {
        pos[i] = R[i] + rc * (cg * u[i] + sg * w[i]);
        vel[i] = speed * ((cpa * b[i] + s * spa * sg * u[i]) - s * spa * cg * w[i]);
    }
}

/* --------------------------------------------------------------------------------------------
 * Particle: Particle.py:59-109 (constructor), :230-309 (advance), :345-384 (isadiabatic)
 * ------------------------------------------------------------------------------------------ */
typedef struct { const ofield_t *f; const oparams_t *p; double mass, charge, gm; } pctx_t;

/* Particle.py:284-298 */
static void particle_eom(double t, const double *Y, double *out, void *vctx)
{
    pctx_t *c = (pctx_t *)vctx;
    double tp[4] = { t, Y[0], Y[1], Y[2] }, E[3], B[3], cr[3];
    if (!c->f->is_static)
        c->gm = sqrt(sq(c->mass) + dot3(Y + 3, Y + 3) / (C_LIGHT * C_LIGHT));
    double gm = c->gm;
    out[0] = Y[3] / gm; out[1] = Y[4] / gm; out[2] = Y[5] / gm;
    field_E(c->f, tp, E); field_B(c->f, tp, B);
    cross3(Y + 3, B, cr);
    for (int i = 0; i < 3; i++) out[3 + i] = c->charge * (E[i] + cr[i] / gm);
    if (c->p->enforce_equatorial) { out[2] = 0; out[5] = 0; }
}

/* Particle.py:380-384 with cycrad :483-488, cycper :489-494 */
static int particle_isadiabatic(const ofield_t *f, const oparams_t *p, const double row[7], double mass, double charge)
{
    const double *mom = row + 4;
    double gm = sqrt(sq(mass) + dot3(mom, mom) / (C_LIGHT * C_LIGHT));
    double v[3] = { mom[0] / gm, mom[1] / gm, mom[2] / gm };
    double tp[4] = { row[0], row[1], row[2], row[3] };
    int sp = cyclotron_radius(f, row[0], row + 1, v, mass, charge) / field_lengthscale(f, tp) < p->epss;
    if (f->is_static) return sp;
    if (!sp) return 0;
    return cyclotron_period(f, row[0], row + 1, v, mass, charge) / field_timescale(f, tp) < p->epst;
}

/* One Particle.advance(delta) call.  state = (t, x,y,z, px,py,pz) = the last trajectory row.
 * rows (optional): capacity max_rows x 8 doubles; stored rows are appended from *nstored on; a row is
 * stored when (global row index % store_every == 0); col 7 = cumulative nstep of this call.
 * percall (optional): capacity max_calls x 4 longs, per-row solver counters (nfcn,nstep,naccpt,nrejct).
 * Returns 1 normal end, 2 Adiabatic raised (check_adiabaticity), negative = solver failure (loop ended
 * silently in the reference: `while r.successful()`). */
static int particle_advance_one(const ofield_t *f, const oparams_t *p, double state[7], double mass,
                                double charge, double delta, int check_adiab,
                                double *rows, long max_rows, long store_every, long *row_index, long *nstored,
                                long *percall, long max_calls, long *ncalls,
                                ocount_t *cnt, double *tcur, double *dt_out)
{
    double t0 = state[0];
    const double *mom = state + 4;
    pctx_t ctx = { f, p, mass, charge, 0 };
    ctx.gm = sqrt(sq(mass) + dot3(mom, mom) / (C_LIGHT * C_LIGHT));             /* :274 */
    double vel[3] = { mom[0] / ctx.gm, mom[1] / ctx.gm, mom[2] / ctx.gm };           /* :275 */
    double dt = cyclotron_period(f, t0, state + 1, vel, mass, charge) / p->cyclotronresolution;  /* :282 */
    if (dt_out) *dt_out = dt;
    double x = t0, y[6];
    memcpy(y, state + 1, 6 * sizeof(double));
    int ok = 1, ret = 1;
    while (ok && x < t0 + delta) {                                                   /* :304 */
        double label = x + dt, xend = x + dt;                                       /* :305 */
        ocount_t c1 = { 0, 0, 0, 0 };
        int idid = dop853(6, particle_eom, &ctx, &x, y, xend, p->rtol, p->atol, p->dop853_reject_rule, &c1);
        if (g_field_err) return ST_FIELD;      /* ValueError out of r.integrate(): no row appended */
        if (cnt) { cnt->nfcn += c1.nfcn; cnt->nstep += c1.nstep; cnt->naccpt += c1.naccpt; cnt->nrejct += c1.nrejct; }
        if (percall && *ncalls < max_calls) {
            long *q = percall + 4 * (*ncalls); q[0] = c1.nfcn; q[1] = c1.nstep; q[2] = c1.naccpt; q[3] = c1.nrejct;
        }
        if (ncalls) (*ncalls)++;
        if (idid < 0) { ok = 0; ret = idid; }
        *tcur = x + dt;                                                              /* :306 */
        state[0] = label; memcpy(state + 1, y, 6 * sizeof(double));                  /* :307 */
        (*row_index)++;
        if (rows && store_every > 0 && (*row_index % store_every) == 0 && *nstored < max_rows) {
            double *r = rows + 8 * (*nstored);
            memcpy(r, state, 7 * sizeof(double)); r[7] = cnt ? (double)cnt->nstep : 0;
            (*nstored)++;
        }
        if (check_adiab && particle_isadiabatic(f, p, state, mass, charge)) return 2; /* :308-309 */
    }
    return ret;
}

/* --------------------------------------------------------------------------------------------
 * GuidingCenter: GuidingCenter.py:63-133 (constructor), :329-395 (EOMs), :397-458 (advance),
 * :287-327 (isadiabatic), :517-541 (cycrad, cycper)
 * ------------------------------------------------------------------------------------------ */
typedef struct { const ofield_t *f; const oparams_t *p; double mass, charge, mu, v; int eom; } gctx_t;

static void gc_eom(double t, const double *Y, double *out, void *vctx)
{
    gctx_t *c = (gctx_t *)vctx;
    const ofield_t *f = c->f;
    double tp[4] = { t, Y[0], Y[1], Y[2] }, ppar = Y[3];
    double B[3], ub[3], cb[3], gB[3], Bs[3], cr[3];
    double m = c->mass, q = c->charge, mu = c->mu;
    field_B(f, tp, B);
    double Bmag = sqrt(dot3(B, B));
    for (int i = 0; i < 3; i++) ub[i] = B[i] / Bmag;
    if (c->eom == EOM_TAOCHANBRIZARD) {            /* GuidingCenter.py:336-355 */
        double gamma = sqrt(1 + 2 * mu * Bmag / (m * C_LIGHT * C_LIGHT) + sq(ppar / (m * C_LIGHT)));
        double E[3], dbdt[3], Es[3];
        field_curlb(f, tp, cb);
        for (int i = 0; i < 3; i++) Bs[i] = B[i] + ppar * cb[i] / q;
        double Bsp = dot3(Bs, ub);
        field_E(f, tp, E);
        field_dbdt(f, tp, dbdt);
        field_gradB(f, tp, gB);
        for (int i = 0; i < 3; i++) Es[i] = E[i] - (ppar * dbdt[i] + mu * gB[i] / gamma) / q;
        cross3(Es, ub, cr);
        for (int i = 0; i < 3; i++) out[i] = (ppar * Bs[i] / (gamma * m) + cr[i]) / Bsp;
        out[3] = q * dot3(Es, Bs) / Bsp;
    } else if (c->eom == EOM_BRIZARDCHAN) {        /* GuidingCenter.py:364-379 */
        double gamma = 1.0 / sqrt(1 - sq(c->v / C_LIGHT));
        field_gradB(f, tp, gB);
        field_curlb(f, tp, cb);
        for (int i = 0; i < 3; i++) Bs[i] = B[i] + ppar * cb[i] / q;
        double Bsp = dot3(Bs, ub);
        cross3(ub, gB, cr);
        for (int i = 0; i < 3; i++) out[i] = (ppar * Bs[i] / (gamma * m) + mu * cr[i] / (q * gamma)) / Bsp;
        out[3] = -mu * dot3(Bs, gB) / (gamma * Bsp);
    } else {                                       /* GuidingCenter.py:382-395 */
        double gamma = 1.0 / sqrt(1 - sq(c->v / C_LIGHT));
        double gm = gamma * m;
        field_gradB(f, tp, gB);
        cross3(ub, gB, cr);
        double s = (gm * sq(c->v) + sq(ppar) / gm) / (2 * q * sq(Bmag));
        for (int i = 0; i < 3; i++) out[i] = s * cr[i] + ppar * ub[i] / gm;
        out[3] = -mu * dot3(ub, gB) / gamma;
    }
    if (c->p->enforce_equatorial) { out[2] = 0; out[3] = 0; }
}

/* GuidingCenter.py:517-529 */
static double gc_cycrad(const ofield_t *f, const double row[5], double mu, double mass, double charge)
{
    double tp[4] = { row[0], row[1], row[2], row[3] }, pp = row[4], vp, v;
    double Bmag = field_magB(f, tp);
    double gamma = sqrt(1 + 2 * mu * Bmag / (mass * C_LIGHT * C_LIGHT) + sq(pp / mass / C_LIGHT));
    if (gamma - 1 < 1e-6) { vp = pp / mass; v = sqrt(2 * mu * Bmag / mass + sq(vp)); }
    else { vp = pp / mass / gamma; v = C_LIGHT * sqrt(1 - 1 / sq(gamma)); }
    return cyclotron_radius2(f, row[0], row + 1, vp, v, mass, charge);
}
/* GuidingCenter.py:531-541 (quirk Q10: pp**2 without /(mc)) */
static double gc_cycper(const ofield_t *f, const double row[5], double mu, double mass, double charge)
{
    double tp[4] = { row[0], row[1], row[2], row[3] }, pp = row[4], v;
    double Bmag = field_magB(f, tp);
    double gamma = sqrt(1 + 2 * mu * Bmag / (mass * C_LIGHT * C_LIGHT) + sq(pp));
    if (gamma - 1 < 1e-6) { double vp = pp / mass; v = sqrt(2 * mu * Bmag / mass + sq(vp)); }
    else v = C_LIGHT * sqrt(1 - 1 / sq(gamma));
    return cyclotron_period2(f, row[0], row + 1, v, mass, charge);
}
/* GuidingCenter.py:323-327 */
static int gc_isadiabatic(const ofield_t *f, const oparams_t *p, const double row[5], double mu, double mass, double charge)
{
    double tp[4] = { row[0], row[1], row[2], row[3] };
    int sp = gc_cycrad(f, row, mu, mass, charge) / field_lengthscale(f, tp) < p->epss;
    if (f->is_static) return sp;
    if (!sp) return 0;
    return gc_cycper(f, row, mu, mass, charge) / field_timescale(f, tp) < p->epst;
}

/* One GuidingCenter.advance(delta, eom) call.  state = (t, X,Y,Z, ppar).  dt = params["GCtimestep"]
 * or bounceperiod()/bounceresolution supplied by the caller (:443-446).  Stored rows: cols 0-4 = state,
 * col 5 = mu, col 7 = cumulative nstep.  Returns 1 normal, 3 NonAdiabatic raised, negative solver failure. */
static int gc_advance_one(const ofield_t *f, const oparams_t *p, int eom, double state[5], double mu, double v,
                          double mass, double charge, double dt, double delta, int check_adiab,
                          double *rows, long max_rows, long store_every, long *row_index, long *nstored,
                          long *percall, long max_calls, long *ncalls, ocount_t *cnt, double *tcur)
{
    gctx_t ctx = { f, p, mass, charge, mu, v, eom };
    double t0 = state[0], x = t0, y[4];
    memcpy(y, state + 1, 4 * sizeof(double));
    int ok = 1, ret = 1;
    while (ok && x < t0 + delta) {                                       /* :452 */
        ocount_t c1 = { 0, 0, 0, 0 };
        int idid = dopri5(4, gc_eom, &ctx, &x, y, x + dt, p->rtol, p->atol, &c1);   /* :453 */
        if (g_field_err) return ST_FIELD;      /* ValueError out of r.integrate(): no row appended */
        if (cnt) { cnt->nfcn += c1.nfcn; cnt->nstep += c1.nstep; cnt->naccpt += c1.naccpt; cnt->nrejct += c1.nrejct; }
        if (percall && *ncalls < max_calls) {
            long *q = percall + 4 * (*ncalls); q[0] = c1.nfcn; q[1] = c1.nstep; q[2] = c1.naccpt; q[3] = c1.nrejct;
        }
        if (ncalls) (*ncalls)++;
        if (idid < 0) { ok = 0; ret = idid; }
        state[0] = x; memcpy(state + 1, y, 4 * sizeof(double));           /* :454-455 */
        *tcur = x;                                                        /* :456 */
        (*row_index)++;
        if (rows && store_every > 0 && (*row_index % store_every) == 0 && *nstored < max_rows) {
            double *r = rows + 8 * (*nstored);
            memcpy(r, state, 5 * sizeof(double)); r[5] = mu; r[6] = 0; r[7] = cnt ? (double)cnt->nstep : 0;
            (*nstored)++;
        }
        if (check_adiab && !gc_isadiabatic(f, p, state, mu, mass, charge)) return 3;   /* :457-458 */
    }
    return ret;
}

/* GuidingCenter.__init__ :124-133: ppar and mu from (v, pa) ; pa == 90 -> vpar = 0 exactly (Q9) */
static void gc_construct(const ofield_t *f, double t0, const double pos[3], double v, double pa_deg, int use_pa,
                         double ppar_in, double mass, double *ppar, double *mu)
{
    double gamma = 1 / sqrt(1 - sq(v / C_LIGHT));
    double pp = ppar_in;
    if (use_pa) {
        double vpar = (pa_deg == 90) ? 0.0 : v * cos(pa_deg * M_PI / 180);
        pp = gamma * mass * vpar;
    }
    *mu = magnetic_moment(f, t0, pos, pp / (mass * gamma), v, mass);
    *ppar = pp;
}

/* GuidingCenter.init(Particle) :168-186 ; returns 0 or ST_GCITER */
static int switch_P2G(const ofield_t *f, const double prow[7], double mass, double charge,
                      double grow[5], double *mu, double *v)
{
    const double *mom = prow + 4;
    double gm = sqrt(sq(mass) + dot3(mom, mom) / (C_LIGHT * C_LIGHT));
    double vel[3] = { mom[0] / gm, mom[1] / gm, mom[2] / gm }, R[3], vp, spd;
    if (guidingcenter(f, prow[0], prow + 1, vel, mass, charge, R, &vp, &spd)) return ST_GCITER;
    double gamma = 1 / sqrt(1 - sq(spd / C_LIGHT));
    double pp;
    gc_construct(f, prow[0], R, spd, 0, 0, mass * gamma * vp, mass, &pp, mu);
    grow[0] = prow[0]; grow[1] = R[0]; grow[2] = R[1]; grow[3] = R[2]; grow[4] = pp;
    *v = spd;
    return 0;
}

/* Particle.init(GuidingCenter) :149-164 ; field evaluated at t_eval = the NEW Particle's tcur (Q12) */
static void switch_G2P(const ofield_t *f, const double grow[5], double mu, double mass, double charge,
                       double t_eval, double prow[7])
{
    double tp[4] = { grow[0], grow[1], grow[2], grow[3] };
    double B = field_magB(f, tp), v, pos[3], vel[3];
    double gammasq = 1 + 2 * mu * B / (mass * C_LIGHT * C_LIGHT) + sq(grow[4] / mass / C_LIGHT);
    if (sqrt(gammasq) - 1 < 1e-6) v = sqrt(2 * mu * B / mass + sq(grow[4] / mass));
    else v = C_LIGHT * sqrt(1 - 1 / gammasq);
    double vpar = grow[4] / mass / sqrt(gammasq);
    GCtoFP(f, t_eval, grow + 1, vpar, v, mass, charge, 0.0, pos, vel);
    /* Particle.__init__ :106-108 */
    double gamma = 1 / sqrt(1 - dot3(vel, vel) / (C_LIGHT * C_LIGHT));
    prow[0] = grow[0];
    for (int i = 0; i < 3; i++) { prow[1 + i] = pos[i]; prow[4 + i] = mass * gamma * vel[i]; }
}

/* --------------------------------------------------------------------------------------------
 * rkf.py:13-143 (RKF45, Burden & Faires) for d[s,x,y,z]/ds = +-[1, b]  (fieldline.py:39-52)
 * Appends accepted points to curve (cap x 4).  Returns number of points appended, or -1 when the
 * reference raises RuntimeError (h < hmin), in which case Fieldline.trace breaks (fieldline.py:67-68)
 * and the partial chunk is discarded.
 * ------------------------------------------------------------------------------------------ */
static void fl_deriv(const ofield_t *f, double time, const double Y[4], double sign, double out[4])
{
    double tp[4] = { time, Y[1], Y[2], Y[3] }, b[3];
    field_unitb(f, tp, b);
    out[0] = sign * 1.0; out[1] = sign * b[0]; out[2] = sign * b[1]; out[3] = sign * b[2];
}

static long rkf_chunk(const ofield_t *f, double time, double sign, const double x0[4], double b, double tol,
                      double hmax, double hmin, double *out, long cap)
{
    const double a2 = 2.500000000000000e-01, a3 = 3.750000000000000e-01, a4 = 9.230769230769231e-01,
                 a5 = 1.000000000000000e+00, a6 = 5.000000000000000e-01;
    const double b21 = 2.500000000000000e-01, b31 = 9.375000000000000e-02, b32 = 2.812500000000000e-01,
                 b41 = 8.793809740555303e-01, b42 = -3.277196176604461e+00, b43 = 3.320892125625853e+00,
                 b51 = 2.032407407407407e+00, b52 = -8.000000000000000e+00, b53 = 7.173489278752436e+00,
                 b54 = -2.058966861598441e-01, b61 = -2.962962962962963e-01, b62 = 2.000000000000000e+00,
                 b63 = -1.381676413255361e+00, b64 = 4.529727095516569e-01, b65 = -2.750000000000000e-01;
    const double r1 = 2.777777777777778e-03, r3 = -2.994152046783626e-02, r4 = -2.919989367357789e-02,
                 r5 = 2.000000000000000e-02, r6 = 3.636363636363636e-02;
    const double c1 = 1.157407407407407e-01, c3 = 5.489278752436647e-01, c4 = 5.353313840155945e-01,
                 c5 = -2.000000000000000e-01;
    (void)a2; (void)a3; (void)a4; (void)a5; (void)a6;   /* f does not depend on t */
    double t = 0, x[4], h = hmax, k1[4], k2[4], k3[4], k4[4], k5[4], k6[4], y[4], d[4];
    long n = 0;
    memcpy(x, x0, sizeof x);
    while (t < b) {
        if (t + h > b) h = b - t;
        int i;
        fl_deriv(f, time, x, sign, d); for (i = 0; i < 4; i++) k1[i] = h * d[i];
        for (i = 0; i < 4; i++) y[i] = x[i] + b21 * k1[i];
        fl_deriv(f, time, y, sign, d); for (i = 0; i < 4; i++) k2[i] = h * d[i];
        for (i = 0; i < 4; i++) y[i] = x[i] + b31 * k1[i] + b32 * k2[i];
        fl_deriv(f, time, y, sign, d); for (i = 0; i < 4; i++) k3[i] = h * d[i];
        for (i = 0; i < 4; i++) y[i] = x[i] + b41 * k1[i] + b42 * k2[i] + b43 * k3[i];
        fl_deriv(f, time, y, sign, d); for (i = 0; i < 4; i++) k4[i] = h * d[i];
        for (i = 0; i < 4; i++) y[i] = x[i] + b51 * k1[i] + b52 * k2[i] + b53 * k3[i] + b54 * k4[i];
        fl_deriv(f, time, y, sign, d); for (i = 0; i < 4; i++) k5[i] = h * d[i];
        for (i = 0; i < 4; i++) y[i] = x[i] + b61 * k1[i] + b62 * k2[i] + b63 * k3[i] + b64 * k4[i] + b65 * k5[i];
        fl_deriv(f, time, y, sign, d); for (i = 0; i < 4; i++) k6[i] = h * d[i];
        double r = 0;
        for (i = 0; i < 4; i++) {
            double ri = fabs(r1 * k1[i] + r3 * k3[i] + r4 * k4[i] + r5 * k5[i] + r6 * k6[i]) / h;
            if (ri > r) r = ri;
        }
        if (r <= tol) {
            t = t + h;
            for (i = 0; i < 4; i++) x[i] = x[i] + c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i];
            if (n < cap) memcpy(out + 4 * n, x, sizeof x);
            n++;
        }
        h = h * fmin(fmax(0.84 * pow(tol / r, 0.25), 0.1), 4.0);      /* r == 0 -> inf -> 4 (Q15) */
        if (h > hmax) h = hmax;
        else if (h < hmin) return -1;
    }
    return n;
}

/* fieldline.py:13-35 + 37-105 with Bmax = Bm, no stopcond, Bmin None (as halfbouncepath uses it).
 * curve: cap x 4 (s,x,y,z) ordered as self.curve; Bout: |B| per point (fieldline.py:123-127).
 * Returns the number of points (may exceed cap -> caller retries); *ds_out = step. */
long oracle_fieldline_trace(const ofield_t *f, const double tpos[4], double Bm, double fieldlineresolution,
                            double *curve, double *Bout, long cap, double *ds_out)
{
    double ds = 1 / field_curvature(f, tpos) / fieldlineresolution;   /* fieldline.py:31-35 */
    if (ds_out) *ds_out = ds;
    double *fw = (double *)malloc(sizeof(double) * 4 * cap), *bw = (double *)malloc(sizeof(double) * 4 * cap);
    long nf = 1, nb = 1;
    double init[4] = { 0, tpos[1], tpos[2], tpos[3] };
    memcpy(fw, init, sizeof init); memcpy(bw, init, sizeof init);
    for (int dir = 0; dir < 2; dir++) {
        double *arr = dir ? bw : fw; long *np_ = dir ? &nb : &nf;
        double sign = dir ? -1.0 : 1.0, tol = dir ? 1e-4 : 1e-3;        /* Q8 */
        for (;;) {
            if (*np_ >= cap) break;
            long m = rkf_chunk(f, tpos[0], sign, arr + 4 * (*np_ - 1), ds, tol, ds, 1e-6, arr + 4 * (*np_), cap - *np_);
            if (m < 0) break;
            if (*np_ + m > cap) { *np_ += m; break; }
            *np_ += m;
            double tp[4] = { tpos[0], arr[4 * (*np_ - 1) + 1], arr[4 * (*np_ - 1) + 2], arr[4 * (*np_ - 1) + 3] };
            if (field_magB(f, tp) > Bm) break;
        }
    }
    long n = (nb - 1) + nf;
    if (n <= cap && nb <= cap && nf <= cap) {
        long k = 0;
        for (long i = nb - 1; i >= 1; i--, k++) memcpy(curve + 4 * k, bw + 4 * i, 4 * sizeof(double));
        for (long i = 0; i < nf; i++, k++) memcpy(curve + 4 * k, fw + 4 * i, 4 * sizeof(double));
        for (long i = 0; i < n; i++) {
            double tp[4] = { tpos[0], curve[4 * i + 1], curve[4 * i + 2], curve[4 * i + 3] };
            Bout[i] = field_magB(f, tp);
        }
    } else n = cap + 1;
    free(fw); free(bw);
    return n;
}

/* GuidingCenter.bounceperiod :593-606 up to the call of flutils.bounceperiod: (Bmirror, v) */
void oracle_gc_mirror(const ofield_t *f, const double state[5], double mu, double mass, double *Bm, double *v)
{
    double tp[4] = { state[0], state[1], state[2], state[3] }, ppar = state[4];
    double Bmag = field_magB(f, tp);
    double gamma = sqrt(1 + 2 * mu * Bmag / (mass * C_LIGHT * C_LIGHT) + sq(ppar / (mass * C_LIGHT)));
    if (gamma - 1 < 1e-6) {
        double p = sqrt(2 * mass * mu * Bmag + sq(ppar));
        *v = p / mass; *Bm = sq(p) / (2 * mass * mu);
    } else {
        double p = mass * C_LIGHT * sqrt((gamma + 1) * (gamma - 1));
        *Bm = sq(p) / ((p - ppar) * (p + ppar)) * Bmag;
        *v = p / mass / gamma;
    }
}

/* --------------------------------------------------------------------------------------------
 * exported entry points (ctypes) -- ensembles are plain loops over independent particles
 * ------------------------------------------------------------------------------------------ */
/* Field operators at npt points (tests): out layout documented in oracle/oracle.py */
void oracle_field_ops(const ofield_t *f, long npt, const double *tpos, double *B, double *E, double *unitb,
                      double *magB, double *gradB, double *jac, double *curlb, double *curv, double *dBdt,
                      double *dbdt, double *lscale, double *tscale)
{
    for (long i = 0; i < npt; i++) {
        const double *tp = tpos + 4 * i; double J[3][3];
        field_B(f, tp, B + 3 * i); field_E(f, tp, E + 3 * i); field_unitb(f, tp, unitb + 3 * i);
        magB[i] = field_magB(f, tp); field_gradB(f, tp, gradB + 3 * i);
        field_jacobianB(f, tp, J); memcpy(jac + 9 * i, J, sizeof J);
        field_curlb(f, tp, curlb + 3 * i); curv[i] = field_curvature(f, tp);
        dBdt[i] = field_dBdt(f, tp); field_dbdt(f, tp, dbdt + 3 * i);
        lscale[i] = field_lengthscale(f, tp);
        tscale[i] = f->is_static ? NAN : field_timescale(f, tp);
    }
}

void oracle_utils(const ofield_t *f, long npt, const double *pos, const double *vel, double mass, double charge,
                  double *cycper, double *cycrad, double *gcR, double *gcvp, double *gcv, double *mu,
                  double *fppos, double *fpvel, double *cycper2, double *cycrad2)
{
    for (long i = 0; i < npt; i++) {
        const double *r = pos + 3 * i, *v = vel + 3 * i;
        cycper[i] = cyclotron_period(f, 0, r, v, mass, charge);
        cycrad[i] = cyclotron_radius(f, 0, r, v, mass, charge);
        double R[3] = { NAN, NAN, NAN }, vp = NAN, sp = NAN;
        guidingcenter(f, 0, r, v, mass, charge, R, &vp, &sp);
        memcpy(gcR + 3 * i, R, sizeof R); gcvp[i] = vp; gcv[i] = sp;
        mu[i] = magnetic_moment(f, 0, R, vp, sp, mass);
        GCtoFP(f, 0, R, vp, sp, mass, charge, 0, fppos + 3 * i, fpvel + 3 * i);
        cycper2[i] = cyclotron_period2(f, 0, R, sp, mass, charge);
        cycrad2[i] = cyclotron_radius2(f, 0, R, vp, sp, mass, charge);
    }
}

void oracle_getperp(const double v[3], double o[3]) { getperp(v, o); }

/* generic solver on built-in test ODEs (pins dop853/dopri5 against scipy's _dop directly):
 * ode 0: y'' = -y (1 + 5000 exp(-((t-1.5)/0.02)^2) + 3000 exp(-((t-2.2)/0.01)^2))   (forces rejections)
 * ode 1: y'' = -y (1 + 50 sin(3t)^2) */
static void test_ode(double t, const double *y, double *dy, void *ctx)
{
    int which = *(int *)ctx;
    dy[0] = y[1];
    if (which == 0) {
        double a = (t - 1.5) / 0.02, b = (t - 2.2) / 0.01;
        dy[1] = -y[0] * (1 + 5000 * exp(-(a * a)) + 3000 * exp(-(b * b)));
    } else {
        double s = sin(3 * t);
        dy[1] = -y[0] * (1 + 50 * (s * s));
    }
}
int oracle_test_solver(int which_solver, int which_ode, double x0, double xend, double *y, double rtol,
                       double atol, int reject_rule, long *counters)
{
    ocount_t c = { 0, 0, 0, 0 }; double x = x0; int idid;
    if (which_solver == 853) idid = dop853(2, test_ode, &which_ode, &x, y, xend, rtol, atol, reject_rule, &c);
    else idid = dopri5(2, test_ode, &which_ode, &x, y, xend, rtol, atol, &c);
    counters[0] = c.nfcn; counters[1] = c.nstep; counters[2] = c.naccpt; counters[3] = c.nrejct;
    return idid;
}

/* dopri5 on a caller-supplied right-hand side (BounceCenter.advance, BounceCenter.py:246-249: the RHS there is
 * scipy quadrature over field-line traces, which oracle.py evaluates with scipy itself, as the reference does) */
typedef void (*cb_rhs_fn)(double t, const double *y, double *dy);
typedef struct { cb_rhs_fn fn; } cb_ctx_t;
static void cb_rhs(double t, const double *y, double *dy, void *ctx) { ((cb_ctx_t *)ctx)->fn(t, y, dy); }
int oracle_dopri5_callback(int n, cb_rhs_fn fn, double *x, double *y, double xend, double rtol, double atol, long *counters)
{
    ocount_t c = { 0, 0, 0, 0 }; cb_ctx_t ctx = { fn };
    g_field_err = 0;
    int idid = dopri5(n, cb_rhs, &ctx, x, y, xend, rtol, atol, &c);
    counters[0] += c.nfcn; counters[1] += c.nstep; counters[2] += c.naccpt; counters[3] += c.nrejct;
    return idid;
}

/* Particle ensemble.  In/out SoA state t,x,y,z,px,py,pz (n each); mass, charge per particle.
 * rows: n x max_rows x 8 (row 0 of each particle = initial state) or NULL.
 * Outputs per particle: nrows (1 + output intervals), nstored, counters[4], status, tcur, dt.
 * percall: only honoured for n == 1 (capacity max_calls x 4). */
void oracle_particle_advance(const ofield_t *f, const oparams_t *p, long n,
                             double *t, double *x, double *y, double *z, double *px, double *py, double *pz,
                             const double *mass, const double *charge, double delta, int check_adiab,
                             long store_every, long max_rows, double *rows, long *nrows, long *nstored,
                             long *counters, int *status, double *tcur, double *dt,
                             long *percall, long max_calls, int nthreads)
{
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (long i = 0; i < n; i++) {
        double st[7] = { t[i], x[i], y[i], z[i], px[i], py[i], pz[i] };
        ocount_t c = { 0, 0, 0, 0 };
        g_field_err = 0; tl_errgap = 1e300;
        long ri = 0, ns = 0, nc = 0;
        double *r = rows ? rows + (size_t)i * max_rows * 8 : NULL;
        if (r && store_every > 0 && max_rows > 0) { memcpy(r, st, sizeof st); r[7] = 0; ns = 1; }
        status[i] = particle_advance_one(f, p, st, mass[i], charge[i], delta, check_adiab, r, max_rows, store_every,
                                         &ri, &ns, n == 1 ? percall : NULL, max_calls, &nc, &c, &tcur[i], &dt[i]);
        t[i] = st[0]; x[i] = st[1]; y[i] = st[2]; z[i] = st[3]; px[i] = st[4]; py[i] = st[5]; pz[i] = st[6];
        if (g_errgap_out) g_errgap_out[i] = tl_errgap;
        nrows[i] = ri + 1; nstored[i] = ns;
        counters[4 * i] = c.nfcn; counters[4 * i + 1] = c.nstep; counters[4 * i + 2] = c.naccpt; counters[4 * i + 3] = c.nrejct;
    }
}

/* GuidingCenter.__init__ for an ensemble: ppar, mu from (v, pa) */
void oracle_gc_construct(const ofield_t *f, long n, const double *t0, const double *x, const double *y, const double *z,
                         const double *v, const double *pa, const double *mass, double *ppar, double *mu)
{
    for (long i = 0; i < n; i++) {
        double pos[3] = { x[i], y[i], z[i] };
        gc_construct(f, t0[i], pos, v[i], pa[i], 1, 0, mass[i], &ppar[i], &mu[i]);
    }
}

/* GuidingCenter ensemble.  State t,X,Y,Z,ppar in/out; mu, v, mass, charge, dt per particle. */
void oracle_gc_advance(const ofield_t *f, const oparams_t *p, int eom, long n,
                       double *t, double *x, double *y, double *z, double *ppar,
                       const double *mu, const double *v, const double *mass, const double *charge, const double *dt,
                       double delta, int check_adiab, long store_every, long max_rows, double *rows,
                       long *nrows, long *nstored, long *counters, int *status, double *tcur,
                       long *percall, long max_calls, int nthreads)
{
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (long i = 0; i < n; i++) {
        double st[5] = { t[i], x[i], y[i], z[i], ppar[i] };
        ocount_t c = { 0, 0, 0, 0 };
        g_field_err = 0; tl_errgap = 1e300;
        long ri = 0, ns = 0, nc = 0;
        double *r = rows ? rows + (size_t)i * max_rows * 8 : NULL;
        if (r && store_every > 0 && max_rows > 0) { memcpy(r, st, sizeof st); r[5] = mu[i]; r[6] = 0; r[7] = 0; ns = 1; }
        status[i] = gc_advance_one(f, p, eom, st, mu[i], v[i], mass[i], charge[i], dt[i], delta, check_adiab,
                                   r, max_rows, store_every, &ri, &ns, n == 1 ? percall : NULL, max_calls, &nc, &c, &tcur[i]);
        t[i] = st[0]; x[i] = st[1]; y[i] = st[2]; z[i] = st[3]; ppar[i] = st[4];
        if (g_errgap_out) g_errgap_out[i] = tl_errgap;
        nrows[i] = ri + 1; nstored[i] = ns;
        counters[4 * i] = c.nfcn; counters[4 * i + 1] = c.nstep; counters[4 * i + 2] = c.naccpt; counters[4 * i + 3] = c.nrejct;
    }
}

/* Adaptive: __init__ (Adaptive.py:96-104) + advance (Adaptive.py:202-222) for one tracer.
 * Requires params.gctimestep != 0 (bounce-period dt needs the host quadrature; see oracle.py).
 * rows: max_rows x 8, every row of every segment (store_every = 1); col 7 holds
 * seg_index * 2 + mode for the row.  seglog: max_segs x 3 doubles (mode, first row, tcur at creation).
 * Returns number of segments (negative: failure code). */
long oracle_adaptive_one(const ofield_t *f, const oparams_t *p, const double pos[3], const double vel[3],
                         double t0, double mass, double charge, double delta,
                         double *rows, long max_rows, long *nrows_out, double *seglog, long max_segs,
                         long *counters)
{
    /* Particle.__init__ :106-108 */
    double gamma = 1 / sqrt(1 - dot3(vel, vel) / (C_LIGHT * C_LIGHT));
    double prow[7] = { t0, pos[0], pos[1], pos[2], mass * gamma * vel[0], mass * gamma * vel[1], mass * gamma * vel[2] };
    double grow[5], mu = 0, v = 0;
    int mode;
    long nseg = 0, nr = 0;
    ocount_t c = { 0, 0, 0, 0 };
    if (particle_isadiabatic(f, p, prow, mass, charge)) {       /* Adaptive.py:98-102 */
        if (switch_P2G(f, prow, mass, charge, grow, &mu, &v)) return ST_GCITER;
        mode = MODE_GC;
    } else mode = MODE_PARTICLE;
#define PUSH_ROW() do { if (nr < max_rows) { double *r = rows + 8 * nr; memset(r, 0, 64); \
        if (mode == MODE_PARTICLE) memcpy(r, prow, 56); else { memcpy(r, grow, 40); r[5] = mu; } \
        r[7] = (double)((nseg - 1) * 2 + mode); } nr++; } while (0)
#define PUSH_SEG(tc) do { if (nseg < max_segs) { seglog[3 * nseg] = mode; seglog[3 * nseg + 1] = (double)nr; seglog[3 * nseg + 2] = (tc); } nseg++; } while (0)
    PUSH_SEG(t0); PUSH_ROW();
    double t = 0, tcur = t0;
    while (t < delta) {                                          /* Adaptive.py:205 */
        int ret;
        double rem = delta - t;
        /* run the current segment row by row so every row is logged */
        if (mode == MODE_PARTICLE) {
            double tstart = prow[0];
            /* single-row stepping would recompute dt; instead call advance_one with a private row sink */
            long ri = 0, ns = 0, ncalls = 0;
            double *sink = rows + 8 * nr; long cap = max_rows - nr; if (cap < 0) cap = 0;
            double dtp;
            ret = particle_advance_one(f, p, prow, mass, charge, rem, 1, sink, cap, 1, &ri, &ns, NULL, 0, &ncalls, &c, &tcur, &dtp);
            for (long k = 0; k < ns; k++) sink[8 * k + 7] = (double)((nseg - 1) * 2 + mode);
            nr += ri; (void)tstart;
            if (ret == 2) {                                       /* Adiabatic -> GuidingCenter (Adaptive.py:215-221) */
                if (switch_P2G(f, prow, mass, charge, grow, &mu, &v)) return ST_GCITER;
                mode = MODE_GC; tcur = grow[0];
                PUSH_SEG(tcur); PUSH_ROW();
            } else if (ret < 0) break;
        } else {
            if (p->gctimestep == 0) return -100;
            long ri = 0, ns = 0, ncalls = 0;
            double *sink = rows + 8 * nr; long cap = max_rows - nr; if (cap < 0) cap = 0;
            ret = gc_advance_one(f, p, EOM_TAOCHANBRIZARD, grow, mu, v, mass, charge, p->gctimestep, rem, 1,
                                 sink, cap, 1, &ri, &ns, NULL, 0, &ncalls, &c, &tcur);
            for (long k = 0; k < ns; k++) sink[8 * k + 7] = (double)((nseg - 1) * 2 + mode);
            nr += ri;
            if (ret == 3) {                                       /* NonAdiabatic -> Particle (Adaptive.py:208-214) */
                switch_G2P(f, grow, mu, mass, charge, 0.0 /* Particle().tcur, Q12 */, prow);
                mode = MODE_PARTICLE; tcur = prow[0];
                PUSH_SEG(tcur); PUSH_ROW();
            } else if (ret < 0) break;
        }
        t = tcur;                                                /* Adaptive.py:222 (Q14) */
    }
    *nrows_out = nr;
    counters[0] = c.nfcn; counters[1] = c.nstep; counters[2] = c.naccpt; counters[3] = c.nrejct;
    return nseg;
}

/* switch transforms exposed for unit tests */
int oracle_switch_P2G(const ofield_t *f, const double prow[7], double mass, double charge, double grow[5], double *mu, double *v)
{ return switch_P2G(f, prow, mass, charge, grow, mu, v); }
void oracle_switch_G2P(const ofield_t *f, const double grow[5], double mu, double mass, double charge, double t_eval, double prow[7])
{ switch_G2P(f, grow, mu, mass, charge, t_eval, prow); }
int oracle_particle_isadiabatic(const ofield_t *f, const oparams_t *p, const double row[7], double mass, double charge)
{ return particle_isadiabatic(f, p, row, mass, charge); }
int oracle_gc_isadiabatic(const ofield_t *f, const oparams_t *p, const double row[5], double mu, double mass, double charge)
{ return gc_isadiabatic(f, p, row, mu, mass, charge); }

/* right-hand sides exposed for unit tests */
void oracle_gc_rhs(const ofield_t *f, const oparams_t *p, int eom, double mass, double charge, double mu, double v,
                   double t, const double *Y, double *out)
{
    gctx_t ctx = { f, p, mass, charge, mu, v, eom };
    gc_eom(t, Y, out, &ctx);
}
void oracle_particle_rhs(const ofield_t *f, const oparams_t *p, double mass, double charge, double gm,
                         double t, const double *Y, double *out)
{
    pctx_t ctx = { f, p, mass, charge, gm };
    particle_eom(t, Y, out, &ctx);
}

/* debug: one dopri5 call on the GC equations with every RHS call logged (t, Y[4], out[4]) */
static double *g_log; static long g_nlog, g_caplog;
static void gc_eom_logged(double t, const double *Y, double *out, void *vctx)
{
    gc_eom(t, Y, out, vctx);
    if (g_nlog < g_caplog) { double *r = g_log + 9 * g_nlog; r[0] = t; memcpy(r + 1, Y, 32); memcpy(r + 5, out, 32); }
    g_nlog++;
}
long oracle_gc_debug_row(const ofield_t *f, const oparams_t *p, int eom, double mass, double charge, double mu, double v,
                         double x0, double *y, double xend, double *log, long cap)
{
    gctx_t ctx = { f, p, mass, charge, mu, v, eom };
    g_log = log; g_nlog = 0; g_caplog = cap;
    double x = x0;
    dopri5(4, gc_eom_logged, &ctx, &x, y, xend, p->rtol, p->atol, NULL);
    return g_nlog;
}
