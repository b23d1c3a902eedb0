// rapt_fields.cuh -- analytic field models as inlined device functions + the finite-difference
// operator layer of rapt/fields.py:_Field.  Compiled by nvcc (built-ins) and by NVRTC (user
// snippets), so no host headers here.
//
// Two arithmetic flavours, selected at compile time:
//   RAPT_STRICT=1 (built with -fmad=false): mirrors the CPU reference's unfused fp64 operation
//                 order (including numpy's fused 3-vector dot), for parity runs;
//   RAPT_STRICT=0 : FMA contraction allowed, pow(r2,2.5) -> rsqrt chain, reciprocal multiplies.
#pragma once

#ifndef RAPT_STRICT
#define RAPT_STRICT 0
#endif
#ifndef RAPT_NS
#define RAPT_NS rapt_fast
#endif
#ifndef RAPT_DIPOLE_SERIES
#define RAPT_DIPOLE_SERIES 1     /* 1: c r^-5 of the scaled dipole straight from the MUFU seed (see Field<0>::Bs) */
#endif

#define RAPT_C_LIGHT 299792458.0           /* rapt/__init__.py:8  */
#define RAPT_EARTH_B0 3.07e-5              /* rapt/__init__.py:9  */
#define RAPT_EARTH_RE 6378137.0            /* rapt/__init__.py:10 */
#define RAPT_PI 3.141592653589793
#define RAPT_NAN __longlong_as_double(0x7ff8000000000000LL)

#define RAPT_DEV __device__ __forceinline__

// Python's scalar `x**2` is glibc pow(x, 2.0), 0.52 ulp and not always x*x ('oracle').  The device has no
// glibc, so the strict flavour squares by multiplication there; the host build of the same source
// (tests/hostcheck/kernel_host.cpp, bit-compared with the reference's trajectories) calls pow as the reference does.
#if defined(RAPT_HOST_BUILD) && RAPT_STRICT
#define RAPT_SQ(x) pow((x), 2.0)
#else
#define RAPT_SQ(x) ((x) * (x))
#endif

#include "rapt_types.h"

namespace RAPT_NS {

using rapt::FieldP;
using rapt::GridP;

// np.dot on 3-vectors as executed by the reference's numpy: fma(a2,b2, fma(a1,b1, a0*b0))
RAPT_DEV double dot3(double ax, double ay, double az, double bx, double by, double bz)
{
    return fma(az, bz, fma(ay, by, ax * bx));
}

RAPT_DEV double sgn(double z) { return z > 0 ? 1.0 : (z < 0 ? -1.0 : 0.0); }

// Branch-free reciprocal and reciprocal square root for the fast flavour: MUFU seed (~2^-22) refined
// to 1-2 ulp with fused Newton steps.  No denormal / special-case slow paths -- every argument on the
// hot path (r^2, |B|^2, error scales, gamma*m) is a normal, strictly positive number.
#ifdef RAPT_HOST_BUILD
// tests/hostcheck/kernel_host.cpp compiles these headers for the CPU (test infrastructure, never loaded by the
// product): the MUFU seeds become the exact value truncated to 23 mantissa bits, the same accuracy class.
RAPT_DEV double mufu_seed(double v) { return __longlong_as_double(__double_as_longlong(v) & ~0x1fffffffLL); }
RAPT_DEV double mufu_rcp(double x) { return mufu_seed(1.0 / x); }
RAPT_DEV double mufu_rsqrt(double x) { return mufu_seed(1.0 / sqrt(x)); }
RAPT_DEV unsigned grid_dynamic_smem() { return 0; }
#else
RAPT_DEV double mufu_rcp(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
RAPT_DEV double mufu_rsqrt(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
RAPT_DEV unsigned grid_dynamic_smem()
{
    unsigned n;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(n));
    return n;
}
#endif
RAPT_DEV double fast_rcp(double x)
{
    double y = mufu_rcp(x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
// log and exp for the step-size controllers: branch-free, ~3e-14 relative, positive normal arguments
// (log) and |y| < 700 (exp).  The library versions are ~100 instructions each with slow paths.
RAPT_DEV double fast_log(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    double m = __hiloint2double(hi, lo);                    // [1, 2)
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }       // [sqrt(1/2), sqrt(2))
    const double s = (m - 1.0) * fast_rcp(m + 1.0), s2 = s * s;
    double p = 1.0 / 15.0;
    p = fma(p, s2, 1.0 / 13.0); p = fma(p, s2, 1.0 / 11.0); p = fma(p, s2, 1.0 / 9.0); p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0); p = fma(p, s2, 1.0 / 3.0); p = fma(p, s2, 1.0);
    return fma((double)e, 0.6931471805599453, 2.0 * s * p);
}
RAPT_DEV double fast_exp(double y)
{
    const double n = rint(y * 1.4426950408889634);
    const double f = fma(-n, 0.6931471805599453, y);        // |f| <= 0.3466
    double p = 1.0 / 39916800.0;
    p = fma(p, f, 1.0 / 3628800.0); p = fma(p, f, 1.0 / 362880.0); p = fma(p, f, 1.0 / 40320.0); p = fma(p, f, 1.0 / 5040.0);
    p = fma(p, f, 1.0 / 720.0); p = fma(p, f, 1.0 / 120.0); p = fma(p, f, 1.0 / 24.0); p = fma(p, f, 1.0 / 6.0);
    p = fma(p, f, 0.5); p = fma(p, f, 1.0); p = fma(p, f, 1.0);
    return __hiloint2double(__double2hiint(p) + (int)n * (1 << 20), __double2loint(p));
}
// one Newton step: ~2^-45; enough for the error-norm scale factors 1/(atol + rtol |y|)
RAPT_DEV double fast_rcp1(double x)
{
    double y = mufu_rcp(x);
    return fma(y, fma(-x, y, 1.0), y);
}
RAPT_DEV double fast_rsqrt(double x)
{
    double y = mufu_rsqrt(x);
    double e = fma(-(x * y), y, 1.0);                 // 1 - x y^2  (~2^-22)
#ifdef RAPT_RSQRT_POLISH
    y = fma(y * e, fma(0.375, e, 0.5), y);
    e = fma(-(x * y), y, 1.0);                        // optional linear polish: last-bit accuracy
    return fma(0.5 * y, e, y);
#else
    // y (1 + e/2 + 3e^2/8): cubic convergence, 2^-22 -> ~2^-66, i.e. ~2 ulp after rounding.  The extra
    // polish step (4 more FP64 instructions per field evaluation) bought nothing in the parity tests and
    // cost 6 % of the config-2 run time (profiles/r1_particle_history.md).
    return fma(y * e, fma(0.375, e, 0.5), y);
#endif
}

// c / r^5 from r2 = r^2 in one go from the 22-bit MUFU seed y0 ~ 1/r: with e = 1 - r2 y0^2 (|e| < 2^-21),
// r^-5 = y0^5 (1 - e)^(-5/2) = y0^5 (1 + e (5/2 + 35/8 e)) + O(e^3 ~ 1e-19).  5 DMUL + 3 DFMA; the rsqrt + powers form is 6 + 3.
RAPT_DEV double fast_c_over_r5(double c, double r2)
{
    const double y0 = mufu_rsqrt(r2), u = y0 * y0;
    const double e = fma(-r2, u, 1.0);
    const double y5 = (c * y0) * (u * u);
    return fma(y5 * e, fma(4.375, e, 2.5), y5);
}

#ifdef RAPT_USER_FIELD
// supplied by the NVRTC-compiled user snippet
__device__ void rapt_user_B(double t, double x, double y, double z, const double *prm, double *B);
#if RAPT_USER_HAS_E
__device__ void rapt_user_E(double t, double x, double y, double z, const double *prm, double *E);
#endif
#endif

// ------------------------------------------------------------------------------------------
// fields.Grid (fields.py:513-814): multilinear interpolation of gridded E/B in (t, x, y, z), or (x, y, z)
// for a single time point -- scipy's RegularGridInterpolator(method="linear") as Grid.Bgrid/Egrid call it:
// interval search per axis (x[i] <= v < x[i+1], closed on the right at the last node), normalised distance
// (v - x[i]) / (x[i+1] - x[i]), then the 2^d cell vertices in itertools.product order (first axis slowest)
// with weight ((w0 w1) w2) w3 and value += node * weight.  The strict flavour executes exactly those
// operations (bit-identical to scipy on the same tables).  Outside the grid the reference raises
// ValueError; here the field is NaN and the advance kernels stop the tracer with RAPT_ST_FIELD.
// The reference keeps a rolling window of three time points on the host (fields.py:697-705); linear
// interpolation between the two bracketing time points does not depend on the window, so all time points
// stay resident in HBM and the window is the bracketing pair.
// ------------------------------------------------------------------------------------------
// FieldP::prm[0] carries the device address of the grid's GridP block (read as a value: taking the address of a
// kernel parameter would force a local copy of the whole argument struct)
RAPT_DEV const GridP *grid_of(const FieldP &f) { return reinterpret_cast<const GridP *>(__double_as_longlong(f.prm[0])); }
struct GridVec { double x, y, z; };

RAPT_DEV bool grid_locate(const double *__restrict__ g, int n, double g0, double ginv, double v, int &idx, double &w)
{
#if !RAPT_STRICT
    if (ginv != 0.0) {
        // fast flavour, uniform axis: index and normalised distance from the arithmetic node positions
        // g0 + k/ginv -- no table look-ups, no division; differs from the stored nodes by their rounding
        // (~1e-13 of a cell)
        const double u = (v - g0) * ginv;
        if (!(u >= 0.0 && u <= (double)(n - 1))) return false;
        const int k0 = min((int)u, n - 2);
        idx = k0; w = u - (double)k0;
        return true;
    }
#endif
    if (!(__ldg(g) <= v && v <= __ldg(g + n - 1))) return false;
    int k;
    if (ginv != 0.0) {                               // uniform axis: direct index, then make it exact
        k = (int)((v - g0) * ginv);
        k = max(0, min(k, n - 2));
        while (k > 0 && v < __ldg(g + k)) k--;
        while (k < n - 2 && v >= __ldg(g + k + 1)) k++;
    } else {                                         // binary search
        int low = 0, high = n - 2;
        if (v == __ldg(g + n - 1)) low = high;
        while (low < high) {
            const int mid = (high + low) >> 1;
            if (v < __ldg(g + mid)) high = mid;
            else if (v >= __ldg(g + mid + 1)) low = mid + 1;
            else { low = mid; break; }
        }
        k = low;
    }
    const double a = __ldg(g + k), b = __ldg(g + k + 1);
    w = (v - a) / (b - a);
    idx = k;
    return true;
}

// one shared copy per kernel (the unrolled particle kernel has 13 call sites: inlined, its hot loop no longer
// fits the instruction cache -- profiles/r1_grid_field.md)
// Per-thread cell cache in dynamic shared memory (advance kernels of a gridded field are launched with
// blockDim * 8 * (1 + 24 * min(nt, 2)) bytes): slot 0 = index of the cached cell, then the 8 (16 with time
// interpolation) vertices of that cell, 3 components each, entry k of thread i at word k * blockDim + i
// (conflict-free).  Consecutive evaluations of one tracer -- the RK stages of a step, the 7 stencil points
// of the guiding-centre right-hand side -- almost always fall in the same cell, so the 16 gathers of 32-byte
// sectors through L1/L2 happen once per cell instead of once per evaluation (profiles/r1_grid_field.md).
RAPT_DEV void grid_cache_reset()
{
    extern __shared__ double rapt_grid_cache[];
    if (grid_dynamic_smem() >= blockDim.x * 8u) rapt_grid_cache[threadIdx.x] = __longlong_as_double(-1LL);
}

static __device__ __noinline__ GridVec grid_eval(const GridP *__restrict__ gp, int which, double t, double x, double y, double z)
{
    extern __shared__ double rapt_grid_cache[];
    const GridP g = *gp;                             // uniform loads, L1-resident
    const double *__restrict__ tab = which ? g.E : g.B;
    GridVec r;
    int it = 0, ix, iy, iz;
    double wt = 0, wx, wy, wz;
    bool ok = grid_locate(g.x, g.nx, g.x0, g.xinv, x, ix, wx);
    ok = grid_locate(g.y, g.ny, g.y0, g.yinv, y, iy, wy) && ok;
    ok = grid_locate(g.z, g.nz, g.z0, g.zinv, z, iz, wz) && ok;
    const int ntp = (g.nt >= 2) ? 2 : 1;
    if (ntp == 2) ok = grid_locate(g.t, g.nt, 0.0, 0.0, t, it, wt) && ok;
    if (!ok) { r.x = r.y = r.z = RAPT_NAN; return r; }
    const size_t sy = (size_t)g.nz, sx = sy * g.ny, st = sx * g.nx;
    const size_t cell = (it * st + ix * sx) + iy * sy + iz;
    const double2 *__restrict__ node = reinterpret_cast<const double2 *>(tab) + 2 * cell;
    const unsigned nthr = blockDim.x;
    const bool cached = (which == 0) && grid_dynamic_smem() >= nthr * 8u * (1u + 24u * ntp);
    double *my = rapt_grid_cache + threadIdx.x;
    if (cached && __double_as_longlong(my[0]) != (long long)cell) {
        // refill: vertex pairs in evaluation order, 6 doubles each (x, y, z of the lower and the upper z-node)
        int k = 1;
        for (int a = 0; a < ntp; a++)
            for (int b = 0; b < 2; b++)
                for (int c = 0; c < 2; c++) {
                    const double2 *q = node + 2 * (a * st + b * sx + c * sy);
                    const double2 lxy = __ldg(q), lz = __ldg(q + 1), hxy = __ldg(q + 2), hz = __ldg(q + 3);
                    my[(k + 0) * nthr] = lxy.x; my[(k + 1) * nthr] = lxy.y; my[(k + 2) * nthr] = lz.x;
                    my[(k + 3) * nthr] = hxy.x; my[(k + 4) * nthr] = hxy.y; my[(k + 5) * nthr] = hz.x;
                    k += 6;
                }
        my[0] = __longlong_as_double((long long)cell);
    }
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int k = 1;
#pragma unroll
    for (int a = 0; a < 2; a++) {
        if (a >= ntp) break;
        const double ft = (ntp == 2) ? (a ? wt : 1 - wt) : 1.0;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const double fx = (ntp == 2) ? ft * (b ? wx : 1 - wx) : (b ? wx : 1 - wx);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double fy = fx * (c ? wy : 1 - wy);
                double lx, ly, lzz, hx, hy, hzz;
                if (cached) {
                    lx = my[(k + 0) * nthr]; ly = my[(k + 1) * nthr]; lzz = my[(k + 2) * nthr];
                    hx = my[(k + 3) * nthr]; hy = my[(k + 4) * nthr]; hzz = my[(k + 5) * nthr];
                    k += 6;
                } else {
                    const double2 *q = node + 2 * (a * st + b * sx + c * sy);      // two z-neighbours: 64 contiguous bytes
                    const double2 lxy = __ldg(q), lz = __ldg(q + 1), hxy = __ldg(q + 2), hz = __ldg(q + 3);
                    lx = lxy.x; ly = lxy.y; lzz = lz.x; hx = hxy.x; hy = hxy.y; hzz = hz.x;
                }
                const double w0 = fy * (1 - wz), w1 = fy * wz;
                v0 = v0 + lx * w0; v1 = v1 + ly * w0; v2 = v2 + lzz * w0;
                v0 = v0 + hx * w1; v1 = v1 + hy * w1; v2 = v2 + hzz * w1;
            }
        }
    }
    r.x = v0; r.y = v1; r.z = v2;
    return r;
}

// ------------------------------------------------------------------------------------------
// Field models.  KIND is a compile-time constant so every kernel contains exactly one model.
// ------------------------------------------------------------------------------------------
template <int KIND> struct Field {
    static constexpr bool HAS_E =
#ifdef RAPT_USER_FIELD
        (KIND == 100) ? (RAPT_USER_HAS_E != 0) :
#endif
        (KIND == 3) || (KIND == 6);
    static constexpr bool TIME_DEP = (KIND == 4) || (KIND == 6) || (KIND == 100);
    static constexpr bool UNIFORM = (KIND == 2) || (KIND == 3);
    static constexpr bool CAN_FAIL = (KIND == 6);      // gridded data: NaN outside the grid
    // B(t, x) = Bspace(tfactor(t), x): lets the guiding-centre stencil evaluate the time factor once per
    // right-hand side instead of once per stencil point (fast flavour; VarEarthDipole, fields.py:469-470)
    static constexpr bool SEPARABLE = (KIND == 4) && !RAPT_STRICT;

    // tfactor: the time factor times the dipole coefficient -B0 Re^3, so that the stencil points pay no multiply for it
    static RAPT_DEV double tfactor(const FieldP &f, double t)
    {
        return (-RAPT_EARTH_B0 * (RAPT_EARTH_RE * RAPT_EARTH_RE * RAPT_EARTH_RE)) * (1 + f.prm[0] * sin(2 * RAPT_PI * t / f.prm[1]));
    }
    static RAPT_DEV void Bspace(const FieldP &f, double tf, double x, double y, double z, double &bx, double &by, double &bz)
    {
#if RAPT_DIPOLE_SERIES
        const double w = fast_c_over_r5(tf, x * x + y * y + z * z);
#else
        const double ir = fast_rsqrt(x * x + y * y + z * z), ir2 = ir * ir;
        const double w = tf * (ir2 * ir2 * ir);
#endif
        const double tz = 3 * z;
        bx = w * (tz * x); by = w * (tz * y); bz = w * fma(2 * z, z, -fma(x, x, y * y));
    }

    static RAPT_DEV void B(const FieldP &f, double t, double x, double y, double z,
                           double &bx, double &by, double &bz)
    {
        if (KIND == 0) {                    // EarthDipole, fields.py:315-317
            double r2 = x * x + y * y + z * z;
#if RAPT_STRICT
            double s = f.prm[0] / pow(r2, 2.5);
            bx = s * (x * z); by = s * (y * z); bz = s * (z * z - r2 / 3);
#else
#if RAPT_DIPOLE_SERIES
            double s = fast_c_over_r5(f.prm[0], r2);
#else
            double ir = fast_rsqrt(r2), ir2 = ir * ir;
            double s = f.prm[0] * (ir2 * ir2 * ir);
#endif
            bx = s * (x * z); by = s * (y * z); bz = s * (z * z - r2 * (1.0 / 3.0));
#endif
        } else if (KIND == 1) {             // DoubleDipole, fields.py:358-362
            double x2 = x - f.prm[1], k = f.prm[2];
#if RAPT_STRICT
            double p1 = pow(x * x + y * y + z * z, 5.0 / 2.0);
            double a0 = 3 * x * z / p1, a1 = 3 * y * z / p1, a2 = (2 * z * z - x * x - y * y) / p1;
            double p2 = pow(x2 * x2 + y * y + z * z, 5.0 / 2.0);
            double b0 = k * (3 * x2 * z) / p2, b1 = k * (3 * y * z) / p2, b2 = k * (2 * z * z - x2 * x2 - y * y) / p2;
            bx = f.prm[0] * (a0 + b0); by = f.prm[0] * (a1 + b1); bz = f.prm[0] * (a2 + b2);
#else
            double yz2 = y * y + z * z, zz2 = 2 * z * z - y * y;
            // B0 folded into the two weights (B0 and B0 k are invariants of the launch): 2 multiplies less per evaluation
#if RAPT_DIPOLE_SERIES
            double w1 = fast_c_over_r5(f.prm[0], x * x + yz2), w2 = fast_c_over_r5(f.prm[0] * k, x2 * x2 + yz2);
#else
            double r1 = fast_rsqrt(x * x + yz2), r2_ = fast_rsqrt(x2 * x2 + yz2);
            double q1 = r1 * r1, q2 = r2_ * r2_;
            double w1 = (f.prm[0] * r1) * (q1 * q1), w2 = ((f.prm[0] * k) * r2_) * (q2 * q2);
#endif
            double tz = 3 * z;
            bx = tz * (x * w1 + x2 * w2);
            by = (tz * y) * (w1 + w2);
            bz = (zz2 - x * x) * w1 + (zz2 - x2 * x2) * w2;
#endif
        } else if (KIND == 2 || KIND == 3) { // UniformBz / UniformCrossedEB, fields.py:390
            bx = 0; by = 0; bz = f.prm[0];
        } else if (KIND == 4) {             // VarEarthDipole, fields.py:469-470
            double s = -RAPT_EARTH_B0 * (RAPT_EARTH_RE * RAPT_EARTH_RE * RAPT_EARTH_RE) *
                       (1 + f.prm[0] * sin(2 * RAPT_PI * t / f.prm[1]));
#if RAPT_STRICT
            double p = pow(x * x + y * y + z * z, 5.0 / 2.0);
            bx = s * (3 * x * z) / p; by = s * (3 * y * z) / p; bz = s * (2 * z * z - x * x - y * y) / p;
#else
            double ir = fast_rsqrt(x * x + y * y + z * z), ir2 = ir * ir;
            double w = s * (ir2 * ir2 * ir);
            bx = w * (3 * x * z); by = w * (3 * y * z); bz = w * (2 * z * z - x * x - y * y);
#endif
        } else if (KIND == 5) {             // Parabolic, fields.py:506-511 (quirk Q6: module B0 outside |z|<=1)
            if (fabs(z) <= 1.0) bx = f.prm[0] * z / f.prm[2];
            else bx = sgn(z) * RAPT_EARTH_B0;
            by = 0; bz = f.prm[1];
        } else if (KIND == 6) {             // Grid.B -> Grid.Bgrid, fields.py:707-741, 774-794
            const GridVec r = grid_eval(grid_of(f), 0, t, x, y, z);
            bx = r.x; by = r.y; bz = r.z;
        }
#ifdef RAPT_USER_FIELD
        else if (KIND == 100) {
            double o[3];
            rapt_user_B(t, x, y, z, f.prm, o);
            bx = o[0]; by = o[1]; bz = o[2];
        }
#endif
        else { bx = by = bz = 0; }
    }

    // sc * B(t, x): the Lorentz kernels need q/(gamma m) * B; folding the factor into the dipole coefficient
    // saves three multiplies per evaluation (fast flavour only)
    static RAPT_DEV void Bs(const FieldP &f, double sc, double t, double x, double y, double z,
                            double &bx, double &by, double &bz)
    {
#if !RAPT_STRICT
        if (KIND == 0) {
            const double r2 = x * x + y * y + z * z;
#if RAPT_DIPOLE_SERIES
            const double w = fast_c_over_r5(sc * f.prm[0], r2);
#else
            const double ir = fast_rsqrt(r2), ir2 = ir * ir;
            const double w = (sc * f.prm[0] * ir) * (ir2 * ir2);
#endif
            const double wz = w * z;
            bx = wz * x; by = wz * y; bz = w * fma(z, z, -(1.0 / 3.0) * r2);
            return;
        }
        if (KIND == 5) {                    // Parabolic: B0/d is an invariant of the launch (no division per evaluation)
            bx = (fabs(z) <= 1.0) ? (sc * (f.prm[0] / f.prm[2])) * z : sc * (sgn(z) * RAPT_EARTH_B0);
            by = 0; bz = sc * f.prm[1];
            return;
        }
#endif
        B(f, t, x, y, z, bx, by, bz);
        bx *= sc; by *= sc; bz *= sc;
    }

    static RAPT_DEV void E(const FieldP &f, double t, double x, double y, double z,
                           double &ex, double &ey, double &ez)
    {
        ex = 0; ey = 0; ez = 0;             // fields.py:59-74
        if (KIND == 3) ey = f.prm[1];       // fields.py:427
        if (KIND == 6) {                    // Grid.E -> Grid.Egrid, fields.py:743-772, 796-814
            if (f.prm[1] != 0.0) {          // the grid has an electric field table
                const GridVec r = grid_eval(grid_of(f), 1, t, x, y, z);
                ex = r.x; ey = r.y; ez = r.z;
            }
        }
#if defined(RAPT_USER_FIELD) && RAPT_USER_HAS_E
        if (KIND == 100) {
            double o[3];
            rapt_user_E(t, x, y, z, f.prm, o);
            ex = o[0]; ey = o[1]; ez = o[2];
        }
#endif
    }

    // fields.py:108-109
    static RAPT_DEV double magB(const FieldP &f, double t, double x, double y, double z)
    {
        double bx, by, bz; B(f, t, x, y, z, bx, by, bz);
        return sqrt(dot3(bx, by, bz, bx, by, bz));
    }
    // fields.py:90-91
    static RAPT_DEV void unitb(const FieldP &f, double t, double x, double y, double z,
                               double &ux, double &uy, double &uz)
    {
        double bx, by, bz; B(f, t, x, y, z, bx, by, bz);
#if RAPT_STRICT
        double m = sqrt(dot3(bx, by, bz, bx, by, bz));
        ux = bx / m; uy = by / m; uz = bz / m;
#else
        double im = fast_rsqrt(dot3(bx, by, bz, bx, by, bz));
        ux = bx * im; uy = by * im; uz = bz * im;
#endif
    }
    // fields.py:125-131 : central differences with the nominal 2*d (perturb-then-round as the reference)
    static RAPT_DEV void gradB(const FieldP &f, double t, double x, double y, double z,
                               double &gx, double &gy, double &gz)
    {
        if (UNIFORM) { gx = gy = gz = 0; return; }      // exact: identical field values subtract to 0
        double d = f.gradstep;
#if RAPT_STRICT
        gx = (magB(f, t, x + d, y, z) - magB(f, t, x - d, y, z)) / (2 * d);
        gy = (magB(f, t, x, y + d, z) - magB(f, t, x, y - d, z)) / (2 * d);
        gz = (magB(f, t, x, y, z + d) - magB(f, t, x, y, z - d)) / (2 * d);
#else
        double i2d = 1.0 / (2 * d);
        gx = (magB(f, t, x + d, y, z) - magB(f, t, x - d, y, z)) * i2d;
        gy = (magB(f, t, x, y + d, z) - magB(f, t, x, y - d, z)) * i2d;
        gz = (magB(f, t, x, y, z + d) - magB(f, t, x, y, z - d)) * i2d;
#endif
    }
    // fields.py:191-200 with _M1 (fields.py:33-36); strict mode keeps the summation order the
    // reference's BLAS uses for np.dot(_M1, beta)
    static RAPT_DEV void curlb(const FieldP &f, double t, double x, double y, double z,
                               double &cx, double &cy, double &cz)
    {
        if (UNIFORM) { cx = cy = cz = 0; return; }
        double d = f.gradstep;
        double pxx, pxy, pxz, mxx, mxy, mxz, pyx, pyy, pyz, myx, myy, myz, pzx, pzy, pzz, mzx, mzy, mzz;
        unitb(f, t, x + d, y, z, pxx, pxy, pxz); unitb(f, t, x - d, y, z, mxx, mxy, mxz);   // beta[0..2], [3..5]
        unitb(f, t, x, y + d, z, pyx, pyy, pyz); unitb(f, t, x, y - d, z, myx, myy, myz);   // beta[6..8], [9..11]
        unitb(f, t, x, y, z + d, pzx, pzy, pzz); unitb(f, t, x, y, z - d, mzx, mzy, mzz);   // beta[12..14], [15..17]
        (void)pxx; (void)mxx; (void)pyy; (void)myy; (void)pzz; (void)mzz;
#if RAPT_STRICT
        cx = ((pyz + (-myz + -pzy)) + mzy) / (2 * d);          // +b8 -b11 -b13 +b16
        cy = ((-pxz + pzx) + (mxz + -mzx)) / (2 * d);          // -b2 +b5 +b12 -b15
        cz = ((pxy + myx) + (-mxy + -pyx)) / (2 * d);          // +b1 -b4 -b6 +b9
#else
        double i2d = 1.0 / (2 * d);
        cx = ((pyz - myz) - (pzy - mzy)) * i2d;
        cy = ((pzx - mzx) - (pxz - mxz)) * i2d;
        cz = ((pxy - mxy) - (pyx - myx)) * i2d;
#endif
    }
    // fields.py:239-245
    static RAPT_DEV void dbdt(const FieldP &f, double t, double x, double y, double z,
                              double &ox, double &oy, double &oz)
    {
        double d = f.tstep, ax, ay, az, bx, by, bz;
        unitb(f, t - d, x, y, z, ax, ay, az);
        unitb(f, t + d, x, y, z, bx, by, bz);
        ox = (bx - ax) / d / 2; oy = (by - ay) / d / 2; oz = (bz - az) / d / 2;
    }
    // fields.py:217-223
    static RAPT_DEV double dBdt(const FieldP &f, double t, double x, double y, double z)
    {
        double d = f.tstep;
        return (magB(f, t + d, x, y, z) - magB(f, t - d, x, y, z)) / d / 2;
    }
    // fields.py:148-153 : J[i][j] = dB_i/dx_j
    static RAPT_DEV void jacobianB(const FieldP &f, double t, double x, double y, double z, double J[9])
    {
        double d = f.gradstep, ax, ay, az, bx, by, bz;
        B(f, t, x + d, y, z, ax, ay, az); B(f, t, x - d, y, z, bx, by, bz);
        J[0] = (ax - bx) / (2 * d); J[3] = (ay - by) / (2 * d); J[6] = (az - bz) / (2 * d);
        B(f, t, x, y + d, z, ax, ay, az); B(f, t, x, y - d, z, bx, by, bz);
        J[1] = (ax - bx) / (2 * d); J[4] = (ay - by) / (2 * d); J[7] = (az - bz) / (2 * d);
        B(f, t, x, y, z + d, ax, ay, az); B(f, t, x, y, z - d, bx, by, bz);
        J[2] = (ax - bx) / (2 * d); J[5] = (ay - by) / (2 * d); J[8] = (az - bz) / (2 * d);
    }
    // fields.py:261
    static RAPT_DEV double lengthscale(const FieldP &f, double t, double x, double y, double z)
    {
        double J[9], m = 0;
        jacobianB(f, t, x, y, z, J);
#pragma unroll
        for (int i = 0; i < 9; i++) m = fmax(m, fabs(J[i]));
        return magB(f, t, x, y, z) / m;
    }
    // fields.py:277-280
    // max_ij |B_i(x + d e_j) - B_i(x - d e_j)| = 2 d max|J_ij|: the denominator of lengthscale() without its nine divisions
    // (fast-flavour adiabaticity predicates compare products instead of quotients)
    static RAPT_DEV double max_central_difference(const FieldP &f, double t, double x, double y, double z)
    {
        if (UNIFORM) return 0.0;
        const double d = f.gradstep;
        double ax, ay, az, bx, by, bz, m;
        B(f, t, x + d, y, z, ax, ay, az); B(f, t, x - d, y, z, bx, by, bz);
        m = fmax(fmax(fabs(ax - bx), fabs(ay - by)), fabs(az - bz));
        B(f, t, x, y + d, z, ax, ay, az); B(f, t, x, y - d, z, bx, by, bz);
        m = fmax(m, fmax(fmax(fabs(ax - bx), fabs(ay - by)), fabs(az - bz)));
        B(f, t, x, y, z + d, ax, ay, az); B(f, t, x, y, z - d, bx, by, bz);
        return fmax(m, fmax(fmax(fabs(ax - bx), fabs(ay - by)), fabs(az - bz)));
    }
    static RAPT_DEV double timescale(const FieldP &f, double t, double x, double y, double z)
    {
        return magB(f, t, x, y, z) / fabs(dBdt(f, t, x, y, z));
    }
    // fields.py:170-174 (quirk Q7: np.dot(gB, B) with the scalar B is element-wise)
    static RAPT_DEV double curvature(const FieldP &f, double t, double x, double y, double z)
    {
        double bx, by, bz, gx, gy, gz;
        B(f, t, x, y, z, bx, by, bz);
        double Bm = sqrt(dot3(bx, by, bz, bx, by, bz));
        gradB(f, t, x, y, z, gx, gy, gz);
        double px = gx - ((gx * Bm) / (Bm * Bm)) * bx;
        double py = gy - ((gy * Bm) / (Bm * Bm)) * by;
        double pz = gz - ((gz * Bm) / (Bm * Bm)) * bz;
        return sqrt(dot3(px, py, pz, px, py, pz)) / Bm;
    }
};

}  // namespace RAPT_NS
