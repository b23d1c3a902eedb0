// kernels_tu.cu -- one translation unit per (arithmetic flavour, kernel family); see Makefile.
//   -DRAPT_STRICT=0|1 -DRAPT_NS=rapt_fast|rapt_strict -DRAPT_TU_PARTICLE | -DRAPT_TU_GC | -DRAPT_TU_AUX | -DRAPT_TU_BC
#include <cuda_runtime.h>
#include <cstdlib>
#include "rapt_launch.h"

// gridded field: dynamic shared memory for the per-thread cell cache (prm[2] = number of time points)
template <class K> static size_t grid_cache_bytes(const rapt::AdvArgs &a, K kernel, int threads)
{
    if (a.f.kind != 6) return 0;
    const size_t bytes = (size_t)threads * 8 * (1 + 24 * (a.f.prm[2] >= 2.0 ? 2 : 1));
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return bytes;
}

// Block shape of the persistent advance kernels.  A launch that fills the GPU runs ONE big block per SM: the warps of one block
// share the SM's issue slots evenly, the warps of four 128-thread blocks do not (profiles/r2_tail.md: -10 % on config 2).
// A launch with fewer tracers than lanes (the late epochs of an adaptive run, single objects) is latency-bound and wants
// its warps spread over as many SMs as possible: 128-thread blocks.  Both shapes are launches of the same kernel.
static int block_threads(int grid128, int big)
{
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    return ((long long)grid128 * 128 >= (long long)sms * big) ? big : 128;
}

#ifdef RAPT_TU_PARTICLE
#include "rapt_particle.cuh"
#if !RAPT_STRICT
#include "rapt_particle_rkn.cuh"
#endif
namespace RAPT_NS {
// fast flavour, static field, no equatorial constraint -> the Nystrom-form kernel (16 warps/SM)
#if !RAPT_STRICT
static bool use_rkn(const rapt::AdvArgs &a) { return a.f.is_static && !a.p.enforce_equatorial && !getenv("RAPT_B200_NO_RKN"); }
#endif
template <int KIND> static cudaError_t go_particle(const rapt::AdvArgs &a, int grid, cudaStream_t s)
{
#if !RAPT_STRICT
    if (use_rkn(a)) {
        const int T = block_threads(grid, RAPT_RKN_THREADS);
        k_particle_rkn<Field<KIND>><<<(grid * 128 + T - 1) / T, T, grid_cache_bytes(a, k_particle_rkn<Field<KIND>>, T), s>>>(a);
    } else
#endif
    k_particle_dop853<Field<KIND>><<<grid, 128, grid_cache_bytes(a, k_particle_dop853<Field<KIND>>, 128), s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_particle(const rapt::AdvArgs &a, int grid, cudaStream_t s)
{
    switch (a.f.kind) {
    case 0: return go_particle<0>(a, grid, s);
    case 1: return go_particle<1>(a, grid, s);
    case 2: return go_particle<2>(a, grid, s);
    case 3: return go_particle<3>(a, grid, s);
    case 4: return go_particle<4>(a, grid, s);
    case 5: return go_particle<5>(a, grid, s);
    case 6: return go_particle<6>(a, grid, s);
    default: return cudaErrorInvalidValue;
    }
}
int particle_blocks_per_sm(int rkn)
{
    int nb = 0;
#if !RAPT_STRICT
    if (rkn) {   // in units of 128 lanes, rounded up: surplus blocks of the persistent kernel find the queue empty
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_particle_rkn<Field<0>>, RAPT_RKN_THREADS, 0);
        return (nb * RAPT_RKN_THREADS + 127) / 128;
    }
#endif
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_particle_dop853<Field<0>>, 128, 0);
    return nb;
}
}  // namespace RAPT_NS
#endif

#ifdef RAPT_TU_GC
#include "rapt_gc.cuh"
#ifndef RAPT_GC_DEFAULT_BLOCKS
#define RAPT_GC_DEFAULT_BLOCKS 4
#endif
namespace RAPT_NS {
static int gc_minb() { const char *e = getenv("RAPT_B200_GC_BLOCKS"); return e ? atoi(e) : RAPT_GC_DEFAULT_BLOCKS; }
template <int KIND> static cudaError_t go_gc(const rapt::AdvArgs &a, int grid, cudaStream_t s)
{
    // `grid` counts 128-lane units (capi.cu:grid_for); the kernel runs RAPT_GC_THREADS threads per block
    const int T = block_threads(grid, RAPT_GC_THREADS), nb = (grid * 128 + T - 1) / T;
    if (gc_minb() >= 4) k_gc_dopri5<Field<KIND>, 4><<<nb, T, grid_cache_bytes(a, k_gc_dopri5<Field<KIND>, 4>, T), s>>>(a);
    else if (gc_minb() == 3) k_gc_dopri5<Field<KIND>, 3><<<nb, T, grid_cache_bytes(a, k_gc_dopri5<Field<KIND>, 3>, T), s>>>(a);
    else k_gc_dopri5<Field<KIND>, 2><<<nb, T, grid_cache_bytes(a, k_gc_dopri5<Field<KIND>, 2>, T), s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_gc(const rapt::AdvArgs &a, int grid, cudaStream_t s)
{
    switch (a.f.kind) {
    case 0: return go_gc<0>(a, grid, s);
    case 1: return go_gc<1>(a, grid, s);
    case 2: return go_gc<2>(a, grid, s);
    case 3: return go_gc<3>(a, grid, s);
    case 4: return go_gc<4>(a, grid, s);
    case 5: return go_gc<5>(a, grid, s);
    case 6: return go_gc<6>(a, grid, s);
    default: return cudaErrorInvalidValue;
    }
}
int gc_blocks_per_sm()
{
    int nb = 0;
    const int T = RAPT_GC_THREADS;
    if (gc_minb() >= 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gc_dopri5<Field<1>, 4>, T, 0);
    else if (gc_minb() == 3) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gc_dopri5<Field<1>, 3>, T, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gc_dopri5<Field<1>, 2>, T, 0);
    return (nb * T + 127) / 128;      // in units of 128 lanes
}
}  // namespace RAPT_NS
#endif

#ifdef RAPT_TU_AUX
#include "rapt_aux.cuh"
namespace RAPT_NS {
// per-particle output step of Particle.advance (Particle.py:282): the work-ordering key (small dt = many rows)
template <class F>
__global__ void k_particle_dt(const rapt::AdvArgs a, double *key, int *idx)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.nwork) return;
    double mass = a.mass[i], q = a.charge[i], px = a.s4[i], py = a.s5[i], pz = a.s6[i];
    double gm = sqrt(mass * mass + dot3(px, py, pz, px, py, pz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    double vx = px / gm, vy = py / gm, vz = pz / gm;
    double gamma = 1.0 / sqrt(1 - dot3(vx, vy, vz, vx, vy, vz) / (RAPT_C_LIGHT * RAPT_C_LIGHT));
    double bx, by, bz;
    F::B(a.f, a.t[i], a.s1[i], a.s2[i], a.s3[i], bx, by, bz);
    double B2 = dot3(bx, by, bz, bx, by, bz), Bm = sqrt(B2);
    double dt = 2 * RAPT_PI * gamma * mass / Bm / fabs(q) / a.p.cyclotronresolution;
    double delta = a.delta_arr ? a.delta_arr[i] : a.delta;
    // predicted work ~ rows x steps per row.  rows = delta / dt; a row takes one step where the orbit stays near the field
    // strength dt was computed for and ~ B_mirror / B = 1 / sin^2(pitch angle) more where the tracer mirrors in a stronger
    // field: measured on config 2 (profiles/r2_work_order.md) steps/row ~ max(1, 0.9 / sin(alpha)), which brings the
    // longest-first schedule from 5.1 % to 0.6 % above the ideal makespan.  Scheduling only: results do not depend on it.
    double pB = dot3(px, py, pz, bx, by, bz), p2 = dot3(px, py, pz, px, py, pz);
    double sa = sqrt(fmax(1.0 - pB * pB / (p2 * B2), 1e-4));
    key[i] = dt / delta * fmin(1.0, sa / 0.9);
    // sort_by_work == 2 (device-resident ensembles advanced in several calls): `counters` still holds the previous call's
    // (nfcn, nstep, naccpt, nrejct) of this tracer; its step count replaces the estimate (same unit: 1 / steps).  Zero =
    // no history.  Scheduling only.
    if (a.p.sort_by_work == 2 && a.counters && !a.append) {
        const int prev = a.counters[4 * i + 1];
        if (prev > 0) key[i] = 1.0 / (double)prev;
    }
    idx[i] = (int)i;
}
#define RAPT_KIND_SWITCH(CALL)                         \
    switch (kind) {                                    \
    case 0: CALL(0); break; case 1: CALL(1); break;    \
    case 2: CALL(2); break; case 3: CALL(3); break;    \
    case 4: CALL(4); break; case 5: CALL(5); break;    \
    case 6: CALL(6); break;                            \
    default: return cudaErrorInvalidValue;             \
    }
cudaError_t launch_particle_dt(const rapt::AdvArgs &a, double *key, int *idx, cudaStream_t s)
{
    const int kind = a.f.kind;
    const int grid = (int)((a.nwork + 255) / 256);
#define CALL(K) k_particle_dt<Field<K>><<<grid, 256, 0, s>>>(a, key, idx)
    RAPT_KIND_SWITCH(CALL)
#undef CALL
    return cudaGetLastError();
}
cudaError_t launch_field_ops(const void *args, cudaStream_t s)
{
    const OpsArgs &a = *static_cast<const OpsArgs *>(args);
    const int kind = a.f.kind;
    const int grid = (int)((a.n + 127) / 128);
#define CALL(K) k_field_ops<Field<K>><<<grid, 128, 0, s>>>(a)
    RAPT_KIND_SWITCH(CALL)
#undef CALL
    return cudaGetLastError();
}
cudaError_t launch_misc(const void *args, cudaStream_t s)
{
    const MiscArgs &a = *static_cast<const MiscArgs *>(args);
    const int kind = a.f.kind;
    const int grid = (int)((a.n + 127) / 128);
#define CALL(K) k_misc<Field<K>><<<grid, 128, 0, s>>>(a)
    RAPT_KIND_SWITCH(CALL)
#undef CALL
    return cudaGetLastError();
}
cudaError_t launch_adaptive_switch(const void *args, cudaStream_t s)
{
    const AdaptArgs &a = *static_cast<const AdaptArgs *>(args);
    const int kind = a.f.kind;
    const int grid = (int)((a.n + 255) / 256);
#define CALL(K) k_adaptive_switch<Field<K>><<<grid, 256, 0, s>>>(a)
    RAPT_KIND_SWITCH(CALL)
#undef CALL
    return cudaGetLastError();
}
cudaError_t launch_bounce(const void *args, cudaStream_t s)
{
    const BounceArgs &a = *static_cast<const BounceArgs *>(args);
    const int kind = a.f.kind;
    const int grid = (int)((a.n + 127) / 128);
#define CALL(K) k_bounce_setup<Field<K>><<<grid, 128, 0, s>>>(a)
    RAPT_KIND_SWITCH(CALL)
#undef CALL
    return cudaGetLastError();
}
}  // namespace RAPT_NS
#endif

#ifdef RAPT_TU_BC
#include "rapt_bc.cuh"
namespace RAPT_NS {
cudaError_t launch_bounce_center(const void *args, int grid, cudaStream_t s)
{
    const BCArgs &a = *static_cast<const BCArgs *>(args);
    const int kind = a.f.kind;
    switch (kind) {
#define CALL(K) case K: k_bounce_center<Field<K>><<<grid, 64, 0, s>>>(a); break
    CALL(0); CALL(1); CALL(2); CALL(3); CALL(4); CALL(5);
#undef CALL
    default: return cudaErrorInvalidValue;      // gridded fields: not offered (per-lane cell cache + curve scratch)
    }
    return cudaGetLastError();
}
}  // namespace RAPT_NS
#endif
