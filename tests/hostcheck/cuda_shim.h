// TEST INFRASTRUCTURE ONLY.  Just enough of the CUDA device vocabulary for g++ to compile the kernel headers of
// rapt_b200/csrc/ (one "thread", one "block") so that tests/test_kernel_host.py can run the *product's own kernel
// source* against the oracle without a GPU.  Nothing in rapt_b200/ includes this file.
#pragma once
#include <cmath>
#include <cstring>
#include <cstddef>
#include <algorithm>

#define RAPT_HOST_BUILD 1
#define __device__
#define __host__
#define __global__
#define __constant__
#define __shared__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __restrict__ __restrict

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
struct hc_dim3 { unsigned x, y, z; };
static const hc_dim3 blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, gridDim = {1, 1, 1};

static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
static inline int __double2loint(double d) { return (int)(__double_as_longlong(d) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo)
{
    return __longlong_as_double(((long long)hi << 32) | (long long)(unsigned)lo);
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }   // a "warp" of one lane
static inline int __all_sync(unsigned, int pred) { return pred ? 1 : 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __syncthreads() {}
static inline int __syncthreads_and(int pred) { return pred ? 1 : 0; }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }   // work queue
using std::min; using std::max;
using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::fma; using std::rint;
