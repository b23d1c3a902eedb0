// rapt_launch.h -- host-side launch entry points exported by each kernels_tu.cu flavour.
#pragma once
#include <cuda_runtime.h>
#include "rapt_types.h"

#define RAPT_DECLARE_FLAVOUR(NS)                                                          \
    namespace NS {                                                                        \
    cudaError_t launch_particle(const rapt::AdvArgs &a, int grid, cudaStream_t s);        \
    int particle_blocks_per_sm(int rkn);                                                       \
    cudaError_t launch_gc(const rapt::AdvArgs &a, int grid, cudaStream_t s);              \
    int gc_blocks_per_sm();                                                               \
    cudaError_t launch_particle_dt(const rapt::AdvArgs &a, double *key, int *idx, cudaStream_t s); \
    cudaError_t launch_field_ops(const void *args, cudaStream_t s);                       \
    cudaError_t launch_misc(const void *args, cudaStream_t s);                            \
    cudaError_t launch_bounce(const void *args, cudaStream_t s);                          \
    cudaError_t launch_adaptive_switch(const void *args, cudaStream_t s);                 \
    cudaError_t launch_bounce_center(const void *args, int grid, cudaStream_t s);         \
    }
RAPT_DECLARE_FLAVOUR(rapt_fast)
RAPT_DECLARE_FLAVOUR(rapt_strict)
