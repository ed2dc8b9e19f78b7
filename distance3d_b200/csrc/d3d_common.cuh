// Error reporting and launch helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// Stores the message for d3d_last_error_string() and returns -1.
int d3d_set_error(const char *fmt, ...);
// Number of SMs of the current device (cached per device).
int d3d_sm_count();

#define D3D_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t err__ = (expr);                                                       \
        if (err__ != cudaSuccess)                                                         \
            return d3d_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,    \
                                 cudaGetErrorString(err__));                              \
    } while (0)

static __host__ __device__ inline int64_t d3d_min64(int64_t a, int64_t b) { return a < b ? a : b; }
