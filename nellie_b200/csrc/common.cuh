// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nellie_b200.h"

namespace nb {

// thread-local last error text, returned by nb200_last_error()
char* last_error_buf();
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? NB200_ERR_OOM : NB200_ERR_CUDA;
    }
    return NB200_OK;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// grid for a grid-stride kernel: a whole number of waves over the SMs
inline unsigned grid_for(long long work_items, int threads, int ctas_per_sm) {
    long long need = (work_items + threads - 1) / threads;
    long long cap = (long long)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// order-preserving map float -> uint32 (handles negatives), for atomicMin/atomicMax
__host__ __device__ inline uint32_t float_to_ordered(float f) {
#if defined(__CUDA_ARCH__)
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

}  // namespace nb

#define NB_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            nb::set_error(__VA_ARGS__);        \
            return (code);                     \
        }                                      \
    } while (0)
