// Serial host emulation of the CUDA launch model for the barrier-free per-voxel kernels (TEST INFRASTRUCTURE ONLY).
//
// A .cu file whose kernels use neither shared memory nor barriers (atomics only through nb_atomic_*) launches through the NB_LAUNCH macro; when it is
// compiled with  g++ -DNB200_HOST_EMU='"<path>/cuda_emu.h"' -x c++  this header replaces common.cuh: __global__ functions become
// ordinary functions, NB_LAUNCH runs them once per (block, thread) with blockIdx / threadIdx set, one after the other.  The
// kernel bodies, the index arithmetic, the grid-stride loops and the extern "C" entry points (argument checks included) are
// then the very code nvcc compiles, so `pytest -m "not gpu"` can check them against scipy and the executed-reference fixtures
// without a GPU.  Never loaded by nellie_b200.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../include/nellie_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
};
static emu_dim3 gridDim, blockDim, blockIdx, threadIdx;
typedef void* cudaStream_t;

namespace nb {
static char g_err[512];
inline void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
inline int check_launch(const char*) { return NB200_OK; }
inline cudaStream_t as_stream(void* s) { return s; }
// a small "device" so that grid-stride loops really stride: 4 SMs
inline unsigned grid_for(long long work_items, int threads, int ctas_per_sm) {
    long long need = (work_items + threads - 1) / threads;
    long long cap = 4ll * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}
}  // namespace nb

#define NB_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            nb::set_error(__VA_ARGS__);        \
            return (code);                     \
        }                                      \
    } while (0)

#define NB_LAUNCH(kernel, grid, block, stream, ...)                            \
    do {                                                                       \
        (void)(stream);                                                        \
        gridDim.x = (grid);                                                    \
        blockDim.x = (block);                                                  \
        for (unsigned nb_b = 0; nb_b < gridDim.x; ++nb_b)                      \
            for (unsigned nb_t = 0; nb_t < blockDim.x; ++nb_t) {               \
                blockIdx.x = nb_b;                                             \
                threadIdx.x = nb_t;                                            \
                kernel(__VA_ARGS__);                                           \
            }                                                                  \
    } while (0)

// the only atomics the emulated kernels use; execution is serial
#define nb_atomic_min_u32(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define nb_atomic_max_i32(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#define nb_atomic_min_i32(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define nb_atomic_max_u64(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#define nb_atomic_add_u64(p, v) do { *(p) += (v); } while (0)

extern "C" const char* nb200_emu_last_error(void) { return nb::g_err; }
