// F10 — multi-scale Laplacian-of-Gaussian "blobness" of the 2-D path
// (nellie/segmentation/filtering.py:772-795 _filter_log, :927-930).
//
// scipy.ndimage.gaussian_laplace(frame, (s, s)) = T0 + T1 with T_a the separable Gaussian of
// derivative order 2 along axis a and order 0 along the other (each a pair of nb200_gauss_axis passes
// with the taps of scipy's _gaussian_kernel1d, truncate 4.0) and one float32 add.
// Per sigma: cur = (-(T0+T1)) * fl32(s^2) * mask; L = running max (first sigma initialises);
// after the loop L[L<0] = 0; blob = (L / (max(L) + 1e-12)) / 10; V = max(V, max(blob, 0)).
#include "common.cuh"
#include "devmath.cuh"

namespace {

__global__ void __launch_bounds__(256)
log_accumulate_kernel(const float* __restrict__ t0, const float* __restrict__ t1, const float* __restrict__ acc,
                      float sigma_sq, int first, long long n, float* __restrict__ L) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float lap = t0[i] + t1[i];
        const float m = acc[i] >= 0.0f ? 1.0f : 0.0f;      // AND-mask of the sigma loop (acc == -1 is dead)
        const float cur = ((-lap) * sigma_sq) * m;
        if (first) L[i] = cur;
        else if (cur > L[i]) L[i] = cur;
    }
}

__global__ void __launch_bounds__(256)
log_max_kernel(const float* __restrict__ L, long long n, long long* __restrict__ max_bits) {
    float m = 0.0f;     // values below 0 are clamped to 0 first (filtering.py:792), so max >= 0
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, L[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)max_bits, (unsigned long long)nb::f2u(m));
}

__global__ void __launch_bounds__(256)
log_combine_kernel(const float* __restrict__ acc, const float* __restrict__ L, const long long* __restrict__ max_bits,
                   long long n, float* __restrict__ v) {
    const float top = nb::u2f((uint32_t)(*max_bits));
    const float denom = top + 1e-12f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float l = L[i];
        if (l < 0.0f) l = 0.0f;
        float blob = (l / denom) / 10.0f;
        blob = fmaxf(blob, 0.0f);
        const float a = acc[i];
        const float ves = a > 0.0f ? a : 0.0f;               // vesselness * masks (filtering.py:926)
        v[i] = fmaxf(ves, blob);
    }
}

__global__ void reset_word_kernel(long long* w) {
    if (threadIdx.x == 0) *w = 0;
}

}  // namespace

extern "C" {

int nb200_log2d_accumulate(const float* t0, const float* t1, const float* acc, float sigma_sq, int first,
                           long long n, float* L, void* stream) {
    NB_REQUIRE(t0 && t1 && acc && L && n > 0, NB200_ERR_ARG, "nb200_log2d_accumulate: bad argument");
    log_accumulate_kernel<<<nb::grid_for(n, 256, 8), 256, 0, nb::as_stream(stream)>>>(t0, t1, acc, sigma_sq, first, n, L);
    return nb::check_launch("log2d_accumulate");
}

int nb200_log2d_combine(const float* acc, const float* L, long long n, long long* max_bits, float* v, void* stream) {
    NB_REQUIRE(acc && L && max_bits && v && n > 0, NB200_ERR_ARG, "nb200_log2d_combine: bad argument");
    cudaStream_t st = nb::as_stream(stream);
    reset_word_kernel<<<1, 32, 0, st>>>(max_bits);
    log_max_kernel<<<nb::grid_for(n, 256, 4), 256, 0, st>>>(L, n, max_bits);
    log_combine_kernel<<<nb::grid_for(n, 256, 8), 256, 0, st>>>(acc, L, max_bits, n, v);
    return nb::check_launch("log2d_combine");
}

}  // extern "C"
