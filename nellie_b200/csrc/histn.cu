// Label thresholds with a bin count other than 256 (labelling.py:23-35 `histogram_nbins`, :440-465; gpu_functions.py:23-94).
//
// The production histogram kernels (thresholds.cu) are specialised for the reference's default of 256 bins: shared-memory
// counters, a fixed state record that the Z-sharded Filter all-gathers, block-cooperative finalisation.  A caller who sets
// `histogram_nbins` to something else gets this separate, general path instead of an error; the 256-bin path and everything
// the Filter does is untouched.  Same numpy semantics, restated for any bin count:
//   * edges = np.linspace(min, max, nbins + 1) in the sample type (float32 for float32 samples; float64 for the intensity
//     Otsu of an integer frame, where np.histogram promotes the bins to float64), min == max widened by 0.5 on both sides;
//   * bin index = ((v - first) / (last - first)) * nbins truncated, the last edge closed, then numpy's +-1 correction against
//     the edges themselves;
//   * Otsu: centres, normalised counts, forward and reverse cumulative sums IN INDEX ORDER in float64 (np.cumsum), class
//     means, between-class variance, first maximum (first NaN if any);
//   * triangle: peak, first / last non-empty bin, flip towards the longer tail, distances to the peak-to-end line, first
//     maximum.
// min / max / count come from the existing nb200_hist_reset + nb200_hist_minmax (state words 0..2).  Bin counts are global
// 64-bit atomics; the finalisation is one thread walking nbins entries (a few microseconds per thousand bins) with its
// float64 work arrays in a caller-provided buffer.  Barrier-free, so the file also compiles for the host (oracle/cuda_emu.h).
#ifdef NB200_HOST_EMU
#include NB200_HOST_EMU
#else
#include "common.cuh"
#define NB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define nb_atomic_add_u64(p, v) atomicAdd((p), (v))
#endif
#include "devmath.cuh"

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ float ordered_to_f32(unsigned u) {
    return nb::u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// the value the histogram is taken of: arr[arr > 0], optionally log10 (numpy's float32 log10, devmath.cuh)
__device__ __forceinline__ bool kept_value(float raw, int log_domain, float& out) {
    if (!(raw > 0.0f)) return false;
    out = log_domain ? nb::np_log10f(raw) : raw;
    return true;
}

template <typename R>
__global__ void __launch_bounds__(THREADS)
histn_edges_kernel(const long long* __restrict__ state, int nbins, R* __restrict__ edges,
                   unsigned long long* __restrict__ counts) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += gridDim.x * blockDim.x) counts[i] = 0ull;
    if (state[NB200_HIST_COUNT] == 0) return;
    R first = (R)ordered_to_f32((unsigned)state[NB200_HIST_MIN]);
    R last = (R)ordered_to_f32((unsigned)state[NB200_HIST_MAX]);
    if (first == last) { first = first - (R)0.5; last = last + (R)0.5; }      // numpy _get_outer_edges
    const R delta = last - first;
    const R step = delta / (R)nbins;                                          // np.linspace
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= nbins; i += gridDim.x * blockDim.x) {
        R y = (R)i;
        if (step == (R)0) { y = y / (R)nbins; y = y * delta; }
        else y = y * step;
        y = y + first;
        if (i == nbins) y = last;
        edges[i] = y;
    }
}

template <typename R>
__global__ void __launch_bounds__(THREADS)
histn_bins_kernel(const float* __restrict__ vals, long long n, int log_domain, const long long* __restrict__ state,
                  int nbins, const R* __restrict__ edges, unsigned long long* counts) {
    if (state[NB200_HIST_COUNT] == 0) return;
    const R e_first = edges[0], e_last = edges[nbins];
    const R denom = e_last - e_first;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if (!kept_value(vals[i], log_domain, v)) continue;
        const R t = (R)v;
        if (!(t >= e_first) || !(t <= e_last)) continue;
        const R f = ((t - e_first) / denom) * (R)nbins;
        int b = (int)f;
        if (b == nbins) b = nbins - 1;
        if (t < edges[b]) --b;
        else if (t >= edges[b + 1] && b != nbins - 1) ++b;
        nb_atomic_add_u64(&counts[b], 1ull);
    }
}

// work: float64[6 * nbins] = p, pc, w_lo, s_lo, w_hi, s_hi.  out (float64[7]) as nb200_finalize_label_threshold:
// [0] threshold, [1] 10**tri or tri, [2] 10**otsu or otsu, [3] 1 = no samples, [4] status, [5] tri, [6] otsu (histogram domain).
// otsu_only: out[0] = out[5] = out[6] = Otsu centre (the intensity threshold), no triangle.
template <typename R>
__global__ void histn_finalize_kernel(const long long* __restrict__ state, const unsigned long long* __restrict__ counts,
                                      const R* __restrict__ edges, int nbins, int log_domain, int otsu_only,
                                      double* __restrict__ work, double* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int k = 0; k < 7; ++k) out[k] = 0.0;
    out[3] = 1.0;
    if (state[NB200_HIST_COUNT] == 0) return;
    out[3] = 0.0;
    double* p = work;
    double* pc = work + nbins;
    double* w_lo = work + 2 * nbins;
    double* s_lo = work + 3 * nbins;
    double* w_hi = work + 4 * nbins;
    double* s_hi = work + 5 * nbins;
    unsigned long long total = 0;
    for (int i = 0; i < nbins; ++i) total += counts[i];
    auto center = [&](int i) -> R { return (edges[i] + edges[i + 1]) / (R)2; };
    for (int i = 0; i < nbins; ++i) {
        p[i] = (double)counts[i] / (double)total;
        pc[i] = p[i] * (double)center(i);
    }
    // ---- Otsu (gpu_functions.py:36-50)
    double aw = 0.0, as = 0.0;
    for (int i = 0; i < nbins; ++i) {
        aw = aw + p[i]; as = as + pc[i];
        w_lo[i] = aw; s_lo[i] = as;
    }
    aw = 0.0; as = 0.0;
    for (int i = nbins - 1; i >= 0; --i) {
        aw = aw + p[i]; as = as + pc[i];
        w_hi[i] = aw; s_hi[i] = as;
    }
    int status = 0;
    double best = 0.0;
    int arg = 0;
    bool have = false, nan_hit = false;
    for (int i = 0; i < nbins - 1; ++i) {
        const double m_lo = s_lo[i] / w_lo[i];
        const double m_hi = s_hi[i + 1] / w_hi[i + 1];
        const double d = m_lo - m_hi;
        const double var = (w_lo[i] * w_hi[i + 1]) * (d * d);
        if (var != var) {                                   // np.argmax returns the first NaN
            if (!nan_hit) { arg = i; nan_hit = true; }
        } else if (!nan_hit && (!have || var > best)) {
            best = var; arg = i; have = true;
        }
    }
    if (nan_hit) status = 1;
    const R otsu = center(arg);
    if (otsu_only) {
        out[0] = (double)otsu; out[1] = (double)otsu; out[2] = (double)otsu;
        out[4] = (double)status; out[5] = (double)otsu; out[6] = (double)otsu;
        return;
    }
    // ---- triangle (gpu_functions.py:65-94)
    int peak = 0;
    double hpk = p[0];
    for (int i = 1; i < nbins; ++i) if (p[i] > hpk) { hpk = p[i]; peak = i; }
    int lo = 0, hi = nbins - 1;
    while (lo < nbins - 1 && !(p[lo] != 0.0)) ++lo;
    while (hi > 0 && !(p[hi] != 0.0)) --hi;
    const bool flip = (peak - lo) < (hi - peak);
    if (flip) { lo = nbins - hi - 1; peak = nbins - peak - 1; }
    const int width = peak - lo;
    R tri;
    if (width <= 0) {
        status = 1;                                          // np.argmax of an empty array raises in the reference
        tri = center(flip ? nbins - lo - 1 : lo);
    } else {
        const double nrm = sqrt(hpk * hpk + (double)((long long)width * width));
        const double hn = hpk / nrm, wn = (double)width / nrm;
        double far = 0.0;
        int at = 0;
        for (int x = 0; x < width; ++x) {
            const int src = x + lo;
            const double y = flip ? p[nbins - 1 - src] : p[src];
            const double len = hn * (double)x - wn * y;
            if (x == 0 || len > far) { far = len; at = x; }
        }
        int lvl = at + lo;
        if (flip) lvl = nbins - lvl - 1;
        tri = center(lvl);
    }
    out[4] = (double)status;
    out[5] = (double)tri;
    out[6] = (double)otsu;
    if (log_domain) {
        const float a = (float)pow(10.0, (double)tri), b = (float)pow(10.0, (double)otsu);
        out[1] = (double)a; out[2] = (double)b;
        out[0] = (double)(a < b ? a : b);
    } else {
        out[1] = (double)tri; out[2] = (double)otsu;
        out[0] = (double)otsu;
    }
}

template <typename R>
int run_histn(const float* vals, long long n, int log_domain, const long long* state, int nbins, int otsu_only, void* edges,
              unsigned long long* counts, double* work, double* out, cudaStream_t st) {
    R* e = static_cast<R*>(edges);
    NB_LAUNCH(histn_edges_kernel<R>, nb::grid_for(nbins + 1, THREADS, 1), THREADS, st, state, nbins, e, counts);
    if (n > 0)
        NB_LAUNCH(histn_bins_kernel<R>, nb::grid_for(n, THREADS, 4), THREADS, st, vals, n, log_domain, state, nbins,
                  (const R*)e, counts);
    NB_LAUNCH(histn_finalize_kernel<R>, 1, 1, st, state, (const unsigned long long*)counts, (const R*)e, nbins, log_domain,
              otsu_only, work, out);
    return nb::check_launch("histn kernels");
}

}  // namespace

extern "C" {

size_t nb200_histn_workspace_bytes(int nbins) {
    // float64 edges (nbins + 1) + uint64 counts (nbins) + float64 work (6 * nbins), 8-byte aligned
    return nbins < 1 ? 0 : sizeof(double) * ((size_t)nbins + 1 + (size_t)nbins + 6 * (size_t)nbins);
}

int nb200_histn_threshold(const float* vals, long long n, int log_domain, int f64_edges, int otsu_only, int nbins,
                          const long long* state, void* workspace, double* out, void* stream) {
    NB_REQUIRE(vals && state && workspace && out && n >= 0, NB200_ERR_ARG, "nb200_histn_threshold: bad argument");
    NB_REQUIRE(nbins >= 2 && nbins <= (1 << 24), NB200_ERR_UNSUPPORTED, "nb200_histn_threshold: %d bins (2 .. 2^24)", nbins);
    NB_REQUIRE(!(f64_edges && log_domain), NB200_ERR_ARG, "nb200_histn_threshold: float64 edges are for integer frames (no log)");
    double* edges = static_cast<double*>(workspace);
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(edges + nbins + 1);
    double* work = reinterpret_cast<double*>(counts + nbins);
    cudaStream_t st = nb::as_stream(stream);
    if (f64_edges) return run_histn<double>(vals, n, log_domain, state, nbins, otsu_only, edges, counts, work, out, st);
    return run_histn<float>(vals, n, log_domain, state, nbins, otsu_only, edges, counts, work, out, st);
}

}  // extern "C"
