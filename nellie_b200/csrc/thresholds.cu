// F2 / F3 / F5 / U1 / U2 / L2 / L3 — sampling lattice and 256-bin histogram thresholds, on device.
//
// Reference: nellie/segmentation/filtering.py:328-380, :407-444; nellie/utils/gpu_functions.py:23-94;
// nellie/segmentation/labelling.py:385-465.  Exact numpy semantics restated in SURVEY.md A.3/A.4:
// np.histogram with float32 scalars builds float32 edges with np.linspace and corrects the
// computed bin index against the edges; Otsu / triangle run in float64 with sequential
// cumulative sums.  The scalars never leave the device: the next kernel reads them from `sp`.
#include "common.cuh"
#include "devmath.cuh"

namespace {

constexpr int NB = NB200_HIST_NBINS;

// ------------------------------------------------------------------------------------------
// sampling
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lattice_sample_kernel(const float* __restrict__ src, nb200_vol v, int sz, int sy, int sx, int g_first,
                      int n_lat_z, int ly, int lx, float* __restrict__ out) {
    const long long total = (long long)n_lat_z * ly * lx;
    const long long plane = (long long)v.ny * v.nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int kz = (int)(i / ((long long)ly * lx));
        const int rem = (int)(i - (long long)kz * ly * lx);
        const int ky = rem / lx, kx = rem - ky * lx;
        const int zb = g_first + kz * sz - v.zg_off;
        out[i] = __ldg(src + (long long)zb * plane + (long long)(ky * sy) * v.nx + kx * sx);
    }
}

__global__ void __launch_bounds__(256)
strided_sample_kernel(const float* __restrict__ src, long long n_out, long long offset, long long step,
                      const float* __restrict__ gate, float gate_thresh, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out;
         i += (long long)gridDim.x * blockDim.x) {
        const long long j = offset + i * step;
        float val = __ldg(src + j);
        if (gate != nullptr && !(__ldg(gate + j) > gate_thresh)) val = 0.0f;
        out[i] = val;
    }
}

// ------------------------------------------------------------------------------------------
// histogram state
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool transformed(float raw, int tf, float divisor, float& out) {
    // keep rule of the reference: arr[arr > 0] on the array the histogram is taken of
    if (tf == NB200_TF_DIV) {
        const float q = raw / divisor;   // frob = sqrt(frob_sq) / max_abs  (filtering.py:562)
        if (!(q > 0.0f)) return false;
        out = q;
        return true;
    }
    if (!(raw > 0.0f)) return false;
    out = (tf == NB200_TF_LOG10) ? nb::np_log10f(raw) : raw;   // numpy float32 log10, bit-exact (devmath.cuh)
    return true;
}

__global__ void hist_reset_kernel(long long* state) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= NB200_HIST_WORDS) return;
    long long val = 0;
    if (i == NB200_HIST_MIN) val = 0xffffffffLL;
    state[i] = val;
}

__global__ void __launch_bounds__(256)
hist_minmax_kernel(const float* __restrict__ vals, long long n, int tf, const double* __restrict__ divisor,
                   long long* __restrict__ state) {
    const float dv = divisor ? (float)(*divisor) : 1.0f;
    uint32_t lo = 0xffffffffu, hi = 0u;
    unsigned long long cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float t;
        if (transformed(__ldg(vals + i), tf, dv, t)) {
            const uint32_t k = nb::float_to_ordered(t);
            lo = min(lo, k);
            hi = max(hi, k);
            ++cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ uint32_t s_lo[8], s_hi[8];
    __shared__ unsigned long long s_cnt[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; s_cnt[w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            lo = min(lo, s_lo[k]); hi = max(hi, s_hi[k]); cnt += s_cnt[k];
        }
        if (cnt) {
            atomicMin((unsigned long long*)&state[NB200_HIST_MIN], (unsigned long long)lo);
            atomicMax((unsigned long long*)&state[NB200_HIST_MAX], (unsigned long long)hi);
            atomicAdd((unsigned long long*)&state[NB200_HIST_COUNT], cnt);
        }
    }
}

// float32 bin edges exactly as np.linspace(first, last, 257, dtype=float32) builds them
__device__ void build_edges(float first, float last, float* edges /*NB+1*/, int tid, int nthreads) {
    if (first == last) { first = first - 0.5f; last = last + 0.5f; }  // numpy _get_outer_edges
    const float delta = last - first;
    const float step = delta / (float)NB;
    for (int i = tid; i <= NB; i += nthreads) {
        float y = (float)i;
        if (step == 0.0f) { y = y / (float)NB; y = y * delta; }
        else y = y * step;
        y = y + first;
        if (i == NB) y = last;
        edges[i] = y;
    }
}

__device__ __forceinline__ void outer_edges(const long long* state, float& first, float& last) {
    first = nb::ordered_to_float((uint32_t)state[NB200_HIST_MIN]);
    last = nb::ordered_to_float((uint32_t)state[NB200_HIST_MAX]);
}

__global__ void __launch_bounds__(256)
hist_bins_kernel(const float* __restrict__ vals, long long n, int tf, const double* __restrict__ divisor,
                 long long* __restrict__ state) {
    __shared__ float edges[NB + 1];
    __shared__ unsigned int local[NB];
    if (state[NB200_HIST_COUNT] == 0) return;
    float first, last;
    outer_edges(state, first, last);
    build_edges(first, last, edges, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < NB; i += blockDim.x) local[i] = 0;
    __syncthreads();
    const float e_first = edges[0], e_last = edges[NB];
    const float denom = e_last - e_first;
    const float dv = divisor ? (float)(*divisor) : 1.0f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float t;
        if (!transformed(__ldg(vals + i), tf, dv, t)) continue;
        if (!(t >= e_first) || !(t <= e_last)) continue;
        // numpy _histogram: index from the scaled offset, then the +-1 edge correction
        const float f = ((t - e_first) / denom) * (float)NB;
        int b = (int)f;
        if (b == NB) b = NB - 1;
        if (t < edges[b]) --b;
        else if (t >= edges[b + 1] && b != NB - 1) ++b;
        atomicAdd(&local[b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NB; i += blockDim.x)
        if (local[i]) atomicAdd((unsigned long long*)&state[NB200_HIST_BINS + i], (unsigned long long)local[i]);
}

// ------------------------------------------------------------------------------------------
// Otsu + triangle on the 256 counts (float64, sequential cumulative sums like np.cumsum)
// ------------------------------------------------------------------------------------------
struct TwoThresholds {
    float tri, otsu;
    int status;  // 0 ok, 1 degenerate (reference raises / yields NaN)
};

// Block-cooperative (256 threads, all must call).  The cumulative sums run in ONE thread in index order, like
// np.cumsum; everything without a loop-carried dependency (normalised counts, products, the 2 x 255 float64 divisions of
// the class means, the between-class variance, the triangle distances) is spread over the block.  The single-thread
// form took ~100 us per call (12 calls per frame, and the part of a Z-sharded step that does not shrink with GPUs).
constexpr int FIN_THREADS = 256;
struct FinalizeScratch {
    float edges[NB + 1];
    float centers[NB];
    double p[NB], pc[NB], w_lo[NB], s_lo[NB], w_hi[NB], s_hi[NB], val[NB];
    long long total;
    int arg;
    TwoThresholds out;
};

__device__ TwoThresholds otsu_triangle(const long long* state, FinalizeScratch& w) {
    const int t = threadIdx.x;
    float first, last;
    outer_edges(state, first, last);
    build_edges(first, last, w.edges, t, FIN_THREADS);
    if (t == 0) {
        long long total = 0;
        for (int i = 0; i < NB; ++i) total += state[NB200_HIST_BINS + i];
        w.total = total;
        w.out.tri = 0.f; w.out.otsu = 0.f; w.out.status = 0;
    }
    __syncthreads();
    if (t < NB) {
        w.centers[t] = (w.edges[t] + w.edges[t + 1]) / 2.0f;
        w.p[t] = (double)state[NB200_HIST_BINS + t] / (double)w.total;
        w.pc[t] = w.p[t] * (double)w.centers[t];
    }
    __syncthreads();
    // ---- Otsu (gpu_functions.py:36-50)
    if (t == 0) {                    // reverse cumulative weight / weighted sum
        double aw = 0.0, as = 0.0;
        for (int i = NB - 1; i >= 0; --i) {
            aw = aw + w.p[i];
            as = as + w.pc[i];
            w.w_hi[i] = aw;
            w.s_hi[i] = as;
        }
    } else if (t == 32) {            // forward ones, on another warp
        double aw = 0.0, as = 0.0;
        for (int i = 0; i < NB; ++i) {
            aw = aw + w.p[i];
            as = as + w.pc[i];
            w.w_lo[i] = aw;
            w.s_lo[i] = as;
        }
    }
    __syncthreads();
    if (t < NB - 1) {
        const double m_lo = w.s_lo[t] / w.w_lo[t];
        const double m_hi = w.s_hi[t + 1] / w.w_hi[t + 1];
        const double d = m_lo - m_hi;
        w.val[t] = (w.w_lo[t] * w.w_hi[t + 1]) * (d * d);
    }
    __syncthreads();
    if (t == 0) {
        double best = 0.0;
        int arg = 0;
        bool have = false, nan_hit = false;
        for (int i = 0; i < NB - 1; ++i) {
            const double var = w.val[i];
            if (var != var) {  // np.argmax returns the first NaN
                if (!nan_hit) { arg = i; nan_hit = true; }
            } else if (!nan_hit && (!have || var > best)) {
                best = var; arg = i; have = true;
            }
        }
        if (nan_hit) w.out.status = 1;
        w.out.otsu = w.centers[arg];
    }
    __syncthreads();
    // ---- triangle (gpu_functions.py:65-94)
    __shared__ int tri_lo, tri_peak, tri_width, tri_flip;
    __shared__ double tri_hn, tri_wn;
    if (t == 0) {
        int peak = 0;
        double hpk = w.p[0];
        for (int i = 1; i < NB; ++i) if (w.p[i] > hpk) { hpk = w.p[i]; peak = i; }
        int lo = 0, hi = NB - 1;
        while (lo < NB - 1 && !(w.p[lo] != 0.0)) ++lo;
        while (hi > 0 && !(w.p[hi] != 0.0)) --hi;
        const bool flip = (peak - lo) < (hi - peak);
        if (flip) { lo = NB - hi - 1; peak = NB - peak - 1; }
        const int width = peak - lo;
        tri_lo = lo; tri_peak = peak; tri_width = width; tri_flip = flip ? 1 : 0;
        if (width <= 0) {
            w.out.status = 1;  // np.argmax of an empty array raises ValueError in the reference
            w.out.tri = w.centers[flip ? NB - lo - 1 : lo];
        } else {
            const double nrm = sqrt(hpk * hpk + (double)((long long)width * width));
            tri_hn = hpk / nrm;
            tri_wn = (double)width / nrm;
        }
    }
    __syncthreads();
    if (tri_width > 0) {
        if (t < tri_width) {
            const int src = t + tri_lo;
            const double y = tri_flip ? w.p[NB - 1 - src] : w.p[src];
            w.val[t] = tri_hn * (double)t - tri_wn * y;
        }
        __syncthreads();
        if (t == 0) {
            double best = 0.0; int arg = 0;
            for (int x = 0; x < tri_width; ++x) {
                const double len = w.val[x];
                if (x == 0 || len > best) { best = len; arg = x; }
            }
            int lvl = arg + tri_lo;
            if (tri_flip) lvl = NB - lvl - 1;
            w.out.tri = w.centers[lvl];
        }
    }
    __syncthreads();
    return w.out;
}

__global__ void __launch_bounds__(FIN_THREADS) finalize_gamma_kernel(const long long* state, double* sp) {
    __shared__ FinalizeScratch work;
    const double eps = 1.1920928955078125e-07;  // np.finfo(np.float32).eps
    double gamma = eps;
    double status = 0.0;
    const bool have = state[NB200_HIST_COUNT] > 0;      // block-uniform
    TwoThresholds t;
    if (have) t = otsu_triangle(state, work);
    if (threadIdx.x != 0) return;
    if (have) {
        sp[NB200_SP_TRI] = (double)t.tri;
        sp[NB200_SP_OTSU] = (double)t.otsu;
        gamma = (double)fminf(t.tri, t.otsu);
        if (t.tri != t.tri || t.otsu != t.otsu) gamma = (double)(t.tri + t.otsu);
        if (gamma <= 0.0) gamma = eps;
        status = (double)t.status;
    }
    sp[NB200_SP_GAMMA] = gamma;
    sp[NB200_SP_GAMMA_SQ] = 2.0 * (gamma * gamma);
    sp[NB200_SP_STATUS] = status;
}

__global__ void finalize_max_abs_kernel(const long long* hstats, double* sp) {
    if (threadIdx.x != 0) return;
    float m = nb::u2f((uint32_t)hstats[NB200_HS_MAX_ABS_BITS]);
    if (!(m > 0.0f)) m = 1.0f;  // filtering.py:560-561
    sp[NB200_SP_MAX_ABS] = (double)m;
    // Fast constant-divisor division (verified for numerators with exponent in [-90, 90] and divisors in
    // [2^-12, 2^12]) is safe when every non-zero |g| lies in [2^-30, 2^60]: a non-zero first difference is then
    // >= ulp(2^-30) = 2^-53 and <= 2^61, a first derivative lies in [2^-65, 2^73], and a non-zero difference of
    // first derivatives in [ulp(2^-65), 2^74] = [2^-88, 2^74].  Otherwise the kernels redo / run with IEEE division.
    const uint32_t compl_min = (uint32_t)hstats[NB200_HS_MIN_NZ_COMPL];
    const float g_min = compl_min == 0u ? 1.0f : nb::u2f(0x7f800000u - compl_min);
    const float g_max = nb::u2f((uint32_t)hstats[NB200_HS_MAX_G_BITS]);
    // hstats[FALLBACK]: nb200_hessian_stats_fast could not prove its maximum exact for this volume
    const bool safe = g_min >= 9.313225746154785e-10f && g_max <= 1.152921504606847e18f && hstats[NB200_HS_FALLBACK] == 0;
    sp[NB200_SP_UNSAFE] = safe ? 0.0 : 1.0;
}

// fast != 0: no exact max frob_sq unless the exact redo pass ran (sp[UNSAFE]); emptiness of the mask from the bounds
// 1 <= max frob <= 3, classification thresholds of nb200_frangi_fast from the proven error bound (hessian_fast.cu)
__global__ void __launch_bounds__(FIN_THREADS)
finalize_frob_kernel(const long long* state, const long long* hstats, double fixed_thresh,
                                     double division, double* sp, int fast, double max_scale, int mask_enabled) {
    __shared__ FinalizeScratch work;
    const bool use_hist = mask_enabled && !(fixed_thresh == fixed_thresh) && state[NB200_HIST_COUNT] > 0;   // block-uniform
    TwoThresholds t;
    if (use_hist) t = otsu_triangle(state, work);
    if (threadIdx.x != 0) return;
    double thr = 0.0;
    double status = sp[NB200_SP_STATUS];
    if (!mask_enabled) {
        thr = 0.0;
    } else if (fixed_thresh == fixed_thresh) {
        thr = fixed_thresh;
    } else if (use_hist) {
        thr = (double)fminf(t.tri, t.otsu);
        if (t.status) status = 1.0;
    }
    const float max_abs = (float)sp[NB200_SP_MAX_ABS];
    const float top = sqrtf(nb::u2f((uint32_t)hstats[NB200_HS_MAX_FROBSQ_BITS])) / max_abs;
    double cut;
    bool any;
    if (division == 0.0) {       // filtering.py:428-430: mask = frob > 0
        cut = 0.0;
    } else {
        cut = thr / division;
    }
    if (!mask_enabled) cut = -INFINITY;                  // every voxel passes (frob > -inf)
    any = top > (float)cut;     // sqrt and division are monotone: max frob decides emptiness
    double ambig = 0.0;
    if (fast && sp[NB200_SP_UNSAFE] == 0.0) {
        // The voxel attaining max|H| = M has frob_sq >= RN(M*M) (rounding is monotone, all terms are >= 0), hence
        // frob >= 0.999; every voxel has frob_sq <= 9.0001 M^2, hence frob <= 3.001.
        const float raw_m = nb::u2f((uint32_t)hstats[NB200_HS_MAX_ABS_BITS]);
        const float cutf = (float)cut;
        if (!(raw_m > 0.0f)) any = 0.0f > cutf;          // all-zero Hessian: frob == 0 everywhere
        else if (cutf < 0.99f) any = true;
        else if (cutf >= 3.01f) any = false;
        else { any = true; ambig = 1.0; }                // nb200_finalize_frob_resolve decides with the exact maximum
    }
    if (!mask_enabled) { any = true; ambig = 0.0; }
    // K3 tests frob_sq directly: sqrt and the division by max_abs are monotone, so the mask
    // "sqrt(fs)/max_abs > cut" is an up-set {fs >= fs_min}; find its smallest member by bisection over the
    // (ordered) bit patterns of the non-negative floats.  NaN cut -> fs_min = NaN -> nothing passes.
    {
        const float cutf = (float)cut;
        uint32_t lo = 0u, hi = 0x7f800000u;           // f(+0) = 0 never passes a cut >= 0; f(inf) = inf
        float fs_min = nb::u2f(0x7fc00000u);
        if (cutf == cutf && sqrtf(nb::u2f(hi)) / max_abs > cutf) {
            if (sqrtf(nb::u2f(lo)) / max_abs > cutf) hi = lo;
            while (hi - lo > 1u) {
                const uint32_t mid = lo + (hi - lo) / 2u;
                if (sqrtf(nb::u2f(mid)) / max_abs > cutf) hi = mid; else lo = mid;
            }
            fs_min = nb::u2f(hi);
        }
        if (!mask_enabled) fs_min = -INFINITY;            // Filter(mask=False): every voxel passes (filtering.py:910-933)
        sp[NB200_SP_FROBSQ_MIN] = (double)fs_min;
        if (fast) {
            const double u = 5.9604644775390625e-08;
            const double gmax = (double)nb::u2f((uint32_t)hstats[NB200_HS_MAX_G_BITS]);
            const double delta = 40.0 * u * gmax * max_scale;
            // the approximate provably-zero test needs ||H||_F <= 1.376 ||H~||_F: frob_sq~ >= 64 delta^2 (and a normal k * fs)
            const double guard = fmax(64.0 * delta * delta, 1e-18);
            double lo, hi;
            if (fs_min != fs_min) { lo = INFINITY; hi = INFINITY; }          // NaN cut: nothing passes
            else if (fs_min == -INFINITY) { lo = -INFINITY; hi = guard; }
            else {
                const double f = (double)fs_min, sq = sqrt(f);
                const double e0 = 6.0 * delta * sq + 9.0 * delta * delta + 64.0 * u * f;
                lo = f - 1.5 * e0;
                hi = (delta <= sq / 8.0) ? fmax(f + 2.5 * e0, guard) : INFINITY;
            }
            sp[NB200_SP_FS_LO] = lo;
            sp[NB200_SP_FS_HI] = hi;
            sp[NB200_SP_ZT_C] = 2.5 * delta;
            sp[NB200_SP_DELTA] = delta;
        }
    }
    sp[NB200_SP_AMBIG] = ambig;
    sp[NB200_SP_FROB_THR] = thr;
    sp[NB200_SP_FROB_CUT] = cut;
    sp[NB200_SP_SKIP] = any ? 0.0 : 1.0;
    sp[NB200_SP_STATUS] = status;
}

// Z-sharded frames: every rank all-gathers its [histogram state | Hessian stats] record; this kernel folds the gathered
// records into the local one with the right operator per word (identical result on every rank).
//   stage 0 (after hist_minmax / hessian stats): word MIN -> min, word MAX -> max, Hessian stats -> max
//   stage 1 (after hist_bins):                    count + 256 bins -> sum,          Hessian stats -> max
// `count` records per rank (one per sigma when a Z-sharded frame reduces all its sigmas at one point): rank r's records
// start at gathered + r * count * NB200_STATE_WORDS
__global__ void fold_records_kernel(const long long* __restrict__ gathered, int world, int count, int stage,
                                    long long* __restrict__ state) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count * NB200_STATE_WORDS) return;
    const int i = j % NB200_STATE_WORDS;           // word inside its record
    const bool is_hs = i >= NB200_HIST_WORDS;
    int op = -1;                                   // 0 min, 1 max, 2 sum
    if (is_hs) op = 1;
    else if (stage == 0) op = i == NB200_HIST_MIN ? 0 : (i == NB200_HIST_MAX ? 1 : -1);
    else op = i >= NB200_HIST_COUNT ? 2 : -1;
    if (op < 0) return;
    const long long stride = (long long)count * NB200_STATE_WORDS;
    long long acc = gathered[j];
    for (int r = 1; r < world; ++r) {
        const long long v = gathered[(long long)r * stride + j];
        acc = op == 0 ? (v < acc ? v : acc) : (op == 1 ? (v > acc ? v : acc) : acc + v);
    }
    state[j] = acc;
}

__global__ void finalize_frob_resolve_kernel(const long long* hstats, double* sp) {
    if (threadIdx.x != 0 || sp[NB200_SP_AMBIG] == 0.0 || sp[NB200_SP_UNSAFE] != 0.0) return;
    const float max_abs = (float)sp[NB200_SP_MAX_ABS];
    const float top = sqrtf(nb::u2f((uint32_t)hstats[NB200_HS_MAX_FROBSQ_BITS])) / max_abs;
    sp[NB200_SP_SKIP] = top > (float)sp[NB200_SP_FROB_CUT] ? 0.0 : 1.0;
    sp[NB200_SP_AMBIG] = 0.0;
}

__global__ void __launch_bounds__(FIN_THREADS) finalize_label_kernel(const long long* state, int log_domain, double* out) {
    __shared__ FinalizeScratch work;
    const bool have = state[NB200_HIST_COUNT] > 0;      // block-uniform
    TwoThresholds t;
    if (have) t = otsu_triangle(state, work);
    if (threadIdx.x != 0) return;
    out[0] = 0.0; out[1] = 0.0; out[2] = 0.0; out[3] = 1.0; out[4] = 0.0; out[5] = 0.0; out[6] = 0.0;
    if (!have) return;
    out[3] = 0.0;
    out[4] = (double)t.status;
    out[5] = (double)t.tri;        // the thresholds in the domain of the histogram (log10 for the frangi threshold):
    out[6] = (double)t.otsu;       // the host applies 10 ** np.float32(.) itself, exactly as labelling.py:452-455 does
    if (log_domain) {
        // 10 ** np.float32 -> float32 power; evaluate in f64 and round once (labelling.py:452-455)
        const float a = (float)pow(10.0, (double)t.tri);
        const float b = (float)pow(10.0, (double)t.otsu);
        out[1] = (double)a;
        out[2] = (double)b;
        out[0] = (double)fminf(a, b);
    } else {
        out[1] = (double)t.tri;
        out[2] = (double)t.otsu;
        out[0] = (double)t.otsu;
    }
}

// ------------------------------------------------------------------------------------------
// Otsu threshold of an INTEGER frame (Label's intensity gate on uint8 / uint16 data, labelling.py:457-465):
// numpy bins integer samples with float64 edges (np.histogram promotes the bin type of integer data to float64) and
// otsu_threshold (gpu_functions.py:23-50) then works on float64 bin centres; the float32 kernels above would round the
// edges and the centre.  The samples arrive as float32 (exact for |v| < 2^24), min / max come from the float32 keys of
// nb200_hist_minmax; the edges, the binning and the centres here are float64.  Separate kernels on purpose: the float32
// path (every fixture of the Filter) is not touched.
// ------------------------------------------------------------------------------------------
__device__ void build_edges_f64(double first, double last, double* edges /*NB+1*/, int tid, int nthreads) {
    if (first == last) { first = first - 0.5; last = last + 0.5; }  // numpy _get_outer_edges
    const double delta = last - first;
    const double step = delta / (double)NB;
    for (int i = tid; i <= NB; i += nthreads) {
        double y = (double)i;
        if (step == 0.0) { y = y / (double)NB; y = y * delta; }
        else y = y * step;
        y = y + first;
        if (i == NB) y = last;
        edges[i] = y;
    }
}

__global__ void __launch_bounds__(256)
hist_bins_f64_kernel(const float* __restrict__ vals, long long n, long long* __restrict__ state) {
    __shared__ double edges[NB + 1];
    __shared__ unsigned int local[NB];
    if (state[NB200_HIST_COUNT] == 0) return;
    float ff, fl;
    outer_edges(state, ff, fl);
    build_edges_f64((double)ff, (double)fl, edges, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < NB; i += blockDim.x) local[i] = 0;
    __syncthreads();
    const double e_first = edges[0], e_last = edges[NB];
    const double denom = e_last - e_first;      // numpy: _unsigned_subtract(last_edge, first_edge), exact for integers
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float raw = __ldg(vals + i);
        if (!(raw > 0.0f)) continue;            // arr[arr > 0] (labelling.py:385-438)
        const double t = (double)raw;
        if (!(t >= e_first) || !(t <= e_last)) continue;
        const double f = ((t - e_first) / denom) * (double)NB;
        int b = (int)f;
        if (b == NB) b = NB - 1;
        if (t < edges[b]) --b;
        else if (t >= edges[b + 1] && b != NB - 1) ++b;
        atomicAdd(&local[b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NB; i += blockDim.x)
        if (local[i]) atomicAdd((unsigned long long*)&state[NB200_HIST_BINS + i], (unsigned long long)local[i]);
}

// out[0] = Otsu threshold (float64 bin centre), out[3] = 1 when there were no samples, out[4] = 1 when degenerate (NaN)
__global__ void __launch_bounds__(FIN_THREADS) finalize_otsu_f64_kernel(const long long* state, double* out) {
    __shared__ double edges[NB + 1], centers[NB], p[NB], pc[NB], w_lo[NB], s_lo[NB], w_hi[NB], s_hi[NB], val[NB];
    __shared__ long long total_s;
    const int t = threadIdx.x;
    const bool have = state[NB200_HIST_COUNT] > 0;      // block-uniform
    if (!have) {
        if (t == 0) { out[0] = 0.0; out[1] = 0.0; out[2] = 0.0; out[3] = 1.0; out[4] = 0.0; out[5] = 0.0; out[6] = 0.0; }
        return;
    }
    float ff, fl;
    outer_edges(state, ff, fl);
    build_edges_f64((double)ff, (double)fl, edges, t, FIN_THREADS);
    if (t == 0) {
        long long total = 0;
        for (int i = 0; i < NB; ++i) total += state[NB200_HIST_BINS + i];
        total_s = total;
    }
    __syncthreads();
    if (t < NB) {
        centers[t] = (edges[t] + edges[t + 1]) / 2.0;
        p[t] = (double)state[NB200_HIST_BINS + t] / (double)total_s;
        pc[t] = p[t] * centers[t];
    }
    __syncthreads();
    if (t == 0) {                    // reverse cumulative sums, in index order like np.cumsum on the reversed array
        double aw = 0.0, as = 0.0;
        for (int i = NB - 1; i >= 0; --i) {
            aw = aw + p[i];
            as = as + pc[i];
            w_hi[i] = aw;
            s_hi[i] = as;
        }
    } else if (t == 32) {
        double aw = 0.0, as = 0.0;
        for (int i = 0; i < NB; ++i) {
            aw = aw + p[i];
            as = as + pc[i];
            w_lo[i] = aw;
            s_lo[i] = as;
        }
    }
    __syncthreads();
    if (t < NB - 1) {
        const double m_lo = s_lo[t] / w_lo[t];
        const double m_hi = s_hi[t + 1] / w_hi[t + 1];
        const double d = m_lo - m_hi;
        val[t] = (w_lo[t] * w_hi[t + 1]) * (d * d);
    }
    __syncthreads();
    if (t != 0) return;
    double best = 0.0;
    int arg = 0;
    bool got = false, nan_hit = false;
    for (int i = 0; i < NB - 1; ++i) {
        const double var = val[i];
        if (var != var) {  // np.argmax returns the first NaN
            if (!nan_hit) { arg = i; nan_hit = true; }
        } else if (!nan_hit && (!got || var > best)) {
            best = var; arg = i; got = true;
        }
    }
    out[0] = centers[arg]; out[1] = centers[arg]; out[2] = centers[arg];
    out[3] = 0.0; out[4] = nan_hit ? 1.0 : 0.0; out[5] = centers[arg]; out[6] = centers[arg];
}

// ------------------------------------------------------------------------------------------
// percentile by 3-pass radix select over the positive samples (F11, filtering.py:963)
// scratch layout (int64): [0..2047] digit histogram, [2048] prefix, [2049] rank (k),
//                         [2050] n_pos, [2051] found bits of s[k], [2052] bits of s[k+1]
// ------------------------------------------------------------------------------------------
constexpr int SEL_BINS = 2048;

__global__ void select_reset_kernel(long long* scratch) {
    for (int i = threadIdx.x; i < SEL_BINS + 8; i += blockDim.x) scratch[i] = 0;
}

__global__ void __launch_bounds__(256)
select_count_kernel(const float* __restrict__ vals, long long n, long long* __restrict__ scratch) {
    unsigned long long c = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        c += (__ldg(vals + i) > 0.0f) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd((unsigned long long*)&scratch[SEL_BINS + 2], c);
}

// pass p in {0,1,2}: digits are bits [21,32), [10,21), [0,10) of the (positive) float pattern
__device__ __forceinline__ void digit_of(int pass, uint32_t u, uint32_t& dig, uint32_t& prefix_bits) {
    if (pass == 0) { dig = u >> 21; prefix_bits = 0; }
    else if (pass == 1) { dig = (u >> 10) & 0x7ffu; prefix_bits = u >> 21; }
    else { dig = u & 0x3ffu; prefix_bits = u >> 10; }
}

__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ vals, long long n, int pass, int which,
                   long long* __restrict__ scratch) {
    __shared__ unsigned int local[SEL_BINS];
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) local[i] = 0;
    __syncthreads();
    const uint32_t want = (uint32_t)scratch[SEL_BINS + 0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float f = __ldg(vals + i);
        if (!(f > 0.0f)) continue;
        uint32_t dig, pre;
        digit_of(pass, nb::f2u(f), dig, pre);
        if (pass == 0 || pre == want) atomicAdd(&local[dig], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
        if (local[i]) atomicAdd((unsigned long long*)&scratch[i], (unsigned long long)local[i]);
    (void)which;
}

// single thread: walk the digit histogram to locate rank k, update prefix / residual rank
__global__ void select_scan_kernel(long long* scratch, int pass, int which) {
    if (threadIdx.x != 0) return;
    long long k = scratch[SEL_BINS + 1];
    long long run = 0;
    int d = 0;
    const int nd = pass == 2 ? 1024 : 2048;
    for (; d < nd; ++d) {
        const long long c = scratch[d];
        if (run + c > k) break;
        run += c;
    }
    if (d >= nd) d = nd - 1;
    const uint32_t prev = (uint32_t)scratch[SEL_BINS + 0];
    uint32_t pre = pass == 0 ? (uint32_t)d : (pass == 1 ? ((prev << 11) | (uint32_t)d) : ((prev << 10) | (uint32_t)d));
    scratch[SEL_BINS + 0] = (long long)pre;
    scratch[SEL_BINS + 1] = k - run;
    if (pass == 2) scratch[SEL_BINS + 3 + which] = (long long)pre;  // full 32-bit pattern
    for (int i = 0; i < SEL_BINS; ++i) scratch[i] = 0;
}

__global__ void select_begin_kernel(long long* scratch, double q_percent, int which) {
    if (threadIdx.x != 0) return;
    // numpy 'linear': q = float32(q)/float32(100); virtual = (n-1)*q in float32; k = floor(virtual)
    const long long n = scratch[SEL_BINS + 2];
    const float q = (float)q_percent / 100.0f;
    const float virt = (float)(n - 1) * q;
    long long k = (long long)floorf(virt);
    if (n > 0 && virt >= (float)(n - 1)) k = n - 1;     // _get_indexes: clamp to the last element
    if (k < 0) k = 0;
    long long kk = k + which;
    if (kk > n - 1) kk = n - 1;
    if (n > 0 && virt >= (float)(n - 1)) kk = n - 1;
    scratch[SEL_BINS + 0] = 0;
    scratch[SEL_BINS + 1] = kk;
}

__global__ void select_finish_kernel(const long long* scratch, double q_percent, double* out) {
    if (threadIdx.x != 0) return;
    const long long n = scratch[SEL_BINS + 2];
    out[1] = (double)n;
    if (n <= 0) { out[0] = 0.0; return; }
    const float a = nb::u2f((uint32_t)scratch[SEL_BINS + 3]);
    const float b = nb::u2f((uint32_t)scratch[SEL_BINS + 4]);
    const float q = (float)q_percent / 100.0f;
    const float virt = (float)(n - 1) * q;
    float g = virt - floorf(virt);
    if (virt >= (float)(n - 1)) g = 0.0f;  // previous == next == last
    // numpy _lerp in float32
    const float diff = b - a;
    float res = a + diff * g;
    if (g >= 0.5f) res = b - diff * (1.0f - g);
    out[0] = (double)res;
}

}  // namespace

// ============================================================================================
extern "C" {

int nb200_lattice_sample(const float* src, const nb200_vol* vol, int sz, int sy, int sx, float* out,
                         void* stream) {
    NB_REQUIRE(src && vol && out && sz > 0 && sy > 0 && sx > 0, NB200_ERR_ARG, "nb200_lattice_sample: bad argument");
    const nb200_vol v = *vol;
    const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off;
    const int g_first = ((g0 + sz - 1) / sz) * sz;
    const int n_lat_z = g_first < g1 ? (g1 - 1 - g_first) / sz + 1 : 0;
    const int ly = (v.ny + sy - 1) / sy, lx = (v.nx + sx - 1) / sx;
    const long long total = (long long)n_lat_z * ly * lx;
    if (total == 0) return NB200_OK;
    lattice_sample_kernel<<<nb::grid_for(total, 256, 8), 256, 0, nb::as_stream(stream)>>>(
        src, v, sz, sy, sx, g_first, n_lat_z, ly, lx, out);
    return nb::check_launch("lattice_sample");
}

int nb200_strided_sample(const float* src, long long n, long long offset, long long step, const float* gate,
                         float gate_thresh, float* out, void* stream) {
    NB_REQUIRE(src && out && step > 0 && offset >= 0, NB200_ERR_ARG, "nb200_strided_sample: bad argument");
    const long long n_out = offset < n ? (n - offset + step - 1) / step : 0;
    if (n_out == 0) return NB200_OK;
    strided_sample_kernel<<<nb::grid_for(n_out, 256, 8), 256, 0, nb::as_stream(stream)>>>(
        src, n_out, offset, step, gate, gate_thresh, out);
    return nb::check_launch("strided_sample");
}

int nb200_hist_reset(long long* state, void* stream) {
    NB_REQUIRE(state, NB200_ERR_ARG, "nb200_hist_reset: null state");
    hist_reset_kernel<<<2, 256, 0, nb::as_stream(stream)>>>(state);
    return nb::check_launch("hist_reset");
}

int nb200_hist_minmax(const float* vals, long long n, int transform, const double* divisor, long long* state,
                      void* stream) {
    NB_REQUIRE(vals && state && n >= 0, NB200_ERR_ARG, "nb200_hist_minmax: bad argument");
    NB_REQUIRE(transform != NB200_TF_DIV || divisor, NB200_ERR_ARG, "nb200_hist_minmax: divisor required");
    if (n == 0) return NB200_OK;
    hist_minmax_kernel<<<nb::grid_for(n, 256, 4), 256, 0, nb::as_stream(stream)>>>(vals, n, transform, divisor, state);
    return nb::check_launch("hist_minmax");
}

int nb200_hist_bins(const float* vals, long long n, int transform, const double* divisor, long long* state,
                    void* stream) {
    NB_REQUIRE(vals && state && n >= 0, NB200_ERR_ARG, "nb200_hist_bins: bad argument");
    NB_REQUIRE(transform != NB200_TF_DIV || divisor, NB200_ERR_ARG, "nb200_hist_bins: divisor required");
    if (n == 0) return NB200_OK;
    hist_bins_kernel<<<nb::grid_for(n, 256, 4), 256, 0, nb::as_stream(stream)>>>(vals, n, transform, divisor, state);
    return nb::check_launch("hist_bins");
}

int nb200_finalize_gamma(const long long* state, double* sp, void* stream) {
    NB_REQUIRE(state && sp, NB200_ERR_ARG, "nb200_finalize_gamma: null argument");
    finalize_gamma_kernel<<<1, FIN_THREADS, 0, nb::as_stream(stream)>>>(state, sp);
    return nb::check_launch("finalize_gamma");
}

int nb200_finalize_max_abs(const long long* hstats, double* sp, void* stream) {
    NB_REQUIRE(hstats && sp, NB200_ERR_ARG, "nb200_finalize_max_abs: null argument");
    finalize_max_abs_kernel<<<1, 32, 0, nb::as_stream(stream)>>>(hstats, sp);
    return nb::check_launch("finalize_max_abs");
}

int nb200_finalize_frob(const long long* state, const long long* hstats, double fixed_thresh, double division,
                        double* sp, void* stream) {
    NB_REQUIRE(state && hstats && sp, NB200_ERR_ARG, "nb200_finalize_frob: null argument");
    finalize_frob_kernel<<<1, FIN_THREADS, 0, nb::as_stream(stream)>>>(state, hstats, fixed_thresh, division, sp, 0, 0.0, 1);
    return nb::check_launch("finalize_frob");
}

int nb200_finalize_frob_fast(const long long* state, const long long* hstats, double fixed_thresh, double division,
                             double max_scale, int mask_enabled, double* sp, void* stream) {
    NB_REQUIRE(state && hstats && sp && max_scale >= 0.0, NB200_ERR_ARG, "nb200_finalize_frob_fast: bad argument");
    finalize_frob_kernel<<<1, FIN_THREADS, 0, nb::as_stream(stream)>>>(state, hstats, fixed_thresh, division, sp,
                                                              max_scale > 0.0 ? 1 : 0, max_scale, mask_enabled);
    return nb::check_launch("finalize_frob_fast");
}

int nb200_fold_records(const long long* gathered, int world, int stage, long long* state, void* stream) {
    NB_REQUIRE(gathered && state && world >= 1 && (stage == 0 || stage == 1), NB200_ERR_ARG, "nb200_fold_records: bad argument");
    fold_records_kernel<<<(NB200_STATE_WORDS + 255) / 256, 256, 0, nb::as_stream(stream)>>>(gathered, world, 1, stage, state);
    return nb::check_launch("fold_records");
}

int nb200_fold_records_n(const long long* gathered, int world, int count, int stage, long long* state, void* stream) {
    NB_REQUIRE(gathered && state && world >= 1 && count >= 1 && count <= 4096 && (stage == 0 || stage == 1), NB200_ERR_ARG,
               "nb200_fold_records_n: bad argument");
    fold_records_kernel<<<(count * NB200_STATE_WORDS + 255) / 256, 256, 0, nb::as_stream(stream)>>>(gathered, world, count, stage, state);
    return nb::check_launch("fold_records_n");
}

int nb200_finalize_frob_resolve(const long long* hstats, double* sp, void* stream) {
    NB_REQUIRE(hstats && sp, NB200_ERR_ARG, "nb200_finalize_frob_resolve: null argument");
    finalize_frob_resolve_kernel<<<1, 32, 0, nb::as_stream(stream)>>>(hstats, sp);
    return nb::check_launch("finalize_frob_resolve");
}

int nb200_finalize_label_threshold(const long long* state, int log_domain, double* out, void* stream) {
    NB_REQUIRE(state && out, NB200_ERR_ARG, "nb200_finalize_label_threshold: null argument");
    finalize_label_kernel<<<1, FIN_THREADS, 0, nb::as_stream(stream)>>>(state, log_domain, out);
    return nb::check_launch("finalize_label_threshold");
}

int nb200_hist_bins_f64(const float* vals, long long n, long long* state, void* stream) {
    NB_REQUIRE(vals && state && n >= 0, NB200_ERR_ARG, "nb200_hist_bins_f64: bad argument");
    if (n == 0) return NB200_OK;
    hist_bins_f64_kernel<<<nb::grid_for(n, 256, 4), 256, 0, nb::as_stream(stream)>>>(vals, n, state);
    return nb::check_launch("hist_bins_f64");
}

int nb200_finalize_otsu_f64(const long long* state, double* out, void* stream) {
    NB_REQUIRE(state && out, NB200_ERR_ARG, "nb200_finalize_otsu_f64: null argument");
    finalize_otsu_f64_kernel<<<1, FIN_THREADS, 0, nb::as_stream(stream)>>>(state, out);
    return nb::check_launch("finalize_otsu_f64");
}

int nb200_percentile(const float* samples, long long n, double q_percent, long long* scratch, double* out,
                     void* stream) {
    NB_REQUIRE(samples && scratch && out && n >= 0, NB200_ERR_ARG, "nb200_percentile: bad argument");
    cudaStream_t st = nb::as_stream(stream);
    select_reset_kernel<<<1, 256, 0, st>>>(scratch);
    if (n > 0) select_count_kernel<<<nb::grid_for(n, 256, 4), 256, 0, st>>>(samples, n, scratch);
    for (int which = 0; which < 2; ++which) {
        select_begin_kernel<<<1, 32, 0, st>>>(scratch, q_percent, which);
        for (int pass = 0; pass < 3 && n > 0; ++pass) {
            select_hist_kernel<<<nb::grid_for(n, 256, 2), 256, 0, st>>>(samples, n, pass, which, scratch);
            select_scan_kernel<<<1, 32, 0, st>>>(scratch, pass, which);
        }
    }
    select_finish_kernel<<<1, 32, 0, st>>>(scratch, q_percent, out);
    return nb::check_launch("percentile");
}

}  // extern "C"
