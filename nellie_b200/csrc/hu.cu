// HuMomentTracking, per-frame feature extraction — nellie/tracking/hu_tracking.py:
//
//   hu_tracking.py:604-612  frangi frame: log10 of the positive response, negatives shifted so their minimum is 0
//   hu_tracking.py:614-616  distance frame: 3^d maximum filter, doubled
//   hu_tracking.py:392-421  _get_im_bounds: ROI box of every marker, half width ceil(2 * max-filtered distance)
//   hu_tracking.py:341-390  _calculate_mean_and_variance of the non-zero ROI voxels (raw and frangi)
//   hu_tracking.py:225-312  _calculate_normalized_moments + _calculate_hu_moments of the ROI (2-D) or of its three
//                           maximum projections (3-D), :314-325 _log_hu
//
// What is reproduced bit for bit and what is not:
//   * the frangi transform (numpy's float32 log10 restated in devmath.cuh), the doubled maximum filter, the boxes;
//   * the statistics: numpy reduces the zero-padded ROI cube (dense path) or the ROI box (streaming path) of a float32
//     frame with its PAIRWISE float32 summation — restated here leaf for leaf (8 interleaved accumulators per block of
//     <= 128, halves cut at multiples of 8) — and an integer frame with exact integer sums, the squares wrapped to the
//     frame's own width as numpy's `uint16 ** 2` does; mean and variance follow in float64 exactly as written;
//   * the moments: raw moments accumulate in raster order like numpy's reduction over the two image axes (exact integers
//     for integer frames); the central moments use powers by multiplication where numpy calls its SIMD pow (not correctly
//     rounded, not reproducible), so eta / Hu / log-Hu agree with the reference to float64 rounding, not to the bit.
//
// Kernels are barrier-free grid-stride loops (one thread per voxel / marker / projection pixel); the two atomics
// (minimum of the negatives, largest ROI half width) go through nb_atomic_* so that the file also compiles for the host
// (oracle/cuda_emu.h), where the test harness runs the same code serially.
#ifdef NB200_HOST_EMU
#include NB200_HOST_EMU
#else
#include "common.cuh"
#define NB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define nb_atomic_min_u32(p, v) atomicMin((p), (v))
#define nb_atomic_max_i32(p, v) atomicMax((p), (v))
#endif
#include "devmath.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int CTAS_PER_SM = 8;
constexpr unsigned NO_NEGATIVE = 0xFFFFFFFFu;

struct Dims {
    int nz, ny, nx;
    long long plane, total;
};

inline Dims make_dims(int nz, int ny, int nx) {
    Dims d;
    d.nz = nz; d.ny = ny; d.nx = nx;
    d.plane = (long long)ny * nx;
    d.total = d.plane * nz;
    return d;
}

inline unsigned grid_of(long long n) { return nb::grid_for(n, THREADS, CTAS_PER_SM); }

__device__ __forceinline__ unsigned ordered_of(float f) {
    const unsigned u = nb::f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_of_ordered(unsigned u) {
    return nb::u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- frangi frame ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
frangi_log_kernel(const float* __restrict__ frangi, long long n, float* __restrict__ out, unsigned* min_neg) {
    unsigned local = NO_NEGATIVE;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        float v = frangi[idx];
        if (v > 0.0f) v = nb::np_log10f(v);
        out[idx] = v;
        if (v < 0.0f) {
            const unsigned o = ordered_of(v);
            local = o < local ? o : local;
        }
    }
    if (local != NO_NEGATIVE) nb_atomic_min_u32(min_neg, local);
}

__global__ void __launch_bounds__(THREADS)
frangi_shift_kernel(float* __restrict__ out, long long n, const unsigned* __restrict__ min_neg) {
    const unsigned m = *min_neg;
    if (m == NO_NEGATIVE) return;
    const float shift = float_of_ordered(m);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const float v = out[idx];
        if (v < 0.0f) out[idx] = v - shift;
    }
}

// ---- distance frame -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
distance_max_kernel(const float* __restrict__ distance, Dims d, float* __restrict__ out) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(idx / d.plane);
        const long long rem = idx - (long long)z * d.plane;
        const int y = (int)(rem / d.nx);
        const int x = (int)(rem - (long long)y * d.nx);
        const int z0 = z > 0 ? -1 : 0, z1 = z + 1 < d.nz ? 1 : 0;
        const int y0 = y > 0 ? -1 : 0, y1 = y + 1 < d.ny ? 1 : 0;
        const int x0 = x > 0 ? -1 : 0, x1 = x + 1 < d.nx ? 1 : 0;
        float m = distance[idx];                       // mode="reflect" repeats the border voxel: in-range window
        for (int dz = z0; dz <= z1; ++dz)
            for (int dy = y0; dy <= y1; ++dy) {
                const long long row = idx + dz * d.plane + (long long)dy * d.nx;
                for (int dx = x0; dx <= x1; ++dx) {
                    const float v = distance[row + dx];
                    m = v > m ? v : m;
                }
            }
        out[idx] = m * 2.0f;
    }
}

// ---- ROI boxes -------------------------------------------------------------------------------------------------------------
// coords: int64 (n, ndim) in raster order (argwhere of the marker frame); bounds: int32 (n, 6) = lo/hi for Z, Y, X
// (Z = [0, 1) for 2-D frames); max_half: largest ceil(radius) over the markers.
__global__ void __launch_bounds__(THREADS)
bounds_kernel(const long long* __restrict__ coords, long long n, int ndim, const float* __restrict__ distance_max, Dims d,
              int* __restrict__ bounds, int* max_half) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int c[3] = {0, 0, 0};
        for (int a = 0; a < ndim; ++a) c[3 - ndim + a] = (int)coords[i * ndim + a];
        const float radius = distance_max[(long long)c[0] * d.plane + (long long)c[1] * d.nx + c[2]];
        const int r = (int)ceilf(radius);
        const int size[3] = {d.nz, d.ny, d.nx};
        for (int a = 0; a < 3; ++a) {
            int lo = c[a] - r, hi = c[a] + r + 1;
            if (a < 3 - ndim) { lo = 0; hi = 1; }
            lo = lo < 0 ? 0 : (lo > size[a] ? size[a] : lo);
            hi = hi < 0 ? 0 : (hi > size[a] ? size[a] : hi);
            bounds[i * 6 + 2 * a] = lo;
            bounds[i * 6 + 2 * a + 1] = hi;
        }
        nb_atomic_max_i32(max_half, r);
    }
}

// ---- mean / variance of the non-zero ROI voxels -----------------------------------------------------------------------------
struct Roi {
    const float* frame;
    long long plane;
    int nx;
    int lo[3];       // box origin
    int ext[3];      // box extents
    int dim[3];      // extents of the array numpy reduces: the box (streaming) or the zero-padded cube (dense)
};

__device__ __forceinline__ float roi_at(const Roi& r, long long k) {
    // element k (C order) of the reduced array: the frame inside the box, the zero padding of the dense cube outside it
    const long long yx = (long long)r.dim[1] * r.dim[2];
    const int cz = (int)(k / yx);
    const long long rem = k - cz * yx;
    const int cy = (int)(rem / r.dim[2]);
    const int cx = (int)(rem - (long long)cy * r.dim[2]);
    if (cz >= r.ext[0] || cy >= r.ext[1] || cx >= r.ext[2]) return 0.0f;
    return r.frame[(long long)(r.lo[0] + cz) * r.plane + (long long)(r.lo[1] + cy) * r.nx + (r.lo[2] + cx)];
}

// images * mask as numpy evaluates it (v * 1 or v * 0), optionally squared in float32
template <bool SQUARE>
__device__ __forceinline__ float term_at(const Roi& r, long long k) {
    const float v = roi_at(r, k);
    const float m = v != 0.0f ? v : v * 0.0f;
    return SQUARE ? m * m : m;
}

// numpy's pairwise summation of n float32 values a(start .. start+n-1) (numpy/_core/src/umath/loops_utils.h.src,
// pairwise_sum): n < 8 sequential from 0; n <= 128: eight interleaved accumulators, combined as a balanced tree, the
// remainder added one by one; larger n: two halves, the first of length n/2 rounded down to a multiple of 8.
template <bool SQUARE>
__device__ float pairwise_leaf(const Roi& r, long long start, long long n) {
    if (n < 8) {
        float res = 0.0f;
        for (long long i = 0; i < n; ++i) res = res + term_at<SQUARE>(r, start + i);
        return res;
    }
    float a[8];
    for (int k = 0; k < 8; ++k) a[k] = term_at<SQUARE>(r, start + k);
    long long i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) a[k] = a[k] + term_at<SQUARE>(r, start + i + k);
    float res = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    for (; i < n; ++i) res = res + term_at<SQUARE>(r, start + i);
    return res;
}

template <bool SQUARE>
__device__ float pairwise_sum(const Roi& r, long long n) {
    // post-order walk of the halving tree with an explicit stack (depth <= 48 for any 64-bit n)
    long long st_start[48], st_n[48];
    float st_left[48];
    unsigned char st_state[48];
    int top = 0;
    st_start[0] = 0; st_n[0] = n; st_state[0] = 0;
    float value = 0.0f;
    while (top >= 0) {
        if (st_n[top] <= 128) {
            value = pairwise_leaf<SQUARE>(r, st_start[top], st_n[top]);
            --top;
            // hand the value to the parents that are waiting for it
            while (top >= 0 && st_state[top] == 2) {
                value = st_left[top] + value;
                --top;
            }
            if (top >= 0) {                            // parent had its left half pending: store it, descend right
                st_left[top] = value;
                st_state[top] = 2;
                long long n2 = st_n[top] / 2;
                n2 -= n2 % 8;
                st_start[top + 1] = st_start[top] + n2;
                st_n[top + 1] = st_n[top] - n2;
                st_state[top + 1] = 0;
                ++top;
            }
        } else {                                       // inner node: descend into the left half first
            long long n2 = st_n[top] / 2;
            n2 -= n2 % 8;
            st_state[top] = 1;
            st_start[top + 1] = st_start[top];
            st_n[top + 1] = n2;
            st_state[top + 1] = 0;
            ++top;
        }
    }
    return value;
}

// one thread per marker; stats (n, 2) float32 = [mean, variance].  int_bits: 0 = float32 frame, 8 / 16 = unsigned integer
// frame of that width (values exact in the float32 device copy), -8 / -16 would be signed (not supported: caller).
__global__ void __launch_bounds__(THREADS)
roi_stats_kernel(const float* __restrict__ frame, Dims d, const int* __restrict__ bounds, long long n, int ndim, int cube,
                 int int_bits, float* __restrict__ stats) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Roi r;
        r.frame = frame; r.plane = d.plane; r.nx = d.nx;
        bool empty = false;
        for (int a = 0; a < 3; ++a) {
            r.lo[a] = bounds[i * 6 + 2 * a];
            r.ext[a] = bounds[i * 6 + 2 * a + 1] - r.lo[a];
            if (r.ext[a] <= 0) empty = true;
            r.dim[a] = (cube > 0 && a >= 3 - ndim) ? cube : r.ext[a];
        }
        if (3 - ndim > 0) r.dim[0] = 1;
        float mean = 0.0f, var = 0.0f;
        if (empty) {
            // dense: the ROI stays an all-zero cube (count 0 -> 0, 0); streaming: the row keeps its initial zeros
        } else {
            const long long len = (long long)r.dim[0] * r.dim[1] * r.dim[2];
            long long count = 0;
            for (int cz = 0; cz < r.ext[0]; ++cz)
                for (int cy = 0; cy < r.ext[1]; ++cy) {
                    const float* row = frame + (long long)(r.lo[0] + cz) * d.plane + (long long)(r.lo[1] + cy) * d.nx + r.lo[2];
                    for (int cx = 0; cx < r.ext[2]; ++cx) count += row[cx] != 0.0f;
                }
            if (count > 0) {
                const double c = (double)count;
                double m, v;
                if (int_bits > 0) {
                    // exact integer sums; squares wrap to the frame's width (numpy: uint16 ** 2 stays uint16), and
                    // sum ** 2 wraps modulo 2^64 like numpy's uint64
                    const unsigned long long wrap = (1ull << int_bits) - 1ull;
                    unsigned long long s = 0, ss = 0;
                    for (int cz = 0; cz < r.ext[0]; ++cz)
                        for (int cy = 0; cy < r.ext[1]; ++cy) {
                            const float* row = frame + (long long)(r.lo[0] + cz) * d.plane + (long long)(r.lo[1] + cy) * d.nx + r.lo[2];
                            for (int cx = 0; cx < r.ext[2]; ++cx) {
                                const unsigned long long u = (unsigned long long)row[cx];
                                s += u;
                                ss += (u * u) & wrap;
                            }
                        }
                    m = (double)s / c;
                    v = ((double)ss - (double)(s * s) / c) / c;
                } else {
                    const float s = pairwise_sum<false>(r, len);
                    const float ss = pairwise_sum<true>(r, len);
                    m = (double)s / c;
                    v = ((double)ss - (double)(s * s) / c) / c;
                }
                mean = (float)m;
                var = (float)v;
            }
        }
        stats[i * 2] = mean;
        stats[i * 2 + 1] = var;
    }
}

// ---- maximum projections ----------------------------------------------------------------------------------------------------
// proj: float32 (n, n_proj, side, side); projection p of a 3-D ROI drops axis p (0 = Z, 1 = Y, 2 = X), rows / columns are
// the two remaining axes in order; a 2-D ROI has one "projection", the ROI itself.  Only the box part is written (the
// rest of the side x side plane is never read).  dense: the reference's ROI cube is zero-padded to `cube` along the
// dropped axis too, so a 0 takes part in the maximum whenever the box is shorter than the cube.
__global__ void __launch_bounds__(THREADS)
project_kernel(const float* __restrict__ frame, Dims d, const int* __restrict__ bounds, long long n, int ndim, int side,
               int cube, float* __restrict__ proj) {
    const int n_proj = ndim == 3 ? 3 : 1;
    const long long per_marker = (long long)n_proj * side * side;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * per_marker;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / per_marker;
        const int rem = (int)(t - i * per_marker);
        const int p = rem / (side * side);
        const int h = (rem - p * side * side) / side, w = rem % side;
        int lo[3], ext[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = bounds[i * 6 + 2 * a];
            ext[a] = bounds[i * 6 + 2 * a + 1] - lo[a];
        }
        if (ext[0] <= 0 || ext[1] <= 0 || ext[2] <= 0) continue;
        int ah, aw, ad;                                  // row axis, column axis, dropped axis
        if (ndim == 2) { ah = 1; aw = 2; ad = 0; }
        else if (p == 0) { ah = 1; aw = 2; ad = 0; }
        else if (p == 1) { ah = 0; aw = 2; ad = 1; }
        else { ah = 0; aw = 1; ad = 2; }
        if (h >= ext[ah] || w >= ext[aw]) continue;
        const long long stride[3] = {d.plane, (long long)d.nx, 1};
        const float* base = frame + (long long)lo[0] * d.plane + (long long)lo[1] * d.nx + lo[2] + h * stride[ah] + w * stride[aw];
        float m = base[0];
        for (int k = 1; k < ext[ad]; ++k) {
            const float v = base[k * stride[ad]];
            m = v > m ? v : m;
        }
        if (ndim == 3 && cube > ext[ad] && !(m > 0.0f)) m = 0.0f;
        proj[t] = m;
    }
}

// ---- moments -> log-Hu ------------------------------------------------------------------------------------------------------
// one thread per (marker, projection): raw moments M00, M10, M01 in raster order, centroid, the seven central moments the
// Hu invariants use, eta = mu / (M00^((i+j+2)/2) + 1e-12), the six invariants, -sign(hu) log10(max(|hu|, tiny)).
// out: float64 (n, 6 * n_proj).  x = column index, y = row index, first moment index = power of x (hu_tracking.py:243-252).
__global__ void __launch_bounds__(THREADS)
hu_moments_kernel(const float* __restrict__ proj, const int* __restrict__ bounds, long long n, int ndim, int side,
                  int integer_frame, double* __restrict__ out) {
    const int n_proj = ndim == 3 ? 3 : 1;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * n_proj;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / n_proj;
        const int p = (int)(t - i * n_proj);
        int ext[3];
        bool empty = false;
        for (int a = 0; a < 3; ++a) {
            ext[a] = bounds[i * 6 + 2 * a + 1] - bounds[i * 6 + 2 * a];
            if (ext[a] <= 0) empty = true;
        }
        int H, W;
        if (ndim == 2 || p == 0) { H = ext[1]; W = ext[2]; }
        else if (p == 1) { H = ext[0]; W = ext[2]; }
        else { H = ext[0]; W = ext[1]; }
        if (empty) { H = 0; W = 0; }
        const float* img = proj + t * (long long)side * side;
        double m00, m10, m01;
        if (integer_frame) {
            long long s00 = 0, s10 = 0, s01 = 0;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const long long v = (long long)img[y * side + x];
                    s00 += v; s10 += v * x; s01 += v * y;
                }
            m00 = (double)s00; m10 = (double)s10; m01 = (double)s01;
        } else {
            m00 = 0.0; m10 = 0.0; m01 = 0.0;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const double v = (double)img[y * side + x];
                    m00 = m00 + v;                       // (v * x^0) * y^0
                    m10 = m10 + v * (double)x;
                    m01 = m01 + v * (double)y;
                }
        }
        const double x_bar = m10 / (m00 + 1e-12), y_bar = m01 / (m00 + 1e-12);
        double mu20 = 0, mu02 = 0, mu11 = 0, mu30 = 0, mu12 = 0, mu21 = 0, mu03 = 0;
        for (int y = 0; y < H; ++y) {
            const double b = (double)y - y_bar, b2 = b * b, b3 = b2 * b;
            for (int x = 0; x < W; ++x) {
                const double v = (double)img[y * side + x];
                const double a = (double)x - x_bar, a2 = a * a, a3 = a2 * a;
                mu20 = mu20 + v * a2;                    // (v * a^i) * b^j, b^0 = 1
                mu02 = mu02 + v * b2;
                mu11 = mu11 + (v * a) * b;
                mu30 = mu30 + v * a3;
                mu12 = mu12 + (v * a) * b2;
                mu21 = mu21 + (v * a2) * b;
                mu03 = mu03 + v * b3;
            }
        }
        const double d2 = m00 * m00 + 1e-12;             // M00 ** 2.0
        const double d25 = m00 * m00 * sqrt(m00) + 1e-12;  // M00 ** 2.5
        const double e20 = mu20 / d2, e02 = mu02 / d2, e11 = mu11 / d2;
        const double e30 = mu30 / d25, e12 = mu12 / d25, e21 = mu21 / d25, e03 = mu03 / d25;
        double hu[6];
        const double s1 = e30 + e12, s2 = e21 + e03, t1 = e30 - 3 * e12, t2 = 3 * e21 - e03;
        hu[0] = e20 + e02;
        hu[1] = (e20 - e02) * (e20 - e02) + 4 * (e11 * e11);
        hu[2] = t1 * t1 + t2 * t2;
        hu[3] = s1 * s1 + s2 * s2;
        hu[4] = (t1 * s1 * (s1 * s1 - 3 * (s2 * s2)) + t2 * s2 * (3 * (s1 * s1) - s2 * s2));
        hu[5] = ((e20 - e02) * (s1 * s1 - s2 * s2) + 4 * e11 * s1 * s2);
        for (int k = 0; k < 6; ++k) {
            const double h = hu[k];
            double a = fabs(h);
            if (!(a >= 2.2250738585072014e-308)) a = 2.2250738585072014e-308;   // np.maximum(|hu|, tiny); NaN stays out
            const double sign = h > 0.0 ? 1.0 : (h < 0.0 ? -1.0 : 0.0);
            double l = -sign * log10(a);
            if (h != h || !(fabs(l) <= 1.7976931348623157e308)) l = 0.0;       // where(isfinite(log_hu), log_hu, 0)
            out[i * (6 * n_proj) + 6 * p + k] = l;
        }
    }
}

inline bool dims_ok(int nz, int ny, int nx) {
    return nz >= 1 && ny >= 1 && nx >= 1 && (long long)nz * ny * nx < (1ll << 40);
}

}  // namespace

extern "C" {

int nb200_hu_frangi_transform(const float* frangi, long long n, float* out, unsigned int* scratch, void* stream) {
    NB_REQUIRE(frangi && out && scratch && n >= 0, NB200_ERR_ARG, "nb200_hu_frangi_transform: bad argument");
    NB_REQUIRE(frangi != out, NB200_ERR_ARG, "nb200_hu_frangi_transform: in-place is not supported");
    if (n == 0) return NB200_OK;
    cudaStream_t st = nb::as_stream(stream);
    NB_LAUNCH(frangi_log_kernel, grid_of(n), THREADS, st, frangi, n, out, scratch);
    NB_LAUNCH(frangi_shift_kernel, grid_of(n), THREADS, st, out, n, (const unsigned*)scratch);
    return nb::check_launch("frangi transform kernels");
}

int nb200_hu_distance_max(const float* distance, int nz, int ny, int nx, float* out, void* stream) {
    NB_REQUIRE(distance && out && distance != out && dims_ok(nz, ny, nx), NB200_ERR_ARG, "nb200_hu_distance_max: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(distance_max_kernel, grid_of(d.total), THREADS, nb::as_stream(stream), distance, d, out);
    return nb::check_launch("distance_max_kernel");
}

int nb200_hu_bounds(const long long* coords, long long n, int ndim, const float* distance_max, int nz, int ny, int nx,
                    int* bounds, int* max_half, void* stream) {
    NB_REQUIRE(coords && distance_max && bounds && max_half && n >= 0 && (ndim == 2 || ndim == 3) && dims_ok(nz, ny, nx) &&
                   (ndim == 3 || nz == 1), NB200_ERR_ARG, "nb200_hu_bounds: bad argument");
    if (n == 0) return NB200_OK;
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(bounds_kernel, grid_of(n), THREADS, nb::as_stream(stream), coords, n, ndim, distance_max, d, bounds, max_half);
    return nb::check_launch("bounds_kernel");
}

int nb200_hu_roi_stats(const float* frame, int nz, int ny, int nx, const int* bounds, long long n, int ndim, int cube,
                       int int_bits, float* stats, void* stream) {
    NB_REQUIRE(frame && bounds && stats && n >= 0 && (ndim == 2 || ndim == 3) && dims_ok(nz, ny, nx) && cube >= 0,
               NB200_ERR_ARG, "nb200_hu_roi_stats: bad argument");
    NB_REQUIRE(int_bits == 0 || int_bits == 8 || int_bits == 16, NB200_ERR_UNSUPPORTED,
               "nb200_hu_roi_stats: integer frames of %d bits are not supported (0 = float32, 8, 16)", int_bits);
    if (n == 0) return NB200_OK;
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(roi_stats_kernel, grid_of(n), THREADS, nb::as_stream(stream), frame, d, bounds, n, ndim, cube, int_bits, stats);
    return nb::check_launch("roi_stats_kernel");
}

int nb200_hu_log_moments(const float* frame, int nz, int ny, int nx, const int* bounds, long long n, int ndim, int side,
                         int cube, int integer_frame, float* proj, double* out, void* stream) {
    NB_REQUIRE(frame && bounds && proj && out && n >= 0 && (ndim == 2 || ndim == 3) && dims_ok(nz, ny, nx) && side >= 1 &&
                   cube >= 0, NB200_ERR_ARG, "nb200_hu_log_moments: bad argument");
    if (n == 0) return NB200_OK;
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    const long long n_proj = ndim == 3 ? 3 : 1;
    NB_LAUNCH(project_kernel, grid_of(n * n_proj * side * side), THREADS, st, frame, d, bounds, n, ndim, side, cube, proj);
    NB_LAUNCH(hu_moments_kernel, grid_of(n * n_proj), THREADS, st, (const float*)proj, bounds, n, ndim, side, integer_frame,
              out);
    return nb::check_launch("hu moment kernels");
}

}  // extern "C"
