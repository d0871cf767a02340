// F1 — one axis of the cascaded Gaussian (nellie/segmentation/filtering.py:828-835).
//
// scipy.ndimage.gaussian_filter1d semantics: float32 line -> double buffer,
// symmetric correlation  acc = x[i]*w0; for j = r..1: acc += (x[i-j] + x[i+j]) * w[j]
// (all double, product and sum rounded separately), result stored as float32; "reflect"
// boundary = half-sample symmetric.  The f64 accumulation is what makes the blurred volume
// bit-identical to the reference, which the downstream histogram thresholds rely on.
//
// Layout: frame (Z,Y,X) float32, X contiguous.  Y/Z passes: one thread per X column segment,
// marching along the filtered axis with the 2r+1 window held in registers (each input is
// loaded and converted to double once per output line; all global accesses are coalesced
// along X).  X pass: a tile of rows is staged in shared memory as double, then each thread
// produces outputs along X from shared memory.
#include "common.cuh"

namespace {

constexpr int kMaxRadius = 40;

struct GaussWeights {
    double w[kMaxRadius + 1];
};

__device__ __forceinline__ int reflect_index(int i, int n) {
    // half-sample symmetric extension, any distance (scipy NI_EXTEND_REFLECT)
    if (i >= 0 && i < n) return i;
    if (n == 1) return 0;
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}

// ----------------------------- generic fallback (any radius) ---------------------------------
// one thread per output voxel; used for radii beyond the register-window variants
__global__ void __launch_bounds__(256)
gauss_axis_generic(const float* __restrict__ src, float* __restrict__ dst, nb200_vol v, int axis,
                   GaussWeights gw, int radius) {
    const long long plane = (long long)v.ny * v.nx;
    const long long total = (long long)(v.zc1 - v.zc0) * plane;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int zb = v.zc0 + (int)(idx / plane);
        const long long rem = idx - (long long)(zb - v.zc0) * plane;
        const int y = (int)(rem / v.nx);
        const int x = (int)(rem - (long long)y * v.nx);
        const long long base = (long long)zb * plane + (long long)y * v.nx + x;
        double acc = (double)src[base] * gw.w[0];
        for (int j = radius; j >= 1; --j) {
            long long lo, hi;
            if (axis == 0) {
                const int g = zb + v.zg_off;
                const int bl = reflect_index(g - j, v.nz_glob) - v.zg_off;
                const int bh = reflect_index(g + j, v.nz_glob) - v.zg_off;
                lo = base + (long long)(bl - zb) * plane;
                hi = base + (long long)(bh - zb) * plane;
            } else if (axis == 1) {
                lo = base + (long long)(reflect_index(y - j, v.ny) - y) * v.nx;
                hi = base + (long long)(reflect_index(y + j, v.ny) - y) * v.nx;
            } else {
                lo = base + (reflect_index(x - j, v.nx) - x);
                hi = base + (reflect_index(x + j, v.nx) - x);
            }
            const double pair = (double)src[lo] + (double)src[hi];
            acc = acc + pair * gw.w[j];
        }
        dst[base] = (float)acc;
    }
}

// ----------------------------- register-window march along Z or Y ----------------------------
// Each thread owns one x column of one (z or y) line segment of length SEG and slides a
// (2R+1)-deep window of doubles kept in registers.
template <int R, int AXIS>
__global__ void __launch_bounds__(128)
gauss_march(const float* __restrict__ src, float* __restrict__ dst, nb200_vol v, GaussWeights gw, int seg) {
    const long long plane = (long long)v.ny * v.nx;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= v.nx) return;
    // blockIdx.y enumerates the orthogonal coordinate, blockIdx.z the segment along the axis
    int n_axis, a0, a1, other;
    if (AXIS == 0) {
        other = blockIdx.y;                       // y
        n_axis = v.nz_glob;
        a0 = v.zc0 + blockIdx.z * seg;            // buffer coordinates
        a1 = min(a0 + seg, v.zc1);
    } else {
        other = v.zc0 + blockIdx.y;               // buffer z
        n_axis = v.ny;
        a0 = blockIdx.z * seg;
        a1 = min(a0 + seg, v.ny);
    }
    if (a0 >= a1) return;
    const long long stride = (AXIS == 0) ? plane : (long long)v.nx;
    const long long col = (AXIS == 0) ? ((long long)other * v.nx + x) : ((long long)other * plane + x);
    const int goff = (AXIS == 0) ? v.zg_off : 0;

    auto load = [&](int a_buf) -> double {
        // a_buf: coordinate along the axis in buffer space (may be outside the frame)
        const int r = reflect_index(a_buf + goff, n_axis) - goff;
        return (double)__ldg(src + col + (long long)r * stride);
    };

    double win[2 * R + 1];
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) win[k + 1] = load(a0 - R + k);
    for (int a = a0; a < a1; a += (2 * R + 1)) {
        // unrolled over one full rotation of the window so indices stay compile-time
#pragma unroll
        for (int u = 0; u < 2 * R + 1; ++u) {
            if (a + u < a1) {
                // shift: logical window position k lives in win[(k + u + 1) % (2R+1)]
                win[u % (2 * R + 1)] = load(a + u + R);
                // logical index helper
#define NB_WIN(k) win[((k) + u + 1) % (2 * R + 1)]
                double acc = NB_WIN(R) * gw.w[0];
#pragma unroll
                for (int j = R; j >= 1; --j) {
                    const double pair = NB_WIN(R - j) + NB_WIN(R + j);
                    acc = acc + pair * gw.w[j];
                }
#undef NB_WIN
                dst[col + (long long)(a + u) * stride] = (float)acc;
            }
        }
    }
}

// ----------------------------- X pass through shared memory ----------------------------------
// Block = 32 x ROWS threads handles ROWS rows by TX outputs; the row segment with its halo is
// staged as double in shared memory (one conversion per input), padded to dodge bank conflicts.
template <int R>
__global__ void __launch_bounds__(256)
gauss_x_smem(const float* __restrict__ src, float* __restrict__ dst, nb200_vol v, GaussWeights gw) {
    constexpr int TX = 256;          // outputs per row per block
    constexpr int ROWS = 4;          // rows per block
    constexpr int W = TX + 2 * R;    // staged width
    __shared__ double tile[ROWS][W + 1];
    const long long nrows = (long long)(v.zc1 - v.zc0) * v.ny;
    const long long row0 = (long long)blockIdx.x * ROWS;
    const int x0 = blockIdx.y * TX;
    for (int rr = 0; rr < ROWS; ++rr) {
        const long long row = row0 + rr;
        if (row >= nrows) break;
        const float* line = src + ((long long)v.zc0 * v.ny + row) * v.nx;
        for (int i = threadIdx.x; i < W; i += blockDim.x) {
            const int xs = reflect_index(x0 - R + i, v.nx);
            tile[rr][i] = (double)__ldg(line + xs);
        }
    }
    __syncthreads();
    for (int rr = 0; rr < ROWS; ++rr) {
        const long long row = row0 + rr;
        if (row >= nrows) break;
        float* out = dst + ((long long)v.zc0 * v.ny + row) * v.nx;
        for (int i = threadIdx.x; i < TX; i += blockDim.x) {
            const int x = x0 + i;
            if (x >= v.nx) break;
            const double* c = &tile[rr][i + R];
            double acc = c[0] * gw.w[0];
#pragma unroll
            for (int j = R; j >= 1; --j) {
                const double pair = c[-j] + c[j];
                acc = acc + pair * gw.w[j];
            }
            out[x] = (float)acc;
        }
    }
}

template <int R>
int launch_fixed(const float* src, float* dst, const nb200_vol& v, int axis, const GaussWeights& gw,
                 cudaStream_t st) {
    const int nzc = v.zc1 - v.zc0;
    if (axis == 2) {
        dim3 grid((unsigned)(((long long)nzc * v.ny + 3) / 4), (v.nx + 255) / 256);
        gauss_x_smem<R><<<grid, 256, 0, st>>>(src, dst, v, gw);
        return nb::check_launch("gauss_x_smem");
    }
    const int n_axis = axis == 0 ? nzc : v.ny;
    const int n_other = axis == 0 ? v.ny : nzc;
    // segments long enough to amortise the 2R window prologue, short enough to fill the GPU
    int seg = 64;
    const long long cols = (long long)((v.nx + 127) / 128) * n_other;
    while (seg > 16 && cols * ((n_axis + seg - 1) / seg) < 4LL * nb::sm_count()) seg /= 2;
    dim3 grid((v.nx + 127) / 128, n_other, (n_axis + seg - 1) / seg);
    if (grid.y > 65535u || grid.z > 65535u) return 1;  // fall back to the generic kernel
    if (axis == 0) gauss_march<R, 0><<<grid, 128, 0, st>>>(src, dst, v, gw, seg);
    else gauss_march<R, 1><<<grid, 128, 0, st>>>(src, dst, v, gw, seg);
    return nb::check_launch("gauss_march");
}

}  // namespace

extern "C" int nb200_gauss_axis(const float* src, float* dst, const nb200_vol* vol, int axis,
                                const double* weights, int radius, void* stream) {
    NB_REQUIRE(src && dst && vol && weights, NB200_ERR_ARG, "nb200_gauss_axis: null argument");
    NB_REQUIRE(src != dst, NB200_ERR_ARG, "nb200_gauss_axis: in-place is not supported (ping-pong buffers)");
    NB_REQUIRE(axis >= 0 && axis <= 2, NB200_ERR_ARG, "nb200_gauss_axis: axis %d", axis);
    NB_REQUIRE(radius >= 0 && radius <= kMaxRadius, NB200_ERR_UNSUPPORTED,
               "nb200_gauss_axis: radius %d exceeds %d", radius, kMaxRadius);
    const nb200_vol v = *vol;
    NB_REQUIRE(v.zc0 >= 0 && v.zc1 <= v.nz_buf && v.zc0 <= v.zc1 && v.ny > 0 && v.nx > 0, NB200_ERR_ARG,
               "nb200_gauss_axis: bad volume window");
    if (v.zc0 == v.zc1) return NB200_OK;
    if (axis == 0) {
        // every reflected tap must land inside the buffer
        const int glo = v.zc0 + v.zg_off - radius, ghi = v.zc1 - 1 + v.zg_off + radius;
        const int need_lo = glo < 0 ? 0 : glo, need_hi = ghi >= v.nz_glob ? v.nz_glob - 1 : ghi;
        NB_REQUIRE(need_lo - v.zg_off >= 0 && need_hi - v.zg_off < v.nz_buf, NB200_ERR_ARG,
                   "nb200_gauss_axis: Z halo too small for radius %d", radius);
    }
    GaussWeights gw;
    for (int i = 0; i <= kMaxRadius; ++i) gw.w[i] = i <= radius ? weights[i] : 0.0;
    cudaStream_t st = nb::as_stream(stream);
    int rc = 1;
    switch (radius) {
#define NB_CASE(R) case R: rc = launch_fixed<R>(src, dst, v, axis, gw, st); break;
        NB_CASE(1) NB_CASE(2) NB_CASE(3) NB_CASE(4) NB_CASE(5) NB_CASE(6) NB_CASE(7) NB_CASE(8)
        NB_CASE(9) NB_CASE(10) NB_CASE(11) NB_CASE(12)
#undef NB_CASE
        default: break;
    }
    if (rc <= 0) return rc;
    const long long total = (long long)(v.zc1 - v.zc0) * v.ny * v.nx;
    gauss_axis_generic<<<nb::grid_for(total, 256, 8), 256, 0, st>>>(src, dst, v, axis, gw, radius);
    return nb::check_launch("gauss_axis_generic");
}
