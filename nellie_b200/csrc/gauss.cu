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

// ----------------------------- Z pass, vectorised register-window march ----------------------
// Fast path of axis 0 (nx % COLS == 0, aligned base): each thread owns COLS adjacent x columns
// (one 16-/8-byte access per plane) and slides a (2R+1)-deep window of doubles along Z.  All
// loads are unconditional (the plane index is clamped to the last plane a valid output needs)
// so the 2R+1 loads of one unrolled rotation are issued back to back; only stores are guarded.
template <int COLS> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int COLS>
__device__ __forceinline__ typename VecT<COLS>::type load_raw(const float* p) {
    return __ldg(reinterpret_cast<const typename VecT<COLS>::type*>(p));
}
template <int COLS>
__device__ __forceinline__ void widen(const typename VecT<COLS>::type& t, double (&out)[COLS]) {
    if constexpr (COLS == 4) {
        out[0] = (double)t.x; out[1] = (double)t.y; out[2] = (double)t.z; out[3] = (double)t.w;
    } else if constexpr (COLS == 2) {
        out[0] = (double)t.x; out[1] = (double)t.y;
    } else {
        out[0] = (double)t;
    }
}
template <int COLS>
__device__ __forceinline__ void store_cols(float* p, const double (&a)[COLS]) {
    if constexpr (COLS == 4) {
        *reinterpret_cast<float4*>(p) = make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
    } else if constexpr (COLS == 2) {
        *reinterpret_cast<float2*>(p) = make_float2((float)a[0], (float)a[1]);
    } else {
        *p = (float)a[0];
    }
}

template <int R, int COLS>
__global__ void __launch_bounds__(128)
gauss_z_vec(const float* __restrict__ src, float* __restrict__ dst, nb200_vol v, GaussWeights gw, int seg) {
    constexpr int N = 2 * R + 1;
    constexpr int D = N < 4 ? N : 4;                    // software prefetch distance (planes in flight per thread)
    using Raw = typename VecT<COLS>::type;
    const long long plane = (long long)v.ny * v.nx;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * COLS;
    if (x >= v.nx) return;
    const int a0 = v.zc0 + blockIdx.z * seg;            // buffer coordinates
    const int a1 = min(a0 + seg, v.zc1);
    if (a0 >= a1) return;
    const long long col = (long long)blockIdx.y * v.nx + x;
    const int goff = v.zg_off, nzg = v.nz_glob, last = a1 - 1 + R;
    auto fetch = [&](int a_buf) -> Raw {
        const int r = reflect_index(min(a_buf, last) + goff, nzg) - goff;
        return load_raw<COLS>(src + col + (long long)r * plane);
    };
    double win[N][COLS];
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) widen<COLS>(fetch(a0 - R + k), win[k + 1]);
    // raw[u] holds the plane that enters the window at step u of a rotation; it is requested D steps early
    Raw raw[N];
#pragma unroll
    for (int u = 0; u < D; ++u) raw[u] = fetch(a0 + u + R);
    for (int a = a0; a < a1; a += N) {
#pragma unroll
        for (int u = 0; u < N; ++u) {
            widen<COLS>(raw[u], win[u % N]);
            raw[(u + D) % N] = fetch(a + u + D + R);
#define NB_WIN(k) win[((k) + u + 1) % N]
            double acc[COLS];
#pragma unroll
            for (int c = 0; c < COLS; ++c) acc[c] = NB_WIN(R)[c] * gw.w[0];
#pragma unroll
            for (int j = R; j >= 1; --j) {
#pragma unroll
                for (int c = 0; c < COLS; ++c) {
                    const double pair = NB_WIN(R - j)[c] + NB_WIN(R + j)[c];
                    acc[c] = acc[c] + pair * gw.w[j];
                }
            }
#undef NB_WIN
            if (a + u < a1) store_cols<COLS>(dst + col + (long long)(a + u) * plane, acc);
        }
    }
}

// ----------------------------- fused Y + X pass on a plane tile ------------------------------
// One CTA = 32 rows x TX columns of one plane.  The tile plus its halo (R rows; HALO = 4 or 8 columns so
// that staged rows start 16-byte aligned) is staged as float32 in shared memory (A): interior tiles with
// one 128-bit load per lane and row, tiles touching the frame border (reflect = half-sample symmetric) or
// unaligned volumes element by element.  The Y pass marches a register window of doubles down each staged
// column (two 16-row segments per column) and stores float32 results (B) exactly as scipy stores the
// intermediate array; the X pass marches along the rows of B (a warp = 4 rows x 8 column segments, pitch
// 132: at most 2-way bank conflicts) into a staging tile (C, aliasing A) that is written out with 128-bit
// coalesced stores.  HBM traffic: 4 B read (+ halo from L2) and 4 B written per voxel for two of the three
// axes; every input is converted to double once per pass.
template <int R> struct YXGeo {
    static constexpr int TY = 32;
    static constexpr int HALO = (R <= 4) ? 4 : 8;
    static constexpr int TX = 128 - 2 * HALO;            // 120 / 112 outputs per row: multiples of 8
    static constexpr int C0 = HALO - R;                  // first staged column the Y pass needs
    static constexpr int W = TX + 2 * R;                 // columns the Y pass produces
    static constexpr int H = TY + 2 * R;                 // staged rows
    static constexpr int PA = 128, PB = 132;
    static constexpr int SX = TX / 8;                    // X-pass outputs per thread
    static constexpr int A_FLOATS = (H * PA > TY * PB) ? H * PA : TY * PB;
};

template <int R>
__global__ void __launch_bounds__(256)
gauss_yx_tile(const float* __restrict__ src, float* __restrict__ dst, nb200_vol v, GaussWeights gy, GaussWeights gx,
              int vec_ok) {
    using G = YXGeo<R>;
    constexpr int N = 2 * R + 1;
    __shared__ __align__(16) float A[G::A_FLOATS];
    __shared__ __align__(16) float B[G::TY * G::PB];
    const int x0 = blockIdx.x * G::TX, y0 = blockIdx.y * G::TY;
    const long long plane = (long long)v.ny * v.nx;
    const float* sp = src + (long long)(v.zc0 + blockIdx.z) * plane;
    float* dp = dst + (long long)(v.zc0 + blockIdx.z) * plane;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- stage rows y0-R .. y0+TY+R-1, columns x0-HALO .. x0-HALO+127 ----
    if (vec_ok) {
        const int xg = x0 - G::HALO + 4 * lane;          // nx % 4 == 0: a group of four is inside or outside as a whole
        const bool x_in = xg >= 0 && xg < v.nx;
        int xr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xr[k] = reflect_index(xg + k, v.nx);
        constexpr int PER = (G::H + 7) / 8;
        float4 t[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int r = warp + 8 * i;
            if (r < G::H) {
                const float* rowp = sp + (long long)reflect_index(y0 - R + r, v.ny) * v.nx;
                if (x_in) t[i] = __ldg(reinterpret_cast<const float4*>(rowp + xg));
                else t[i] = make_float4(__ldg(rowp + xr[0]), __ldg(rowp + xr[1]), __ldg(rowp + xr[2]), __ldg(rowp + xr[3]));
            }
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int r = warp + 8 * i;
            if (r < G::H) *reinterpret_cast<float4*>(&A[r * G::PA + 4 * lane]) = t[i];
        }
    } else {
        const int c = tid & 127;
        const int xs = reflect_index(x0 - G::HALO + c, v.nx);
        for (int r = tid >> 7; r < G::H; r += 2) {
            const int ys = reflect_index(y0 - R + r, v.ny);
            A[r * G::PA + c] = __ldg(sp + (long long)ys * v.nx + xs);
        }
    }
    __syncthreads();
    // ---- Y pass: staged column C0 + c, rows [16*seg, 16*seg+16) of the tile ----
    {
        const int c = tid & 127, r0 = (tid >> 7) * 16;
        if (c < G::W) {
            const float* a = A + r0 * G::PA + G::C0 + c;
            double win[N];
#pragma unroll
            for (int k = 0; k < 2 * R; ++k) win[k] = (double)a[k * G::PA];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                win[(j + 2 * R) % N] = (double)a[(j + 2 * R) * G::PA];
#define NB_WIN(k) win[((k) + j) % N]
                double acc = NB_WIN(R) * gy.w[0];
#pragma unroll
                for (int t = R; t >= 1; --t) {
                    const double pair = NB_WIN(R - t) + NB_WIN(R + t);
                    acc = acc + pair * gy.w[t];
                }
#undef NB_WIN
                B[(r0 + j) * G::PB + c] = (float)acc;
            }
        }
    }
    __syncthreads();
    // ---- X pass: row = 4*warp + (lane & 3), columns [SX*seg, SX*seg + SX), seg = lane >> 2 ----
    {
        const int row = 4 * warp + (lane & 3), c0 = (lane >> 2) * G::SX;
        const float* b = B + row * G::PB + c0;
        float* cdst = A + row * G::PB + c0;              // C aliases A (dead since the last barrier)
        double win[N];
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) win[k] = (double)b[k];
#pragma unroll
        for (int j = 0; j < G::SX; ++j) {
            win[(j + 2 * R) % N] = (double)b[j + 2 * R];
#define NB_WIN(k) win[((k) + j) % N]
            double acc = NB_WIN(R) * gx.w[0];
#pragma unroll
            for (int t = R; t >= 1; --t) {
                const double pair = NB_WIN(R - t) + NB_WIN(R + t);
                acc = acc + pair * gx.w[t];
            }
#undef NB_WIN
            cdst[j] = (float)acc;
        }
    }
    __syncthreads();
    // ---- coalesced write-out ----
    if (vec_ok) {
        constexpr int Q = G::TX / 4;                     // float4 groups per row
        for (int i = tid; i < G::TY * Q; i += 256) {
            const int r = i / Q, q = i - r * Q;
            const int y = y0 + r;
            if (y < v.ny && x0 + 4 * q < v.nx)
                *reinterpret_cast<float4*>(dp + (long long)y * v.nx + x0 + 4 * q) =
                    *reinterpret_cast<const float4*>(&A[r * G::PB + 4 * q]);
        }
    } else {
        for (int i = tid; i < G::TY * G::TX; i += 256) {
            const int r = i / G::TX, c = i - r * G::TX;
            const int y = y0 + r, x = x0 + c;
            if (y < v.ny && x < v.nx) dp[(long long)y * v.nx + x] = A[r * G::PB + c];
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
    if (axis == 0 && R <= 12 && v.ny <= 65535) {
        // vectorised march: 4 columns per thread up to R = 5 (window of 44 doubles), else 2
        constexpr int COLS = R <= 5 ? 4 : 2;
        const bool aligned = (v.nx % COLS == 0) && (((uintptr_t)src | (uintptr_t)dst) % (4 * COLS) == 0);
        if (aligned) {
            const int gx = (v.nx / COLS + 127) / 128;
            int seg = 256;
            while (seg > 32 && (long long)gx * v.ny * ((nzc + seg - 1) / seg) < 16LL * nb::sm_count()) seg /= 2;
            dim3 grid(gx, v.ny, (nzc + seg - 1) / seg);
            if (grid.z <= 65535u) {
                gauss_z_vec<R, COLS><<<grid, 128, 0, st>>>(src, dst, v, gw, seg);
                return nb::check_launch("gauss_z_vec");
            }
        }
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

extern "C" int nb200_gauss_yx(const float* src, float* dst, const nb200_vol* vol, const double* wy, const double* wx,
                              int radius, void* stream) {
    NB_REQUIRE(src && dst && vol && wy && wx, NB200_ERR_ARG, "nb200_gauss_yx: null argument");
    NB_REQUIRE(src != dst, NB200_ERR_ARG, "nb200_gauss_yx: in-place is not supported (ping-pong buffers)");
    NB_REQUIRE(radius >= 1 && radius <= 8, NB200_ERR_UNSUPPORTED,
               "nb200_gauss_yx: radius %d outside 1..8 (use two nb200_gauss_axis calls)", radius);
    const nb200_vol v = *vol;
    NB_REQUIRE(v.zc0 >= 0 && v.zc1 <= v.nz_buf && v.zc0 <= v.zc1 && v.ny > 0 && v.nx > 0, NB200_ERR_ARG,
               "nb200_gauss_yx: bad volume window");
    if (v.zc0 == v.zc1) return NB200_OK;
    NB_REQUIRE(v.zc1 - v.zc0 <= 65535, NB200_ERR_UNSUPPORTED, "nb200_gauss_yx: more than 65535 planes");
    GaussWeights gy, gx;
    for (int i = 0; i <= kMaxRadius; ++i) {
        gy.w[i] = i <= radius ? wy[i] : 0.0;
        gx.w[i] = i <= radius ? wx[i] : 0.0;
    }
    const int vec_ok = (v.nx % 4 == 0) && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0);
    cudaStream_t st = nb::as_stream(stream);
    switch (radius) {
#define NB_CASE(R)                                                                                      \
    case R: {                                                                                           \
        dim3 grid((v.nx + YXGeo<R>::TX - 1) / YXGeo<R>::TX, (v.ny + YXGeo<R>::TY - 1) / YXGeo<R>::TY,    \
                  v.zc1 - v.zc0);                                                                       \
        NB_REQUIRE(grid.y <= 65535u, NB200_ERR_UNSUPPORTED, "nb200_gauss_yx: ny too large");            \
        gauss_yx_tile<R><<<grid, 256, 0, st>>>(src, dst, v, gy, gx, vec_ok);                                    \
        break;                                                                                          \
    }
        NB_CASE(1) NB_CASE(2) NB_CASE(3) NB_CASE(4) NB_CASE(5) NB_CASE(6) NB_CASE(7) NB_CASE(8)
#undef NB_CASE
        default: break;
    }
    return nb::check_launch("gauss_yx_tile");
}
