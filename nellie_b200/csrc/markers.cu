// Markers stage — nellie/segmentation/mocap_marking.py, full-volume branch:
//
//   mocap_marking.py:657      mask = label frame > 0
//   mocap_marking.py:440      border = binary_dilation(mask, iterations=1) ^ mask   (6-/4-neighbour shell outside the mask)
//   mocap_marking.py:444-447  distance = distance_transform_edt(mask).astype(float32), clamped to 2 * max_radius_px
//   mocap_marking.py:488-505  per scale: -LoG * sigma^2, negatives to 0, 3^d local maxima inside the mask, best scale wins
//   mocap_marking.py:595-606  non-maximum suppression of the peaks on the raw intensity, (2 d + 1)^d window
//
// The Gaussian second-derivative passes of scipy.ndimage.gaussian_laplace are nb200_gauss_axis / nb200_gauss_yx calls with
// order-2 taps (gauss.cu); everything here is integer / comparison work plus one float32 multiply, i.e. exact, so the stage's
// three outputs are bit-identical to the reference's.  All kernels are plain grid-stride loops over voxels without shared
// memory or barriers: coalesced along X, HBM/L2-bound, and — through the NB_LAUNCH macro — compilable for the host, where
// the test harness (oracle/markers_host.cpp) runs the very same kernel bodies and entry points against scipy.
//
// Work is skipped where the reference's result cannot depend on it: the distance passes return at once for background
// voxels (squared distance 0 ends the min-plus scan), the 27-voxel maximum test runs only for voxels inside the mask whose
// response beats the best scale so far, and the suppression window is scanned only at peak voxels — so the expensive parts
// touch the few percent of the frame that is foreground, and the rest of each pass streams one or two bytes per voxel.
#ifdef NB200_HOST_EMU
#include NB200_HOST_EMU
#else
#include "common.cuh"
#define NB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

namespace {

constexpr int THREADS = 256;
constexpr int CTAS_PER_SM = 8;
constexpr int EDT_INF = 0xFFFF;        // "no background voxel inside the window"; real squared distances stay below
constexpr int EDT_MAX_WINDOW = 147;    // 3 * 147^2 = 64827 < EDT_INF

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

struct Pos {
    int z, y, x;
};

__device__ __forceinline__ Pos pos_of(long long idx, const Dims& d) {
    Pos p;
    p.z = (int)(idx / d.plane);
    const long long rem = idx - (long long)p.z * d.plane;
    p.y = (int)(rem / d.nx);
    p.x = (int)(rem - (long long)p.y * d.nx);
    return p;
}

// ---- mask + border shell ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
mask_border_kernel(const int* __restrict__ labels, Dims d, unsigned char* __restrict__ mask,
                   unsigned char* __restrict__ border) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const bool m = labels[idx] > 0;
        bool shell = false;
        if (!m) {
            const Pos p = pos_of(idx, d);
            shell = (p.x > 0 && labels[idx - 1] > 0) || (p.x + 1 < d.nx && labels[idx + 1] > 0) ||
                    (p.y > 0 && labels[idx - d.nx] > 0) || (p.y + 1 < d.ny && labels[idx + d.nx] > 0) ||
                    (p.z > 0 && labels[idx - d.plane] > 0) || (p.z + 1 < d.nz && labels[idx + d.plane] > 0);
        }
        mask[idx] = m ? 1 : 0;
        border[idx] = shell ? 1 : 0;
    }
}

// ---- exact Euclidean distance transform, windowed -------------------------------------------------------------------------
// X pass: squared distance along the row to the nearest background voxel within `window` (0 for background itself).
__global__ void __launch_bounds__(THREADS)
edt_x_kernel(const unsigned char* __restrict__ mask, Dims d, int window, unsigned short* __restrict__ g) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        int best = 0;
        if (mask[idx]) {
            const int x = (int)(idx % d.nx);
            best = EDT_INF;
            for (int k = 1; k <= window; ++k) {
                const bool left = x - k >= 0 && !mask[idx - k];
                const bool right = x + k < d.nx && !mask[idx + k];
                if (left || right) {
                    best = k * k;
                    break;
                }
            }
        }
        g[idx] = (unsigned short)best;
    }
}

// Y / Z pass: best = min over |k| <= window of src[a + k] + k^2 along the axis.  A candidate at offset k costs at least
// k^2, so the scan ends as soon as k^2 >= best (immediately for background, after a few steps inside thin structures).
// FINAL: writes min(float32(sqrt(float64(best))), clamp) — the reference's astype + np.minimum — instead of the squares.
template <bool FINAL>
__global__ void __launch_bounds__(THREADS)
edt_axis_kernel(const unsigned short* __restrict__ src, Dims d, int axis, int window, float clamp,
                unsigned short* __restrict__ dst, float* __restrict__ distance) {
    const long long stride = axis == 0 ? d.plane : (long long)d.nx;
    const int n_axis = axis == 0 ? d.nz : d.ny;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const Pos p = pos_of(idx, d);
        const int a = axis == 0 ? p.z : p.y;
        int best = src[idx];
        for (int k = 1; k <= window; ++k) {
            const int kk = k * k;
            if (kk >= best) break;
            if (a - k >= 0) {
                const int c = (int)src[idx - k * stride] + kk;
                best = c < best ? c : best;
            }
            if (a + k < n_axis) {
                const int c = (int)src[idx + k * stride] + kk;
                best = c < best ? c : best;
            }
        }
        if (FINAL) {
            float v = clamp;
            if (best < EDT_INF) {
                const float r = (float)sqrt((double)best);
                v = r < clamp ? r : clamp;
            }
            distance[idx] = v;
        } else {
            dst[idx] = (unsigned short)(best < EDT_INF ? best : EDT_INF);
        }
    }
}

// 1-row frames (ny == 1 and nz == 1): the X pass is the whole transform
__global__ void __launch_bounds__(THREADS)
edt_finish_kernel(const unsigned short* __restrict__ src, long long n, float clamp, float* __restrict__ distance) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int best = src[idx];
        float v = clamp;
        if (best < EDT_INF) {
            const float r = (float)sqrt((double)best);
            v = r < clamp ? r : clamp;
        }
        distance[idx] = v;
    }
}

// ---- scale-normalised LoG response ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
log_response_kernel(const float* d0, const float* __restrict__ d1, const float* __restrict__ d2, long long n,
                    float sigma_sq, float* resp) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        float lap = d0[idx] + d1[idx];                 // generic_laplace: output = D_axis0; output += D_axis1; ...
        if (d2) lap = lap + d2[idx];
        float r = (-lap) * sigma_sq;
        if (r < 0.0f) r = 0.0f;
        resp[idx] = r;
    }
}

// ---- 3^d local maxima of one scale, best scale so far wins ------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
peak_update_kernel(const float* __restrict__ resp, const unsigned char* __restrict__ mask,
                   const float* __restrict__ distance, Dims d, float* __restrict__ best,
                   unsigned char* __restrict__ peak) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (!mask[idx] || !(distance[idx] > 0.0f)) continue;
        const float r = resp[idx];
        if (!(r > best[idx])) continue;
        const Pos p = pos_of(idx, d);
        const int z0 = p.z > 0 ? -1 : 0, z1 = p.z + 1 < d.nz ? 1 : 0;
        const int y0 = p.y > 0 ? -1 : 0, y1 = p.y + 1 < d.ny ? 1 : 0;
        const int x0 = p.x > 0 ? -1 : 0, x1 = p.x + 1 < d.nx ? 1 : 0;
        bool is_max = true;                            // mode="nearest": the window only repeats in-range voxels
        for (int dz = z0; dz <= z1 && is_max; ++dz)
            for (int dy = y0; dy <= y1 && is_max; ++dy) {
                const long long row = idx + dz * d.plane + (long long)dy * d.nx;
                for (int dx = x0; dx <= x1; ++dx)
                    if (resp[row + dx] > r) is_max = false;
            }
        if (is_max) {
            best[idx] = r;
            peak[idx] = 1;
        }
    }
}

// The same step without the response volume: the response of a voxel is three loads and three float32 operations, so the
// kernel evaluates it where it is needed — at the mask voxels, and at the 26 neighbours of those whose own response beats
// the best scale so far — instead of streaming a float32 volume out and in again (25 -> ~3 B/voxel for a 10 % mask).
__device__ __forceinline__ float response_at(const float* __restrict__ d0, const float* __restrict__ d1,
                                             const float* __restrict__ d2, long long idx, float sigma_sq) {
    float lap = d0[idx] + d1[idx];
    if (d2) lap = lap + d2[idx];
    float r = (-lap) * sigma_sq;
    if (r < 0.0f) r = 0.0f;
    return r;
}

__global__ void __launch_bounds__(THREADS)
peak_update_fused_kernel(const float* __restrict__ d0, const float* __restrict__ d1, const float* __restrict__ d2,
                         float sigma_sq, const unsigned char* __restrict__ mask, const float* __restrict__ distance, Dims d,
                         float* __restrict__ best, unsigned char* __restrict__ peak) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (!mask[idx] || !(distance[idx] > 0.0f)) continue;
        const float r = response_at(d0, d1, d2, idx, sigma_sq);
        if (!(r > best[idx])) continue;
        const Pos p = pos_of(idx, d);
        const int z0 = p.z > 0 ? -1 : 0, z1 = p.z + 1 < d.nz ? 1 : 0;
        const int y0 = p.y > 0 ? -1 : 0, y1 = p.y + 1 < d.ny ? 1 : 0;
        const int x0 = p.x > 0 ? -1 : 0, x1 = p.x + 1 < d.nx ? 1 : 0;
        bool is_max = true;
        for (int dz = z0; dz <= z1 && is_max; ++dz)
            for (int dy = y0; dy <= y1 && is_max; ++dy) {
                const long long row = idx + dz * d.plane + (long long)dy * d.nx;
                for (int dx = x0; dx <= x1; ++dx)
                    if (response_at(d0, d1, d2, row + dx, sigma_sq) > r) is_max = false;
            }
        if (is_max) {
            best[idx] = r;
            peak[idx] = 1;
        }
    }
}

// ---- non-maximum suppression on the raw intensity ---------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
nms_kernel(const unsigned char* __restrict__ peak, const float* __restrict__ intensity, Dims d, int radius,
           unsigned char* __restrict__ marker) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        unsigned char keep = 0;
        if (peak[idx]) {
            const float s = intensity[idx];
            if (s > 0.0f) {
                const Pos p = pos_of(idx, d);
                const int z0 = p.z - radius > 0 ? p.z - radius : 0, z1 = p.z + radius < d.nz - 1 ? p.z + radius : d.nz - 1;
                const int y0 = p.y - radius > 0 ? p.y - radius : 0, y1 = p.y + radius < d.ny - 1 ? p.y + radius : d.ny - 1;
                const int x0 = p.x - radius > 0 ? p.x - radius : 0, x1 = p.x + radius < d.nx - 1 ? p.x + radius : d.nx - 1;
                bool top = true;
                for (int z = z0; z <= z1 && top; ++z)
                    for (int y = y0; y <= y1 && top; ++y) {
                        const long long row = (long long)z * d.plane + (long long)y * d.nx;
                        for (int x = x0; x <= x1; ++x)
                            if (peak[row + x] && intensity[row + x] > s) top = false;
                    }
                keep = top ? 1 : 0;
            }
        }
        marker[idx] = keep;
    }
}

inline bool dims_ok(int nz, int ny, int nx) {
    return nz >= 1 && ny >= 1 && nx >= 1 && (long long)nz * ny * nx < (1ll << 40);
}

}  // namespace

extern "C" {

int nb200_markers_mask_border(const int* labels, int nz, int ny, int nx, unsigned char* mask, unsigned char* border,
                              void* stream) {
    NB_REQUIRE(labels && mask && border && dims_ok(nz, ny, nx), NB200_ERR_ARG, "nb200_markers_mask_border: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(mask_border_kernel, grid_of(d.total), THREADS, nb::as_stream(stream), labels, d, mask, border);
    return nb::check_launch("mask_border_kernel");
}

int nb200_markers_edt(const unsigned char* mask, int nz, int ny, int nx, int window, float clamp, unsigned short* scratch,
                      float* distance, void* stream) {
    NB_REQUIRE(mask && scratch && distance && dims_ok(nz, ny, nx), NB200_ERR_ARG, "nb200_markers_edt: bad argument");
    NB_REQUIRE(window >= 1 && window <= EDT_MAX_WINDOW, NB200_ERR_UNSUPPORTED,
               "nb200_markers_edt: window %d outside 1..%d", window, EDT_MAX_WINDOW);
    NB_REQUIRE(clamp >= 0.0f && clamp <= (float)window, NB200_ERR_ARG,
               "nb200_markers_edt: the window (%d) must cover the clamp (%g)", window, (double)clamp);
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    const unsigned grid = grid_of(d.total);
    unsigned short* a = scratch;
    unsigned short* b = scratch + d.total;
    NB_LAUNCH(edt_x_kernel, grid, THREADS, st, mask, d, window, a);
    if (nz > 1) {
        NB_LAUNCH(edt_axis_kernel<false>, grid, THREADS, st, a, d, 1, window, clamp, b, (float*)nullptr);
        NB_LAUNCH(edt_axis_kernel<true>, grid, THREADS, st, b, d, 0, window, clamp, (unsigned short*)nullptr, distance);
    } else if (ny > 1) {
        NB_LAUNCH(edt_axis_kernel<true>, grid, THREADS, st, a, d, 1, window, clamp, (unsigned short*)nullptr, distance);
    } else {
        NB_LAUNCH(edt_finish_kernel, grid, THREADS, st, a, d.total, clamp, distance);
    }
    return nb::check_launch("edt kernels");
}

int nb200_markers_log_response(const float* d0, const float* d1, const float* d2, long long n, float sigma_sq,
                               float* resp, void* stream) {
    NB_REQUIRE(d0 && d1 && resp && n >= 0, NB200_ERR_ARG, "nb200_markers_log_response: bad argument");
    if (n == 0) return NB200_OK;
    NB_LAUNCH(log_response_kernel, grid_of(n), THREADS, nb::as_stream(stream), d0, d1, d2, n, sigma_sq, resp);
    return nb::check_launch("log_response_kernel");
}

int nb200_markers_peak_update(const float* resp, const unsigned char* mask, const float* distance, int nz, int ny, int nx,
                              float* best, unsigned char* peak, void* stream) {
    NB_REQUIRE(resp && mask && distance && best && peak && dims_ok(nz, ny, nx), NB200_ERR_ARG,
               "nb200_markers_peak_update: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(peak_update_kernel, grid_of(d.total), THREADS, nb::as_stream(stream), resp, mask, distance, d, best, peak);
    return nb::check_launch("peak_update_kernel");
}

int nb200_markers_peak_update_fused(const float* d0, const float* d1, const float* d2, float sigma_sq,
                                    const unsigned char* mask, const float* distance, int nz, int ny, int nx, float* best,
                                    unsigned char* peak, void* stream) {
    NB_REQUIRE(d0 && d1 && mask && distance && best && peak && dims_ok(nz, ny, nx), NB200_ERR_ARG,
               "nb200_markers_peak_update_fused: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(peak_update_fused_kernel, grid_of(d.total), THREADS, nb::as_stream(stream), d0, d1, d2, sigma_sq, mask,
              distance, d, best, peak);
    return nb::check_launch("peak_update_fused_kernel");
}

int nb200_markers_nms(const unsigned char* peak, const float* intensity, int nz, int ny, int nx, int radius,
                      unsigned char* marker, void* stream) {
    NB_REQUIRE(peak && intensity && marker && dims_ok(nz, ny, nx), NB200_ERR_ARG, "nb200_markers_nms: bad argument");
    NB_REQUIRE(radius >= 0 && radius <= 64, NB200_ERR_UNSUPPORTED, "nb200_markers_nms: radius %d outside 0..64", radius);
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(nms_kernel, grid_of(d.total), THREADS, nb::as_stream(stream), peak, intensity, d, radius, marker);
    return nb::check_launch("nms_kernel");
}

}  // extern "C"
