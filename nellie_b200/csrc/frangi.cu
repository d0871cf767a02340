// F4-F9 — Hessian statistics (K2) and the fused Hessian + Frobenius mask + 3x3 eigenvalues +
// vesselness + max/AND accumulate kernel (K3).  Reference: nellie/segmentation/filtering.py
// :446-562 (_compute_hessian), :407-444 (_get_frob_mask), :651-767 (eigvalsh + _filter_hessian),
// :842-851 (max over sigma, AND of masks, skip of an empty sigma).
//
// Data layout: blurred frame g (Z,Y,X) float32, X contiguous; accumulator acc same shape:
//   acc >= 0  : running max of the vesselness over the sigmas processed so far (voxel alive)
//   acc == -1 : voxel failed the Frobenius mask at some non-skipped sigma (dead for good)
// so the reference's separate bool `masks` volume never exists.  Algorithmic HBM traffic per
// sigma: K2 reads g (4 B/voxel); K3 reads g, reads+writes acc (12 B/voxel).
//
// Tiling (both kernels): a CTA owns a (TZ x TY x TX) brick of outputs and stages the brick plus
// a 2-voxel halo of g in shared memory with coalesced 128-byte row loads; the 19-point stencil
// and all re-use then run out of shared memory, so g is read from HBM ~once (halo overhead only).
#include "common.cuh"
#include "hessian.cuh"

namespace {

constexpr int TX = 64, TY = 8, TZ = 8;      // outputs per CTA
constexpr int HALO = 2;
constexpr int SX = TX + 2 * HALO;           // 68
constexpr int SY = TY + 2 * HALO;           // 12
constexpr int SZ = TZ + 2 * HALO;           // 12
constexpr int SXP = SX + 1;                 // padded row pitch (floats)
constexpr int NTHREADS = 256;

struct Tile {
    float g[SZ][SY][SXP];
};

// stage brick + halo; coordinates outside the GLOBAL frame are clamped (their values are never
// used: the one-sided rules at the frame border only touch in-frame samples)
__device__ __forceinline__ void load_tile(Tile& t, const float* __restrict__ g, const nb200_vol& v,
                                          int zb0, int y0, int x0) {
    const long long plane = (long long)v.ny * v.nx;
    const int zlo = max(-v.zg_off, 0), zhi = min(v.nz_glob - v.zg_off, v.nz_buf) - 1;  // valid buffer planes
    for (int i = threadIdx.x; i < SZ * SY * SX; i += NTHREADS) {
        const int lx = i % SX;
        const int ly = (i / SX) % SY;
        const int lz = i / (SX * SY);
        int zb = zb0 - HALO + lz, y = y0 - HALO + ly, x = x0 - HALO + lx;
        zb = min(max(zb, zlo), zhi);
        y = min(max(y, 0), v.ny - 1);
        x = min(max(x, 0), v.nx - 1);
        t.g[lz][ly][lx] = __ldg(g + (long long)zb * plane + (long long)y * v.nx + x);
    }
}

struct TileLoad {
    const Tile* t;
    int lz, ly, lx;
    __device__ __forceinline__ float operator()(int dz, int dy, int dx) const {
        return t->g[lz + dz][ly + dy][lx + dx];
    }
};

__device__ __forceinline__ void brick_origin(const nb200_vol& v, int& zb0, int& y0, int& x0, bool& valid) {
    const int nbx = (v.nx + TX - 1) / TX, nby = (v.ny + TY - 1) / TY;
    long long b = blockIdx.x;
    const int bx = (int)(b % nbx); b /= nbx;
    const int by = (int)(b % nby); b /= nby;
    zb0 = v.zc0 + (int)b * TZ;
    y0 = by * TY;
    x0 = bx * TX;
    valid = zb0 < v.zc1;
}

// --------------------------------------------------------------------------------------------
// K2: max |component|, max frob_sq, sqrt(frob_sq) at the sampling lattice
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
hessian_stats_kernel(const float* __restrict__ g, nb200_vol v, nb::Spacing3 sp, int sz, int sy, int sx,
                     int g_first, int ly_n, int lx_n, float* __restrict__ frob_samples,
                     long long* __restrict__ hstats) {
    __shared__ Tile tile;
    __shared__ float red_a[NTHREADS / 32], red_f[NTHREADS / 32];
    int zb0, y0, x0; bool valid;
    brick_origin(v, zb0, y0, x0, valid);
    load_tile(tile, g, v, zb0, y0, x0);
    __syncthreads();
    const int n[3] = {v.nz_glob, v.ny, v.nx};
    float m_abs = 0.0f, m_frob = 0.0f;
    const int tx = threadIdx.x % TX, ty0 = threadIdx.x / TX;   // 64 x 4 threads
    for (int lz = 0; lz < TZ; ++lz) {
        const int zb = zb0 + lz;
        if (zb >= v.zc1) break;
        const int zg = zb + v.zg_off;
        for (int ly = ty0; ly < TY; ly += NTHREADS / TX) {
            const int y = y0 + ly, x = x0 + tx;
            if (y >= v.ny || x >= v.nx) continue;
            TileLoad L{&tile, lz + HALO, ly + HALO, tx + HALO};
            float a, b, c, d, e, f;
            nb::hessian3(L, zg, y, x, n, sp, a, b, c, d, e, f);
            const float fs = nb::frob_sq3(a, b, c, d, e, f);
            m_abs = fmaxf(m_abs, fmaxf(fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), fabsf(d))),
                                       fmaxf(fabsf(e), fabsf(f))));
            m_frob = fmaxf(m_frob, fs);
            if (frob_samples && (zg % sz == 0) && (y % sy == 0) && (x % sx == 0)) {
                const long long k = ((long long)((zg - g_first) / sz) * ly_n + y / sy) * lx_n + x / sx;
                frob_samples[k] = sqrtf(fs);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        m_abs = fmaxf(m_abs, __shfl_xor_sync(0xffffffffu, m_abs, o));
        m_frob = fmaxf(m_frob, __shfl_xor_sync(0xffffffffu, m_frob, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red_a[w] = m_abs; red_f[w] = m_frob; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < NTHREADS / 32; ++k) { m_abs = fmaxf(m_abs, red_a[k]); m_frob = fmaxf(m_frob, red_f[k]); }
        // non-negative floats order like their bit patterns
        atomicMax((unsigned long long*)&hstats[NB200_HS_MAX_ABS_BITS], (unsigned long long)nb::f2u(m_abs));
        atomicMax((unsigned long long*)&hstats[NB200_HS_MAX_FROBSQ_BITS], (unsigned long long)nb::f2u(m_frob));
    }
}

// --------------------------------------------------------------------------------------------
// K3: fused Hessian + mask + eig + vesselness + accumulate
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
frangi_accumulate_kernel(const float* __restrict__ g, float* __restrict__ acc, nb200_vol v, nb::Spacing3 sp,
                         float alpha_sq, float beta_sq, const double* __restrict__ spd) {
    __shared__ Tile tile;
    if (spd[NB200_SP_SKIP] != 0.0) return;        // empty mask: the sigma contributes nothing (:843-844)
    const float gamma_sq = (float)spd[NB200_SP_GAMMA_SQ];
    const float frob_cut = (float)spd[NB200_SP_FROB_CUT];
    const float max_abs = (float)spd[NB200_SP_MAX_ABS];
    int zb0, y0, x0; bool valid;
    brick_origin(v, zb0, y0, x0, valid);
    load_tile(tile, g, v, zb0, y0, x0);
    __syncthreads();
    const int n[3] = {v.nz_glob, v.ny, v.nx};
    const long long plane = (long long)v.ny * v.nx;
    const int tx = threadIdx.x % TX, ty0 = threadIdx.x / TX;
    for (int lz = 0; lz < TZ; ++lz) {
        const int zb = zb0 + lz;
        if (zb >= v.zc1) break;
        const int zg = zb + v.zg_off;
        for (int ly = ty0; ly < TY; ly += NTHREADS / TX) {
            const int y = y0 + ly, x = x0 + tx;
            if (y >= v.ny || x >= v.nx) continue;
            const long long idx = (long long)zb * plane + (long long)y * v.nx + x;
            const float prev = acc[idx];
            if (prev < 0.0f) continue;            // already dead: output is 0 whatever this sigma says
            TileLoad L{&tile, lz + HALO, ly + HALO, tx + HALO};
            float a, b, c, d, e, f;
            nb::hessian3(L, zg, y, x, n, sp, a, b, c, d, e, f);
            const float frob = sqrtf(nb::frob_sq3(a, b, c, d, e, f)) / max_abs;
            if (!(frob > frob_cut)) { acc[idx] = -1.0f; continue; }
            float l1, l2, l3;
            nb::eig3_sym<2>(a, b, c, d, e, f, l1, l2, l3);
            const float vv = nb::vesselness3(l1, l2, l3, alpha_sq, beta_sq, gamma_sq);
            if (vv > prev) acc[idx] = vv;
        }
    }
}

// --------------------------------------------------------------------------------------------
// 2-D variants (one thread per pixel, neighbours through L1; frames are small)
// --------------------------------------------------------------------------------------------
struct GlobalLoad2 {
    const float* p; int nx;
    __device__ __forceinline__ float operator()(int dy, int dx) const { return __ldg(p + dy * nx + dx); }
};

__global__ void __launch_bounds__(256)
hessian_stats_2d_kernel(const float* __restrict__ g, int ny, int nx, float h1y, float h2y, float h1x, float h2x,
                        int sy, int sx, int lx_n, float* __restrict__ frob_samples, long long* __restrict__ hstats) {
    __shared__ float red_a[8], red_f[8];
    const float h1[2] = {h1y, h1x}, h2[2] = {h2y, h2x};
    float m_abs = 0.0f, m_frob = 0.0f;
    const long long total = (long long)ny * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        GlobalLoad2 L{g + i, nx};
        float a, b, c;
        nb::hessian2(L, y, x, ny, nx, h1, h2, a, b, c);
        const float fs = nb::frob_sq2(a, b, c);
        m_abs = fmaxf(m_abs, fmaxf(fabsf(a), fmaxf(fabsf(b), fabsf(c))));
        m_frob = fmaxf(m_frob, fs);
        if (frob_samples && (y % sy == 0) && (x % sx == 0)) frob_samples[(long long)(y / sy) * lx_n + x / sx] = sqrtf(fs);
    }
    for (int o = 16; o > 0; o >>= 1) {
        m_abs = fmaxf(m_abs, __shfl_xor_sync(0xffffffffu, m_abs, o));
        m_frob = fmaxf(m_frob, __shfl_xor_sync(0xffffffffu, m_frob, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red_a[w] = m_abs; red_f[w] = m_frob; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { m_abs = fmaxf(m_abs, red_a[k]); m_frob = fmaxf(m_frob, red_f[k]); }
        atomicMax((unsigned long long*)&hstats[NB200_HS_MAX_ABS_BITS], (unsigned long long)nb::f2u(m_abs));
        atomicMax((unsigned long long*)&hstats[NB200_HS_MAX_FROBSQ_BITS], (unsigned long long)nb::f2u(m_frob));
    }
}

__global__ void __launch_bounds__(256)
frangi_accumulate_2d_kernel(const float* __restrict__ g, float* __restrict__ acc, int ny, int nx, float h1y,
                            float h2y, float h1x, float h2x, float beta_sq, const double* __restrict__ spd) {
    if (spd[NB200_SP_SKIP] != 0.0) return;
    const float gamma_sq = (float)spd[NB200_SP_GAMMA_SQ];
    const float frob_cut = (float)spd[NB200_SP_FROB_CUT];
    const float max_abs = (float)spd[NB200_SP_MAX_ABS];
    const float h1[2] = {h1y, h1x}, h2[2] = {h2y, h2x};
    const long long total = (long long)ny * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const float prev = acc[i];
        if (prev < 0.0f) continue;
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        GlobalLoad2 L{g + i, nx};
        float a, b, c;
        nb::hessian2(L, y, x, ny, nx, h1, h2, a, b, c);
        const float frob = sqrtf(nb::frob_sq2(a, b, c)) / max_abs;
        if (!(frob > frob_cut)) { acc[i] = -1.0f; continue; }
        float l1, l2;
        nb::eig2_sym(a, b, c, l1, l2);
        const float vv = nb::vesselness2(l1, l2, beta_sq, gamma_sq);
        if (vv > prev) acc[i] = vv;
    }
}

__global__ void hstats_reset_kernel(long long* hstats) {
    if (threadIdx.x < NB200_HS_WORDS) hstats[threadIdx.x] = 0;
}

int check_vol(const nb200_vol& v, const char* who) {
    NB_REQUIRE(v.ny >= 2 && v.nx >= 2 && v.nz_glob >= 2, NB200_ERR_ARG,
               "%s: every axis needs >= 2 samples (numpy.gradient)", who);
    NB_REQUIRE(v.zc0 >= 0 && v.zc1 <= v.nz_buf && v.zc0 <= v.zc1, NB200_ERR_ARG, "%s: bad Z window", who);
    // interior slab sides need 2 halo planes of g
    const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off;
    const int need_lo = g0 - 2 < 0 ? 0 : g0 - 2, need_hi = g1 + 1 >= v.nz_glob ? v.nz_glob - 1 : g1 + 1;
    NB_REQUIRE(g0 >= 0 && g1 <= v.nz_glob && need_lo - v.zg_off >= 0 && need_hi - v.zg_off < v.nz_buf,
               NB200_ERR_ARG, "%s: Z halo of 2 planes missing", who);
    return NB200_OK;
}

unsigned brick_grid(const nb200_vol& v) {
    const long long nbx = (v.nx + TX - 1) / TX, nby = (v.ny + TY - 1) / TY, nbz = (v.zc1 - v.zc0 + TZ - 1) / TZ;
    return (unsigned)(nbx * nby * nbz);
}

nb::Spacing3 spacing_from(const float* s) {
    nb::Spacing3 sp;
    for (int a = 0; a < 3; ++a) { sp.h1[a] = s[2 * a]; sp.h2[a] = s[2 * a + 1]; }
    return sp;
}

}  // namespace

extern "C" {

int nb200_hstats_reset(long long* hstats, void* stream) {
    NB_REQUIRE(hstats, NB200_ERR_ARG, "nb200_hstats_reset: null");
    hstats_reset_kernel<<<1, 32, 0, nb::as_stream(stream)>>>(hstats);
    return nb::check_launch("hstats_reset");
}

int nb200_hessian_stats(const float* gauss, const nb200_vol* vol, const float* spacing, int sz, int sy, int sx,
                        float* frob_samples, long long* hstats, void* stream) {
    NB_REQUIRE(gauss && vol && spacing && hstats && sz > 0 && sy > 0 && sx > 0, NB200_ERR_ARG,
               "nb200_hessian_stats: bad argument");
    const nb200_vol v = *vol;
    int rc = check_vol(v, "nb200_hessian_stats");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    const int g0 = v.zc0 + v.zg_off;
    const int g_first = ((g0 + sz - 1) / sz) * sz;
    const int ly_n = (v.ny + sy - 1) / sy, lx_n = (v.nx + sx - 1) / sx;
    hessian_stats_kernel<<<brick_grid(v), NTHREADS, 0, nb::as_stream(stream)>>>(
        gauss, v, spacing_from(spacing), sz, sy, sx, g_first, ly_n, lx_n, frob_samples, hstats);
    return nb::check_launch("hessian_stats");
}

int nb200_frangi_accumulate(const float* gauss, float* acc, const nb200_vol* vol, const float* spacing,
                            float alpha_sq, float beta_sq, const double* sp, void* stream) {
    NB_REQUIRE(gauss && acc && vol && spacing && sp, NB200_ERR_ARG, "nb200_frangi_accumulate: null argument");
    const nb200_vol v = *vol;
    int rc = check_vol(v, "nb200_frangi_accumulate");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    frangi_accumulate_kernel<<<brick_grid(v), NTHREADS, 0, nb::as_stream(stream)>>>(
        gauss, acc, v, spacing_from(spacing), alpha_sq, beta_sq, sp);
    return nb::check_launch("frangi_accumulate");
}

int nb200_hessian_stats_2d(const float* gauss, int ny, int nx, const float* spacing, int sy, int sx,
                           float* frob_samples, long long* hstats, void* stream) {
    NB_REQUIRE(gauss && spacing && hstats && ny >= 2 && nx >= 2 && sy > 0 && sx > 0, NB200_ERR_ARG,
               "nb200_hessian_stats_2d: bad argument");
    const long long total = (long long)ny * nx;
    hessian_stats_2d_kernel<<<nb::grid_for(total, 256, 4), 256, 0, nb::as_stream(stream)>>>(
        gauss, ny, nx, spacing[0], spacing[1], spacing[2], spacing[3], sy, sx, (nx + sx - 1) / sx, frob_samples, hstats);
    return nb::check_launch("hessian_stats_2d");
}

int nb200_frangi_accumulate_2d(const float* gauss, float* acc, int ny, int nx, const float* spacing, float beta_sq,
                               const double* sp, void* stream) {
    NB_REQUIRE(gauss && acc && spacing && sp && ny >= 2 && nx >= 2, NB200_ERR_ARG,
               "nb200_frangi_accumulate_2d: bad argument");
    const long long total = (long long)ny * nx;
    frangi_accumulate_2d_kernel<<<nb::grid_for(total, 256, 4), 256, 0, nb::as_stream(stream)>>>(
        gauss, acc, ny, nx, spacing[0], spacing[1], spacing[2], spacing[3], beta_sq, sp);
    return nb::check_launch("frangi_accumulate_2d");
}

}  // extern "C"
