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
// Both kernels are epilogues of the Z-marching Hessian in hessian_march.cuh (shared-memory ring of
// blurred planes fed by cp.async, first-derivative planes shared between the second derivatives,
// 4 voxels per thread with 128-bit shared/global accesses).
#include <stdlib.h>

#include "common.cuh"
#include "hessian.cuh"
#include "hessian_march.cuh"

namespace {

using hm::Hess4;

// --------------------------------------------------------------------------------------------
// epilogues
// --------------------------------------------------------------------------------------------
struct StatsParams {
    int sz, sy, sx, g_first, ly_n, lx_n;
    float* frob_samples;
    long long* hstats;
};

// frob_sq of 4 voxels, evaluated exactly like filtering.py:538-543 (each product and each sum rounded once).
// Squares run packed (FMUL2); the sums are scalar FADDs on purpose: ptxas contracts a packed multiply that
// feeds a packed add into FFMA2 even for mul.rn/add.rn.f32x2 and --fmad=false, which would skip a rounding.
__device__ __forceinline__ float4 frob_sq4(const Hess4& h) {
    auto half = [](float2 zz, float2 zy, float2 zx, float2 yy, float2 yx, float2 xx) -> float2 {
        const float2 a = __fmul2_rn(zz, zz), b = __fmul2_rn(yy, yy), c = __fmul2_rn(xx, xx);
        const float2 d = __fmul2_rn(zy, zy), e = __fmul2_rn(zx, zx), f = __fmul2_rn(yx, yx);
        float2 r;
        r.x = ((a.x + b.x) + c.x) + 2.0f * ((d.x + e.x) + f.x);
        r.y = ((a.y + b.y) + c.y) + 2.0f * ((d.y + e.y) + f.y);
        return r;
    };
    const float2 lo = half(make_float2(h.zz.x, h.zz.y), make_float2(h.zy.x, h.zy.y), make_float2(h.zx.x, h.zx.y),
                           make_float2(h.yy.x, h.yy.y), make_float2(h.yx.x, h.yx.y), make_float2(h.xx.x, h.xx.y));
    const float2 hi = half(make_float2(h.zz.z, h.zz.w), make_float2(h.zy.z, h.zy.w), make_float2(h.zx.z, h.zx.w),
                           make_float2(h.yy.z, h.yy.w), make_float2(h.yx.z, h.yx.w), make_float2(h.xx.z, h.xx.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float absmax4(const float4& a) {
    return fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
}

struct StatsEpi {
    const StatsParams& p;
    const nb200_vol& v;
    float m_abs = 0.0f, m_frob = 0.0f;
    int xmask = 0, xlat = 0;            // lattice columns inside this thread's 4-voxel group (loop invariant)
    int ylat[hm::TY / hm::NW];          // lattice row index of each output row, or -1
    long long zrow = -1;                // lattice plane offset of the current plane, or -1 (uniform)
    __device__ StatsEpi(const StatsParams& p_, const nb200_vol& v_, int x0, int y0, void*) : p(p_), v(v_) {
        const int x = x0 + 4 * (threadIdx.x & 31);
        for (int k = 0; k < 4; ++k) xmask |= ((x + k) % p.sx == 0) ? (1 << k) : 0;
        xlat = (x + p.sx - 1) / p.sx;
        for (int i = 0; i < hm::TY / hm::NW; ++i) {
            const int y = y0 + (threadIdx.x >> 5) + i * hm::NW;
            ylat[i] = (p.frob_samples && xmask && (y % p.sy == 0)) ? y / p.sy : -1;
        }
    }
    __device__ __forceinline__ void plane(int zg) {
        zrow = (zg % p.sz == 0) ? (long long)((zg - p.g_first) / p.sz) * p.ly_n : -1;
    }
    __device__ __forceinline__ void preload(int, int, int, int, int) {}
    __device__ __forceinline__ bool skip4(int, int, int, int, int) { return false; }
    __device__ __forceinline__ void voxels4(int row, bool active, int, int, int x, int nvalid, const Hess4& h) {
        if (!active) return;
        const float4 fs = frob_sq4(h);
        if (nvalid == 4) {
            m_abs = fmaxf(m_abs, fmaxf(fmaxf(fmaxf(absmax4(h.zz), absmax4(h.zy)), fmaxf(absmax4(h.zx), absmax4(h.yy))),
                                       fmaxf(absmax4(h.yx), absmax4(h.xx))));
            m_frob = fmaxf(m_frob, fmaxf(fmaxf(fs.x, fs.y), fmaxf(fs.z, fs.w)));
        } else {
            const float* c[6] = {&h.zz.x, &h.zy.x, &h.zx.x, &h.yy.x, &h.yx.x, &h.xx.x};
            const float* f = &fs.x;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nvalid) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) m_abs = fmaxf(m_abs, fabsf(c[j][k]));
                    m_frob = fmaxf(m_frob, f[k]);
                }
        }
        if (zrow >= 0 && ylat[row] >= 0) {            // lattice row: a few threads per plane
            const float* f = &fs.x;
            long long idx = (zrow + ylat[row]) * p.lx_n + xlat;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nvalid && (xmask >> k & 1)) p.frob_samples[idx++] = sqrtf(f[k]);
        }
    }
    __device__ void finish() {
        __shared__ float red_a[hm::NW], red_f[hm::NW];
        for (int o = 16; o > 0; o >>= 1) {
            m_abs = fmaxf(m_abs, __shfl_xor_sync(0xffffffffu, m_abs, o));
            m_frob = fmaxf(m_frob, __shfl_xor_sync(0xffffffffu, m_frob, o));
        }
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { red_a[w] = m_abs; red_f[w] = m_frob; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < hm::NW; ++k) { m_abs = fmaxf(m_abs, red_a[k]); m_frob = fmaxf(m_frob, red_f[k]); }
            // non-negative floats order like their bit patterns
            atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_ABS_BITS], (unsigned long long)nb::f2u(m_abs));
            atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_FROBSQ_BITS], (unsigned long long)nb::f2u(m_frob));
        }
    }
};

// test/diagnostic epilogue: writes the six second derivatives (order zz, zy, zx, yy, yx, xx) as six volumes
struct DumpParams {
    float* out;          // 6 x nz_buf x ny x nx
};
struct DumpEpi {
    const DumpParams& p;
    const nb200_vol& v;
    __device__ DumpEpi(const DumpParams& p_, const nb200_vol& v_, int, int, void*) : p(p_), v(v_) {}
    __device__ __forceinline__ void plane(int) {}
    __device__ __forceinline__ void preload(int, int, int, int, int) {}
    __device__ __forceinline__ bool skip4(int, int, int, int, int) { return false; }
    __device__ __forceinline__ void voxels4(int, bool active, int zb, int y, int x, int nvalid, const Hess4& h) {
        if (!active) return;
        const long long vol = (long long)v.nz_buf * v.ny * v.nx;
        const long long idx = ((long long)zb * v.ny + y) * v.nx + x;
        const float* c[6] = {&h.zz.x, &h.zy.x, &h.zx.x, &h.yy.x, &h.yx.x, &h.xx.x};
        for (int j = 0; j < 6; ++j)
            for (int k = 0; k < 4; ++k)
                if (k < nvalid) p.out[j * vol + idx + k] = c[j][k];
    }
    __device__ void finish() {}
};

// one out-of-line copy of the eigen-solver + vesselness (keeps the marching loop inside the I-cache)
__device__ __noinline__ float eig_vesselness(float a00, float a01, float a02, float a11, float a12, float a22,
                                             float alpha_sq, float beta_sq, float gamma_sq) {
    float l1, l2, l3;
    nb::eig3_sym<2>(a00, a01, a02, a11, a12, a22, l1, l2, l3);
    if (l3 > 0.0f || l2 > 0.0f) return 0.0f;           // filtering.py:759-761 zeroes these responses
    return nb::vesselness3(l1, l2, l3, alpha_sq, beta_sq, gamma_sq);
}

// two independent voxels per call: the long float64 dependency chains of the two solves interleave (ILP 2)
struct H6 {
    float a00, a01, a02, a11, a12, a22;
};
__device__ __noinline__ float2 eig_vesselness_pair(H6 a, H6 b, float alpha_sq, float beta_sq, float gamma_sq) {
    float l1, l2, l3, m1, m2, m3;
    nb::eig3_sym<2>(a.a00, a.a01, a.a02, a.a11, a.a12, a.a22, l1, l2, l3);
    nb::eig3_sym<2>(b.a00, b.a01, b.a02, b.a11, b.a12, b.a22, m1, m2, m3);
    float2 r = make_float2(0.0f, 0.0f);
    const bool need_a = !(l3 > 0.0f || l2 > 0.0f), need_b = !(m3 > 0.0f || m2 > 0.0f);   // filtering.py:759-761
    if (need_a) r.x = nb::vesselness3(l1, l2, l3, alpha_sq, beta_sq, gamma_sq);
    if (need_b) r.y = nb::vesselness3(m1, m2, m3, alpha_sq, beta_sq, gamma_sq);
    return r;
}

struct FrangiParams {
    float* acc;
    float alpha_sq, beta_sq;
    const double* spd;
    int acc_vec_ok;
    int debug;        // profiling experiments only (NB200_K3_DEBUG): 1 = skip the solves, 2 = eigenvalues only
};

// Per-warp work queue in shared memory.  The Hessian phase is dense (every live voxel), but only the
// voxels that pass the Frobenius mask need eigenvalues + vesselness (a few hundred instructions with
// float64 Newton steps): they are pushed here and popped 32 at a time, so that expensive part always
// runs with full warps instead of diverging on the speckled mask.
constexpr int QCAP = 96;      // entries per warp: < 64 left over + <= 32 pushed per round
constexpr int QWORDS = 7;     // six second derivatives + voxel index
struct FrangiQueue {
    float w[hm::NW][QWORDS][QCAP];
};

struct FrangiEpi {
    const FrangiParams& p;
    const nb200_vol& v;
    float gamma_sq, fs_min;
    long long plane_sz;
    float prev_[hm::TY / hm::NW][4];       // accumulator values of this thread's groups, one set per output row
    float next_[hm::TY / hm::NW][4];       // ... of the next plane, in flight
    long long idx_[hm::TY / hm::NW];
    float (*q)[QCAP];          // this warp's queue: q[word][entry]
    int fill = 0;              // warp-uniform
    int lane;
    __device__ FrangiEpi(const FrangiParams& p_, const nb200_vol& v_, int, int, void* queue_mem) : p(p_), v(v_) {
        gamma_sq = (float)p.spd[NB200_SP_GAMMA_SQ];
        fs_min = (float)p.spd[NB200_SP_FROBSQ_MIN];    // mask <=> frob_sq >= fs_min (see finalize_frob_kernel)
        plane_sz = (long long)v.ny * v.nx;
        q = reinterpret_cast<FrangiQueue*>(queue_mem)->w[threadIdx.x >> 5];
        lane = threadIdx.x & 31;
    }
    __device__ __forceinline__ void plane(int) {}
    // accumulator values of plane zb are requested one iteration ahead (preload) and consumed by skip4
    __device__ __forceinline__ void preload(int row, int zb, int y, int x, int nvalid) {
        float* nxt = next_[row];
        const long long idx = (long long)zb * plane_sz + (long long)y * v.nx + x;
        if (nvalid == 4 && p.acc_vec_ok) {
            const float4 a = *reinterpret_cast<const float4*>(p.acc + idx);
            nxt[0] = a.x; nxt[1] = a.y; nxt[2] = a.z; nxt[3] = a.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) nxt[k] = k < nvalid ? p.acc[idx + k] : -1.0f;
        }
    }
    __device__ __forceinline__ bool skip4(int row, int zb, int y, int x, int) {
        float* prev = prev_[row];
        idx_[row] = (long long)zb * plane_sz + (long long)y * v.nx + x;
#pragma unroll
        for (int k = 0; k < 4; ++k) prev[k] = next_[row][k];
        // dead voxels stay dead whatever this sigma says (AND of masks): skip their Hessian
        return prev[0] < 0.0f && prev[1] < 0.0f && prev[2] < 0.0f && prev[3] < 0.0f;
    }
    // pop `n` entries (n <= 64) from the top of the queue: two entries per lane, dense
    __device__ __forceinline__ void drain(int n) {
        __syncwarp();
        const int base = fill - n;
        const bool has_a = lane < n, has_b = lane + 32 < n;
        if (has_a) {
            const int ea = base + lane, eb = has_b ? ea + 32 : ea;
            H6 a, b;
            a.a00 = q[0][ea]; a.a01 = q[1][ea]; a.a02 = q[2][ea]; a.a11 = q[3][ea]; a.a12 = q[4][ea]; a.a22 = q[5][ea];
            b.a00 = q[0][eb]; b.a01 = q[1][eb]; b.a02 = q[2][eb]; b.a11 = q[3][eb]; b.a12 = q[4][eb]; b.a22 = q[5][eb];
            const int at_a = __float_as_int(q[6][ea]), at_b = __float_as_int(q[6][eb]);
            const float cur_a = p.acc[at_a], cur_b = p.acc[at_b];      // issued before the solves: latency hidden
            float2 vv = make_float2(0.0f, 0.0f);
            if (p.debug == 0) vv = eig_vesselness_pair(a, b, p.alpha_sq, p.beta_sq, gamma_sq);
            else if (p.debug == 2) {
                float l1, l2, l3, m1, m2, m3;
                nb::eig3_sym<2>(a.a00, a.a01, a.a02, a.a11, a.a12, a.a22, l1, l2, l3);
                nb::eig3_sym<2>(b.a00, b.a01, b.a02, b.a11, b.a12, b.a22, m1, m2, m3);
                vv = make_float2(fabsf(l3) * 1e-30f, fabsf(m3) * 1e-30f);
            }
            if (vv.x > cur_a) p.acc[at_a] = vv.x;                      // acc >= 0 here: a zero response changes nothing
            if (has_b && vv.y > cur_b) p.acc[at_b] = vv.y;
        }
        fill -= n;
        __syncwarp();
    }
    // called by ALL lanes of the warp (active = this lane has a group with a fresh Hessian)
    __device__ __forceinline__ void voxels4(int row, bool active, int, int, int, int nvalid, const Hess4& h) {
        const float* prev = prev_[row];
        const long long idx = idx_[row];
        bool pass[4] = {false, false, false, false};
        if (active) {
            const float4 fs4 = frob_sq4(h);
            const float* fs = &fs4.x;
            float out[4];
            bool killed = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                out[k] = prev[k];
                const bool alive = k < nvalid && prev[k] >= 0.0f;
                pass[k] = alive && fs[k] >= fs_min;
                if (alive && !pass[k]) { out[k] = -1.0f; killed = true; }
                // A non-zero response needs the two largest-|lambda| eigenvalues <= 0 and the third no larger
                // in magnitude, hence trace <= 0 (also for the float32-rounded eigenvalues the reference
                // tests).  A trace that is positive beyond its own rounding error therefore means V = 0:
                // the voxel stays alive and unchanged, no solve.
                const float dz = (&h.zz.x)[k], dy = (&h.yy.x)[k], dx = (&h.xx.x)[k];
                if (pass[k] && (dz + dy) + dx > 2.4e-7f * ((fabsf(dz) + fabsf(dy)) + fabsf(dx))) pass[k] = false;
            }
            if (killed) {
                if (nvalid == 4 && p.acc_vec_ok) {
                    *reinterpret_cast<float4*>(p.acc + idx) = make_float4(out[0], out[1], out[2], out[3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < nvalid) p.acc[idx + k] = out[k];
                }
            }
        }
        const float* c[6] = {&h.zz.x, &h.zy.x, &h.zx.x, &h.yy.x, &h.yx.x, &h.xx.x};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned bits = __ballot_sync(0xffffffffu, pass[k]);
            if (bits == 0u) continue;
            if (pass[k]) {
                const int e = fill + __popc(bits & ((1u << lane) - 1u));
#pragma unroll
                for (int j = 0; j < 6; ++j) q[j][e] = c[j][k];
                q[6][e] = __int_as_float((int)(idx + k));
            }
            fill += __popc(bits);
            if (fill >= 64) drain(64);
        }
    }
    __device__ void finish() {
        while (fill > 0) drain(fill < 64 ? fill : 64);
    }
};

// --------------------------------------------------------------------------------------------
// kernel wrapper: tile decode, division-mode / edge dispatch
// --------------------------------------------------------------------------------------------
template <int MODE, class Epi, class Params>
__global__ void __launch_bounds__(hm::NT, 2)
march_kernel(const float* __restrict__ g, nb200_vol v, hm::Divs dv, const double* __restrict__ flags, int run_if_unsafe,
             int zchunk, int check_skip, Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    hm::Smem& s = *reinterpret_cast<hm::Smem*>(smem_raw);
    if (flags != nullptr) {
        if (check_skip && flags[NB200_SP_SKIP] != 0.0) return;   // empty mask: the sigma contributes nothing (:843-844)
        // the fast-division launch and its IEEE twin are both enqueued; exactly one of them does the work
        const bool unsafe = flags[NB200_SP_UNSAFE] != 0.0;
        if (run_if_unsafe >= 0 && unsafe != (run_if_unsafe != 0)) return;
    }
    const int ntx = (v.nx + hm::TX - 1) / hm::TX, nty = (v.ny + hm::TY - 1) / hm::TY;
    long long b = blockIdx.x;
    const int bx = (int)(b % ntx); b /= ntx;
    const int by = (int)(b % nty); b /= nty;
    hm::Geo q;
    q.v = v;
    q.x0 = bx * hm::TX;
    q.y0 = by * hm::TY;
    q.plane = (long long)v.ny * v.nx;
    q.vec_ok = (q.x0 + hm::TX <= v.nx) && (v.nx % 4 == 0) && ((reinterpret_cast<unsigned long long>(g) & 15ull) == 0);
    const int zs = v.zc0 + v.zg_off + (int)b * zchunk;               // global planes
    const int ze = min(zs + zchunk, v.zc1 + v.zg_off);
    if (zs >= ze) return;
    Epi epi(p, v, q.x0, q.y0, smem_raw + sizeof(hm::Smem));
    const bool edge = q.x0 < 2 || q.x0 + hm::TX + 2 > v.nx || q.y0 < 2 || q.y0 + hm::TY + 2 > v.ny;
    if (edge) hm::march<MODE, true>(s, g, q, dv, zs, ze, epi);
    else hm::march<MODE, false>(s, g, q, dv, zs, ze, epi);
    epi.finish();
}

// exhaustive check of the reciprocal-multiply division against IEEE division for one divisor:
// every numerator with exponent in [-90, 90] plus +-0 (the range the FAST path is allowed to see)
__global__ void __launch_bounds__(256)
verify_divisor_kernel(float d, float r, unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    const unsigned lo_e = 127 - 90, hi_e = 127 + 90;
    const unsigned long long total = (unsigned long long)(hi_e - lo_e + 1) << 24;   // exponent x sign x mantissa
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned mant = (unsigned)(i & 0x7fffffu);
        const unsigned sign = (unsigned)((i >> 23) & 1u);
        const unsigned e = lo_e + (unsigned)(i >> 24);
        const float n = nb::u2f((sign << 31) | (e << 23) | mant);
        const float a = hm::divc<hm::DIV_FAST>(n, d, r);
        const float b = n / d;
        bad += (nb::f2u(a) != nb::f2u(b)) ? 1 : 0;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // +0 must map to +0.  (-0 maps to +0 instead of -0: the sign of a zero never reaches a result —
        // Hessian entries are squared, compared by magnitude, or added to non-zero terms.)
        const float z0 = hm::divc<hm::DIV_FAST>(0.0f, d, r), z1 = hm::divc<hm::DIV_FAST>(-0.0f, d, r);
        bad += (nb::f2u(z0) != nb::f2u(0.0f / d)) + (z1 != 0.0f);
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
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

// Z chunk length: enough CTAs for a few waves (2 CTAs/SM), chunks no shorter than 8 planes
int pick_zchunk(const nb200_vol& v, long long* n_ctas) {
    const long long tiles = (long long)((v.nx + hm::TX - 1) / hm::TX) * ((v.ny + hm::TY - 1) / hm::TY);
    const int nz = v.zc1 - v.zc0;
    const long long want = 8LL * nb::sm_count();
    long long chunks = (want + tiles - 1) / tiles;
    if (chunks < 1) chunks = 1;
    int zchunk = (int)((nz + chunks - 1) / chunks);
    if (zchunk < 8) zchunk = nz < 8 ? nz : 8;
    chunks = (nz + zchunk - 1) / zchunk;
    *n_ctas = tiles * chunks;
    return zchunk;
}

hm::Divs divs_from(const float* s) {
    hm::Divs dv;
    for (int a = 0; a < 3; ++a) {
        dv.a[a].d1 = s[2 * a];
        dv.a[a].r1 = 1.0f / s[2 * a];
        dv.a[a].d2 = s[2 * a + 1];
        dv.a[a].r2 = 1.0f / s[2 * a + 1];
    }
    return dv;
}

constexpr size_t kSmemBytes = sizeof(hm::Smem) + sizeof(FrangiQueue);

template <class K>
int set_smem(K kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) {
        nb::set_error("cudaFuncSetAttribute(smem=%zu): %s", kSmemBytes, cudaGetErrorString(e));
        return NB200_ERR_CUDA;
    }
    return NB200_OK;
}

template <int MODE, class Epi, class Params>
int launch_one(const float* g, const nb200_vol& v, const float* spacing, const double* sp, int run_if_unsafe,
               int check_skip, const Params& p, cudaStream_t st) {
    auto kernel = march_kernel<MODE, Epi, Params>;
    static bool smem_set = false;       // one flag per <MODE, Epi> instantiation
    if (!smem_set) {
        int rc = set_smem(kernel);
        if (rc) return rc;
        smem_set = true;
    }
    long long n_ctas = 0;
    const int zchunk = pick_zchunk(v, &n_ctas);
    kernel<<<(unsigned)n_ctas, hm::NT, kSmemBytes, st>>>(g, v, divs_from(spacing), sp, run_if_unsafe, zchunk,
                                                                check_skip, p);
    return NB200_OK;
}

// FAST mode enqueues the fast kernel and its IEEE twin; the device flag sp[UNSAFE] picks the one that runs
template <class Epi, class Params>
int launch_march(const float* g, const nb200_vol& v, const float* spacing, int div_mode, const double* sp,
                 int check_skip, const Params& p, cudaStream_t st, const char* what) {
    int rc;
    if (div_mode == hm::DIV_POW2) rc = launch_one<hm::DIV_POW2, Epi>(g, v, spacing, sp, -1, check_skip, p, st);
    else if (div_mode == hm::DIV_IEEE || sp == nullptr) rc = launch_one<hm::DIV_IEEE, Epi>(g, v, spacing, sp, -1, check_skip, p, st);
    else {
        rc = launch_one<hm::DIV_FAST, Epi>(g, v, spacing, sp, 0, check_skip, p, st);
        if (rc == NB200_OK) rc = launch_one<hm::DIV_IEEE, Epi>(g, v, spacing, sp, 1, check_skip, p, st);
    }
    if (rc) return rc;
    return nb::check_launch(what);
}

}  // namespace

extern "C" {

int nb200_hstats_reset(long long* hstats, void* stream) {
    NB_REQUIRE(hstats, NB200_ERR_ARG, "nb200_hstats_reset: null");
    hstats_reset_kernel<<<1, 32, 0, nb::as_stream(stream)>>>(hstats);
    return nb::check_launch("hstats_reset");
}

int nb200_divisor_mode(float d, int* mode_out, void* stream) {
    NB_REQUIRE(mode_out && d > 0.0f && d < INFINITY, NB200_ERR_ARG, "nb200_divisor_mode: bad divisor");
    int e = 0;
    const float m = frexpf(d, &e);
    if (m == 0.5f && e > -60 && e < 60) { *mode_out = hm::DIV_POW2; return NB200_OK; }
    cudaStream_t st = nb::as_stream(stream);
    unsigned long long* dev = nullptr;
    cudaError_t ce = cudaMalloc(&dev, sizeof(unsigned long long));   // init-time only, never on the frame path
    NB_REQUIRE(ce == cudaSuccess, NB200_ERR_OOM, "nb200_divisor_mode: out of memory");
    cudaMemsetAsync(dev, 0, sizeof(unsigned long long), st);
    verify_divisor_kernel<<<nb::sm_count() * 8, 256, 0, st>>>(d, 1.0f / d, dev);
    unsigned long long bad = 1;
    ce = cudaMemcpyAsync(&bad, dev, sizeof(bad), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(dev);
    NB_REQUIRE(ce == cudaSuccess, NB200_ERR_CUDA, "nb200_divisor_mode: %s", cudaGetErrorString(ce));
    *mode_out = bad == 0 ? hm::DIV_FAST : hm::DIV_IEEE;
    return NB200_OK;
}

int nb200_hessian_stats(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode, const double* sp,
                        int sz, int sy, int sx, float* frob_samples, long long* hstats, void* stream) {
    NB_REQUIRE(gauss && vol && spacing && hstats && sz > 0 && sy > 0 && sx > 0, NB200_ERR_ARG,
               "nb200_hessian_stats: bad argument");
    NB_REQUIRE(div_mode >= 0 && div_mode <= 2, NB200_ERR_ARG, "nb200_hessian_stats: div_mode %d", div_mode);
    const nb200_vol v = *vol;
    int rc = check_vol(v, "nb200_hessian_stats");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    StatsParams p;
    p.sz = sz; p.sy = sy; p.sx = sx;
    const int g0 = v.zc0 + v.zg_off;
    p.g_first = ((g0 + sz - 1) / sz) * sz;
    p.ly_n = (v.ny + sy - 1) / sy;
    p.lx_n = (v.nx + sx - 1) / sx;
    p.frob_samples = frob_samples;
    p.hstats = hstats;
    return launch_march<StatsEpi>(gauss, v, spacing, div_mode, sp, 0, p, nb::as_stream(stream), "hessian_stats");
}

int nb200_hessian_components(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode, float* out6,
                             void* stream) {
    NB_REQUIRE(gauss && vol && spacing && out6, NB200_ERR_ARG, "nb200_hessian_components: null argument");
    NB_REQUIRE(div_mode >= 0 && div_mode <= 2, NB200_ERR_ARG, "nb200_hessian_components: div_mode %d", div_mode);
    const nb200_vol v = *vol;
    int rc = check_vol(v, "nb200_hessian_components");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    DumpParams p;
    p.out = out6;
    cudaStream_t st = nb::as_stream(stream);
    if (div_mode == hm::DIV_POW2) rc = launch_one<hm::DIV_POW2, DumpEpi>(gauss, v, spacing, nullptr, -1, 0, p, st);
    else if (div_mode == hm::DIV_FAST) rc = launch_one<hm::DIV_FAST, DumpEpi>(gauss, v, spacing, nullptr, -1, 0, p, st);
    else rc = launch_one<hm::DIV_IEEE, DumpEpi>(gauss, v, spacing, nullptr, -1, 0, p, st);
    if (rc) return rc;
    return nb::check_launch("hessian_components");
}

int nb200_frangi_accumulate(const float* gauss, float* acc, const nb200_vol* vol, const float* spacing, int div_mode,
                            float alpha_sq, float beta_sq, const double* sp, void* stream) {
    NB_REQUIRE(gauss && acc && vol && spacing && sp, NB200_ERR_ARG, "nb200_frangi_accumulate: null argument");
    NB_REQUIRE(div_mode >= 0 && div_mode <= 2, NB200_ERR_ARG, "nb200_frangi_accumulate: div_mode %d", div_mode);
    const nb200_vol v = *vol;
    int rc = check_vol(v, "nb200_frangi_accumulate");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    FrangiParams p;
    p.acc = acc;
    p.alpha_sq = alpha_sq;
    p.beta_sq = beta_sq;
    p.spd = sp;
    p.acc_vec_ok = (v.nx % 4 == 0) && ((reinterpret_cast<unsigned long long>(acc) & 15ull) == 0);
    {
        const char* dbg = getenv("NB200_K3_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
    }
    return launch_march<FrangiEpi>(gauss, v, spacing, div_mode, sp, 1, p, nb::as_stream(stream), "frangi_accumulate");
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
