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
#include "march_host.cuh"

namespace {

using hm::Hess4;

// --------------------------------------------------------------------------------------------
// shared per-voxel pieces
// --------------------------------------------------------------------------------------------
// frob_sq of 4 voxels, evaluated exactly like filtering.py:538-543:
//   ((zz^2 + yy^2) + xx^2) + 2*((zy^2 + zx^2) + yx^2), every product and sum rounded once.
// Packed: squares are FMUL2; a sum a+b is FFMA2(a, 1, b) (the product a*1 is exact, so the result is RN(a+b));
// the last step FFMA2(2, s2, s1) = RN(2*s2 + s1) equals RN(RN(2*s2) + s1) because doubling is exact.
// The products are passed through an opaque asm so that ptxas cannot contract a square into the following
// addition (it does that for mul.rn.f32x2 feeding add.rn.f32x2, which would skip a rounding).
__device__ __forceinline__ float2 opaque2(float2 v) {
    unsigned long long u = (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
    asm volatile("" : "+l"(u));
    return make_float2(__uint_as_float((unsigned)u), __uint_as_float((unsigned)(u >> 32)));
}
__device__ __forceinline__ float2 frob_sq2(float2 zz, float2 zy, float2 zx, float2 yy, float2 yx, float2 xx) {
    const float2 a = opaque2(__fmul2_rn(zz, zz)), b = opaque2(__fmul2_rn(yy, yy)), c = opaque2(__fmul2_rn(xx, xx));
    const float2 d = opaque2(__fmul2_rn(zy, zy)), e = opaque2(__fmul2_rn(zx, zx)), f = opaque2(__fmul2_rn(yx, yx));
    const float2 s1 = hm::add2(hm::add2(a, b), c);
    const float2 s2 = hm::add2(hm::add2(d, e), f);
    return __ffma2_rn(make_float2(2.0f, 2.0f), s2, s1);
}
__device__ __forceinline__ float4 frob_sq4(const Hess4& h) {
    const float2 lo = frob_sq2(make_float2(h.zz.x, h.zz.y), make_float2(h.zy.x, h.zy.y), make_float2(h.zx.x, h.zx.y),
                               make_float2(h.yy.x, h.yy.y), make_float2(h.yx.x, h.yx.y), make_float2(h.xx.x, h.xx.y));
    const float2 hi = frob_sq2(make_float2(h.zz.z, h.zz.w), make_float2(h.zy.z, h.zy.w), make_float2(h.zx.z, h.zx.w),
                               make_float2(h.yy.z, h.yy.w), make_float2(h.yx.z, h.yx.w), make_float2(h.xx.z, h.xx.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float absmax4(const float4& a) {
    return fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
}

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

// voxel_code of 4 voxels on the packed pipes (same arithmetic: every sum / product rounded once)
__device__ __forceinline__ float2 voxel_code2(float2 fs, float2 zz, float2 yy, float2 xx) {
    const float2 a = hm::add2(zz, yy), b = hm::add2(zz, xx), c = hm::add2(yy, xx);
    const float2 m = make_float2(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y));
    const float2 mm = __fmul2_rn(m, m), lim = __fmul2_rn(make_float2(1.001e-10f, 1.001e-10f), fs);
    const bool z0 = m.x > 0.0f && mm.x > lim.x && fs.x > 1e-20f && fs.x < 1e20f;
    const bool z1 = m.y > 0.0f && mm.y > lim.y && fs.y > 1e-20f && fs.y < 1e20f;
    return make_float2(z0 ? __uint_as_float(__float_as_uint(fs.x) | 0x80000000u) : fs.x,
                       z1 ? __uint_as_float(__float_as_uint(fs.y) | 0x80000000u) : fs.y);
}
__device__ __forceinline__ float4 voxel_code4(const float4& fs, const Hess4& h) {
    const float2 lo = voxel_code2(make_float2(fs.x, fs.y), make_float2(h.zz.x, h.zz.y), make_float2(h.yy.x, h.yy.y),
                                  make_float2(h.xx.x, h.xx.y));
    const float2 hi = voxel_code2(make_float2(fs.z, fs.w), make_float2(h.zz.z, h.zz.w), make_float2(h.yy.z, h.yy.w),
                                  make_float2(h.xx.z, h.xx.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// --------------------------------------------------------------------------------------------
// epilogues of the interior march
// --------------------------------------------------------------------------------------------
struct StatsParams {
    int sz, sy, sx, g_first, ly_n, lx_n;
    float* frob_samples;
    long long* hstats;
    float* code;            // optional per-voxel record for nb200_frangi_sparse (see voxel_code)
    int code_vec_ok;
};

// Per-voxel record K2 leaves for the sparse K3:  |code| = frob_sq (exactly the value the mask tests), sign bit
// set when the vesselness is PROVABLY zero.  The proof is the diagonal test of devmath.cuh (pd_reject_diag) with
// the margin taken relative to the voxel's own Frobenius norm instead of the sigma-wide bound: a pair of diagonal
// entries with a_ii + a_jj > 1e-5 * F, F >= ||H||_F, forces lambda_2 > 0 or lambda_3 > 0 far beyond any
// rounding of the eigenvalues, so filtering.py:759-761 zeroes the response.  ||H||_F^2 is frob_sq itself; the
// comparison is done on squares (m > 0, m^2 > 1.001e-10 * frob_sq; the 1e-3 slack dwarfs the float32 rounding
// of frob_sq and m^2) and only for norms in the range pd_margins accepts, where nothing under- or overflows.
__device__ __forceinline__ float voxel_code(float fs, float zz, float yy, float xx) {
    const float m = fmaxf(fmaxf(zz + yy, zz + xx), yy + xx);
    const bool zero = m > 0.0f && m * m > 1.001e-10f * fs && fs > 1e-20f && fs < 1e20f;
    return zero ? __uint_as_float(__float_as_uint(fs) | 0x80000000u) : fs;
}

// range record of the blurred values (see NB200_HS_MIN_NZ_COMPL): NaN / inf count as "out of range"
__device__ __forceinline__ void store_range(long long* hstats, float g_min, float g_max) {
    if (g_min < INFINITY)
        atomicMax((unsigned long long*)&hstats[NB200_HS_MIN_NZ_COMPL], (unsigned long long)(0x7f800000u - nb::f2u(g_min)));
    if (!(g_max <= 3.0e38f)) g_max = INFINITY;     // NaN propagates as "unsafe"
    atomicMax((unsigned long long*)&hstats[NB200_HS_MAX_G_BITS], (unsigned long long)nb::f2u(g_max));
}

struct StatsEpi {
    const StatsParams& p;
    float m_abs = 0.0f, m_frob = 0.0f;
    float g_min = INFINITY, g_max = 0.0f;   // smallest non-zero and largest |blurred value| seen by this thread
    int xmask = 0, xlat = 0;            // lattice columns inside this thread's 4-voxel group (loop invariant)
    int ylat[2];                        // lattice row index of each output row, or -1
    long long zrow = -1;                // lattice plane offset of the current plane, or -1 (uniform)
    int zmod = -1, zq = 0;              // plane index modulo sz and lattice plane counter, kept incrementally
    __device__ StatsEpi(const StatsParams& p_, const nb200_vol&, int x0, int y0, void*) : p(p_) {
        const int x = x0 + 4 * (threadIdx.x & 31);
        for (int k = 0; k < 4; ++k) xmask |= ((x + k) % p.sx == 0) ? (1 << k) : 0;
        xlat = (x + p.sx - 1) / p.sx;
        for (int i = 0; i < 2; ++i) {
            const int y = y0 - 1 + 2 * (threadIdx.x >> 5) + i;
            ylat[i] = (p.frob_samples && xmask && y >= 0 && (y % p.sy == 0)) ? y / p.sy : -1;
        }
    }
    __device__ __forceinline__ void plane(int zg) {      // consecutive planes: one division per chunk, then counters
        if (zmod < 0) {
            zmod = zg % p.sz;
            zq = (zg - zmod - p.g_first) / p.sz;
        } else if (++zmod == p.sz) {
            zmod = 0;
            ++zq;
        }
        zrow = zmod == 0 ? (long long)zq * p.ly_n : -1;
    }
    __device__ __forceinline__ void prefetch(int, bool, long long) {}
    __device__ __forceinline__ void cta_sync_point() {}
    // range of the non-zero blurred values (zero-filled out-of-frame parts of a tile do not count)
    __device__ __forceinline__ void center(const float4& a, const float4& b) {
        const float v[8] = {fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w), fabsf(b.x), fabsf(b.y), fabsf(b.z), fabsf(b.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            g_max = fmaxf(g_max, v[k]);
            g_min = fminf(g_min, v[k] > 0.0f ? v[k] : INFINITY);
        }
    }
    __device__ __forceinline__ void voxels4(int row, bool valid, long long idx, const Hess4& h) {
        if (!valid) return;
        const float4 fs = frob_sq4(h);
        m_abs = fmaxf(m_abs, fmaxf(fmaxf(fmaxf(absmax4(h.zz), absmax4(h.zy)), fmaxf(absmax4(h.zx), absmax4(h.yy))),
                                   fmaxf(absmax4(h.yx), absmax4(h.xx))));
        m_frob = fmaxf(m_frob, fmaxf(fmaxf(fs.x, fs.y), fmaxf(fs.z, fs.w)));
        if (p.code) {
            const float4 c = voxel_code4(fs, h);
            if (p.code_vec_ok) *reinterpret_cast<float4*>(p.code + idx) = c;
            else { p.code[idx] = c.x; p.code[idx + 1] = c.y; p.code[idx + 2] = c.z; p.code[idx + 3] = c.w; }
        }
        if (zrow >= 0 && ylat[row] >= 0) {            // lattice row: a few threads per plane
            const float* f = &fs.x;
            long long idx = (zrow + ylat[row]) * p.lx_n + xlat;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (xmask >> k & 1) p.frob_samples[idx++] = sqrtf(f[k]);
        }
    }
    __device__ void finish() {
        for (int o = 16; o > 0; o >>= 1) {
            m_abs = fmaxf(m_abs, __shfl_xor_sync(0xffffffffu, m_abs, o));
            m_frob = fmaxf(m_frob, __shfl_xor_sync(0xffffffffu, m_frob, o));
            g_min = fminf(g_min, __shfl_xor_sync(0xffffffffu, g_min, o));
            g_max = fmaxf(g_max, __shfl_xor_sync(0xffffffffu, g_max, o));
        }
        if ((threadIdx.x & 31) == 0) {
            // non-negative floats order like their bit patterns
            atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_ABS_BITS], (unsigned long long)nb::f2u(m_abs));
            atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_FROBSQ_BITS], (unsigned long long)nb::f2u(m_frob));
            store_range(p.hstats, g_min, g_max);
        }
    }
};

// test/diagnostic epilogue: writes the six second derivatives (order zz, zy, zx, yy, yx, xx) as six volumes
struct DumpParams {
    float* out;          // 6 x nz_buf x ny x nx
};
struct DumpEpi {
    const DumpParams& p;
    long long vol;
    __device__ DumpEpi(const DumpParams& p_, const nb200_vol& v, int, int, void*)
        : p(p_), vol((long long)v.nz_buf * v.ny * v.nx) {}
    __device__ __forceinline__ void plane(int) {}
    __device__ __forceinline__ void prefetch(int, bool, long long) {}
    __device__ __forceinline__ void cta_sync_point() {}
    __device__ __forceinline__ void center(const float4&, const float4&) {}
    __device__ __forceinline__ void voxels4(int, bool valid, long long idx, const Hess4& h) {
        if (!valid) return;
        const float* c[6] = {&h.zz.x, &h.zy.x, &h.zx.x, &h.yy.x, &h.yx.x, &h.xx.x};
        for (int j = 0; j < 6; ++j)
            for (int k = 0; k < 4; ++k) p.out[j * vol + idx + k] = c[j][k];
    }
    __device__ void finish() {}
};

struct FrangiParams {
    float* acc;
    float alpha_sq, beta_sq;
    const double* spd;
    int acc_vec_ok;
    int debug;        // profiling experiments only (NB200_K3_DEBUG): 1 = skip the solves
};

// Per-sigma constants of the "response is provably zero" tests (devmath.cuh, pd_reject_*): margins relative
// to F = sqrt(6) * max|H| >= ||H||_F of every voxel.
struct ZeroTests {
    float tau1, tau2, tau3;
};
__device__ __forceinline__ ZeroTests zero_tests_from(const double* spd) {
    ZeroTests z;
    nb::pd_margins((float)spd[NB200_SP_MAX_ABS], z.tau1, z.tau2, z.tau3);
    return z;
}

// Work queues in shared memory.  The Hessian phase is dense (every voxel), but only voxels that pass the
// Frobenius mask AND are not provably zero need eigenvalues + vesselness (a few hundred instructions with
// float64 Newton steps).
//   RAW   (per warp)  voxels that passed the cheap in-lane tests; drained 32 at a time through the full
//                     positive-definiteness test by the warp that filled it (cheap, ~40 instructions);
//   READY (per CTA)   survivors of all warps.  At the end of a plane, when READY holds >= 512 entries, ALL 256
//                     threads solve two entries each.  The expensive part therefore never diverges on the
//                     speckled mask and never leaves seven warps waiting at a barrier for the eighth.
// A warp that finds READY full solves its survivors itself (dense regions, where every warp is busy anyway).
constexpr int RAW_CAP = 64;      // < 32 left over + <= 32 pushed per round
constexpr int RDY_CAP = 768;
constexpr int SOLVE_AT = 2 * hm::NT;
constexpr int QWORDS = 7;        // six second derivatives + voxel index
struct FrangiQueue {
    float raw[hm::NW][QWORDS][RAW_CAP];
    float rdy[QWORDS][RDY_CAP];
    int n_rdy;
    int pad[3];
};

struct FrangiEpi {
    const FrangiParams& p;
    float gamma_sq, fs_min;
    ZeroTests zt;
    float4 cur[2];             // accumulator values of this thread's two groups (current plane)
    FrangiQueue* q;
    float (*raw)[RAW_CAP];     // this warp's RAW queue: raw[word][entry]
    int n_raw = 0;             // warp-uniform
    int lane;
    __device__ FrangiEpi(const FrangiParams& p_, const nb200_vol&, int, int, void* queue_mem) : p(p_) {
        gamma_sq = (float)p.spd[NB200_SP_GAMMA_SQ];
        fs_min = (float)p.spd[NB200_SP_FROBSQ_MIN];    // mask <=> frob_sq >= fs_min (see finalize_frob_kernel)
        zt = zero_tests_from(p.spd);
        q = reinterpret_cast<FrangiQueue*>(queue_mem);
        raw = q->raw[threadIdx.x >> 5];
        lane = threadIdx.x & 31;
        if (threadIdx.x == 0) q->n_rdy = 0;            // visible after the first barrier of the march
    }
    __device__ __forceinline__ void plane(int) {}
    __device__ __forceinline__ void center(const float4&, const float4&) {}
    __device__ __forceinline__ void prefetch(int i, bool inb, long long idx) {
        float4 a = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
        if (inb) {
            if (p.acc_vec_ok) a = *reinterpret_cast<const float4*>(p.acc + idx);
            else a = make_float4(p.acc[idx], p.acc[idx + 1], p.acc[idx + 2], p.acc[idx + 3]);
        }
        cur[i] = a;
    }
    __device__ __forceinline__ void accumulate(int at, float cur_v, float v) {
        if (v > cur_v) p.acc[at] = v;                  // acc >= 0 here: a zero response changes nothing
    }
    // all threads of the CTA (call sites are CTA-uniform, right after a __syncthreads): solve READY[n-m, n)
    __device__ __forceinline__ void solve_ready(int n, int m) {
        const int ea = n - m + (int)threadIdx.x, eb = ea + hm::NT;
        const bool has_a = ea < n, has_b = eb < n;
        if (has_a) {
            const int ib = has_b ? eb : ea;
            H6 a, b;
            a.a00 = q->rdy[0][ea]; a.a01 = q->rdy[1][ea]; a.a02 = q->rdy[2][ea]; a.a11 = q->rdy[3][ea]; a.a12 = q->rdy[4][ea]; a.a22 = q->rdy[5][ea];
            b.a00 = q->rdy[0][ib]; b.a01 = q->rdy[1][ib]; b.a02 = q->rdy[2][ib]; b.a11 = q->rdy[3][ib]; b.a12 = q->rdy[4][ib]; b.a22 = q->rdy[5][ib];
            const int at_a = __float_as_int(q->rdy[6][ea]), at_b = __float_as_int(q->rdy[6][ib]);
            const float cur_a = p.acc[at_a], cur_b = p.acc[at_b];      // issued before the solves: latency hidden
            float2 vv = make_float2(0.0f, 0.0f);
            if (p.debug == 0) vv = eig_vesselness_pair(a, b, p.alpha_sq, p.beta_sq, gamma_sq);
            accumulate(at_a, cur_a, vv.x);
            if (has_b) accumulate(at_b, cur_b, vv.y);
        }
        __syncthreads();                               // every thread has read n and its entries
        if (threadIdx.x == 0) q->n_rdy = n - m;        // ordered before the next pushes by the march's barriers
    }
    // CTA-uniform point at the end of every plane (after the barrier): no pushes are in flight
    __device__ __forceinline__ void cta_sync_point() {
        const int n = *reinterpret_cast<volatile int*>(&q->n_rdy);
        if (n >= SOLVE_AT) solve_ready(n, SOLVE_AT);
    }
    // pop up to 32 entries from the top of RAW, keep those whose response is not provably zero
    __device__ __forceinline__ void drain_raw() {
        __syncwarp();
        const int n = n_raw < 32 ? n_raw : 32;
        const int e = n_raw - n + lane;
        bool keep = false;
        float w[QWORDS];
        if (lane < n) {
#pragma unroll
            for (int j = 0; j < QWORDS; ++j) w[j] = raw[j][e];
            keep = !nb::pd_reject_full(w[0], w[1], w[2], w[3], w[4], w[5], zt.tau2, zt.tau3);
        }
        n_raw -= n;
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if (bits != 0u) {
            const int cnt = __popc(bits);
            int base = 0;
            if (lane == 0) {                           // reserve cnt slots of READY, or learn that it is full
                int old = *reinterpret_cast<volatile int*>(&q->n_rdy);
                while (true) {
                    if (old + cnt > RDY_CAP) { old = -1; break; }
                    const int prev = atomicCAS(&q->n_rdy, old, old + cnt);
                    if (prev == old) break;
                    old = prev;
                }
                base = old;
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                if (base >= 0) {
                    const int o = base + __popc(bits & ((1u << lane) - 1u));
#pragma unroll
                    for (int j = 0; j < QWORDS; ++j) q->rdy[j][o] = w[j];
                } else {
                    const int at = __float_as_int(w[6]);
                    const float cur_v = p.acc[at];
                    float v = 0.0f;
                    if (p.debug == 0) v = eig_vesselness(w[0], w[1], w[2], w[3], w[4], w[5], p.alpha_sq, p.beta_sq, gamma_sq);
                    accumulate(at, cur_v, v);
                }
            }
        }
        __syncwarp();
    }
    // called by ALL lanes of the warp (valid = this lane's group belongs to the interior)
    __device__ __forceinline__ void voxels4(int i, bool valid, long long idx, const Hess4& h) {
        unsigned cand = 0;
        const float4 pv = cur[i];
        if (valid && (pv.x >= 0.0f || pv.y >= 0.0f || pv.z >= 0.0f || pv.w >= 0.0f)) {
            const float4 fs = frob_sq4(h);
            const bool a0 = pv.x >= 0.0f, a1 = pv.y >= 0.0f, a2 = pv.z >= 0.0f, a3 = pv.w >= 0.0f;
            const bool p0 = a0 && fs.x >= fs_min, p1 = a1 && fs.y >= fs_min, p2 = a2 && fs.z >= fs_min,
                       p3 = a3 && fs.w >= fs_min;
            // dead voxels stay dead (AND of masks); voxels failing this sigma's mask die now
            if ((a0 && !p0) || (a1 && !p1) || (a2 && !p2) || (a3 && !p3)) {
                const float4 o = make_float4(p0 ? pv.x : -1.0f, p1 ? pv.y : -1.0f, p2 ? pv.z : -1.0f, p3 ? pv.w : -1.0f);
                if (p.acc_vec_ok) *reinterpret_cast<float4*>(p.acc + idx) = o;
                else { p.acc[idx] = o.x; p.acc[idx + 1] = o.y; p.acc[idx + 2] = o.z; p.acc[idx + 3] = o.w; }
            }
            // cheap zero test: a pair of diagonal entries with a clearly positive sum (devmath.cuh)
            const float2 s_lo_a = hm::add2(make_float2(h.zz.x, h.zz.y), make_float2(h.yy.x, h.yy.y));
            const float2 s_lo_b = hm::add2(make_float2(h.zz.x, h.zz.y), make_float2(h.xx.x, h.xx.y));
            const float2 s_lo_c = hm::add2(make_float2(h.yy.x, h.yy.y), make_float2(h.xx.x, h.xx.y));
            const float2 s_hi_a = hm::add2(make_float2(h.zz.z, h.zz.w), make_float2(h.yy.z, h.yy.w));
            const float2 s_hi_b = hm::add2(make_float2(h.zz.z, h.zz.w), make_float2(h.xx.z, h.xx.w));
            const float2 s_hi_c = hm::add2(make_float2(h.yy.z, h.yy.w), make_float2(h.xx.z, h.xx.w));
            const float m0 = fmaxf(fmaxf(s_lo_a.x, s_lo_b.x), s_lo_c.x), m1 = fmaxf(fmaxf(s_lo_a.y, s_lo_b.y), s_lo_c.y);
            const float m2 = fmaxf(fmaxf(s_hi_a.x, s_hi_b.x), s_hi_c.x), m3 = fmaxf(fmaxf(s_hi_a.y, s_hi_b.y), s_hi_c.y);
            cand = (p0 && !(m0 > zt.tau1) ? 1u : 0u) | (p1 && !(m1 > zt.tau1) ? 2u : 0u) |
                   (p2 && !(m2 > zt.tau1) ? 4u : 0u) | (p3 && !(m3 > zt.tau1) ? 8u : 0u);
        }
        if (!__any_sync(0xffffffffu, cand != 0u)) return;
        const float* c[6] = {&h.zz.x, &h.zy.x, &h.zx.x, &h.yy.x, &h.yx.x, &h.xx.x};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool mine = (cand >> k) & 1u;
            const unsigned bits = __ballot_sync(0xffffffffu, mine);
            if (bits == 0u) continue;
            if (mine) {
                const int e = n_raw + __popc(bits & ((1u << lane) - 1u));
#pragma unroll
                for (int j = 0; j < 6; ++j) raw[j][e] = c[j][k];
                raw[6][e] = __int_as_float((int)(idx + k));
            }
            n_raw += __popc(bits);
            if (n_raw >= 32) drain_raw();
        }
    }
    __device__ void finish() {                         // called by every thread of the CTA
        while (n_raw > 0) drain_raw();
        __syncthreads();
        while (true) {
            const int n = *reinterpret_cast<volatile int*>(&q->n_rdy);
            if (n <= 0) break;
            solve_ready(n, n < SOLVE_AT ? n : SOLVE_AT);
            __syncthreads();
        }
    }
};

// --------------------------------------------------------------------------------------------
// interior kernel: tile decode, division-mode dispatch
// --------------------------------------------------------------------------------------------
template <int MODE, class Epi, class Params>
__global__ void __launch_bounds__(hm::NT, 2)
march_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ g, nb200_vol v, hm::Divs dv,
             const double* __restrict__ flags, int run_if_unsafe, int zi0, int zi1, int zchunk, int check_skip,
             int use_tma, Params p) {
    // no static __shared__ in this kernel: the dynamic window starts at the (1 KB aligned) base of shared memory,
    // which the TMA destinations (128-byte aligned) rely on; keeping the address space visible to the compiler
    // is what turns every access below into LDS/STS with immediate offsets
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (flags != nullptr) {
        if (check_skip && flags[NB200_SP_SKIP] != 0.0) return;   // empty mask: the sigma contributes nothing (:843-844)
        // the fast-division launch and its IEEE twin are both enqueued; exactly one of them does the work.
        // run_if_unsafe: -1 always, 0 / 1 = only when sp[UNSAFE] is clear / set, 2 = only when the emptiness of the
        // Frobenius mask is still undecided (sp[AMBIG] set and no exact pass has run: UNSAFE clear)
        const bool unsafe = flags[NB200_SP_UNSAFE] != 0.0;
        if (run_if_unsafe == 2) {
            if (unsafe || flags[NB200_SP_AMBIG] == 0.0) return;
        } else if (run_if_unsafe >= 0 && unsafe != (run_if_unsafe != 0)) return;
    }
    hm::Smem& s = *reinterpret_cast<hm::Smem*>(smem_raw);
    const int ntx = (v.nx + hm::TX - 1) / hm::TX, nty = (v.ny - 4 + hm::TYO - 1) / hm::TYO;
    long long b = blockIdx.x;
    const int bx = (int)(b % ntx); b /= ntx;
    const int by = (int)(b % nty); b /= nty;
    hm::Geo q;
    q.v = v;
    q.x0 = bx * hm::TX;
    q.y0 = 2 + by * hm::TYO;
    q.tma = use_tma != 0;
    const int zs = zi0 + (int)b * zchunk;               // global planes, interior only
    const int ze = min(zs + zchunk, zi1);
    if (zs >= ze) return;
    Epi epi(p, v, q.x0, q.y0, smem_raw + sizeof(hm::Smem));
    hm::march<MODE>(s, &map, g, q, dv, zs, ze, epi);
    epi.finish();
}

// --------------------------------------------------------------------------------------------
// border shell: generic per-voxel evaluation with the one-sided rules of numpy.gradient (hessian.cuh)
// --------------------------------------------------------------------------------------------
struct Shell {
    nb200_vol v;
    int xhi;                 // interior x range is [4, xhi)
    int zlo_n, zhi0, zhi_n;  // z-border planes (buffer coords): [zc0, zc0+zlo_n) and [zhi0, zhi0+zhi_n)
    int zi0, zi1;            // interior planes (buffer coords)
    int yb_lo, yb_hi0;       // y-border rows: [0, yb_lo) and [yb_hi0, ny)
    int xb_lo, xb_hi0;       // x-border columns: [0, xb_lo) and [xb_hi0, nx)
    long long n1, n2, n3;    // voxel counts of the three regions
};

__host__ Shell make_shell(const nb200_vol& v) {
    Shell sh;
    sh.v = v;
    sh.xhi = 4 * ((v.nx - 2) / 4);
    const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off, n = v.nz_glob;
    const int lo_end = min(g1, max(g0, 2));                       // global planes [g0, lo_end) have z < 2
    const int hi_beg = max(lo_end, min(g1, max(n - 2, 2)));       // global planes [hi_beg, g1) have z >= n-2
    sh.zlo_n = lo_end - g0;
    sh.zhi0 = hi_beg - v.zg_off;
    sh.zhi_n = g1 - hi_beg;
    sh.zi0 = lo_end - v.zg_off;
    sh.zi1 = hi_beg - v.zg_off;
    sh.yb_lo = min(2, v.ny);
    sh.yb_hi0 = max(sh.yb_lo, v.ny - 2);
    sh.xb_lo = min(4, v.nx);
    sh.xb_hi0 = max(sh.xb_lo, min(v.nx, max(sh.xhi, 4)));
    const long long nzi = max(0, sh.zi1 - sh.zi0);
    const long long nyb = sh.yb_lo + (v.ny - sh.yb_hi0), nyi = v.ny - nyb;
    const long long nxb = sh.xb_lo + (v.nx - sh.xb_hi0);
    sh.n1 = (long long)(sh.zlo_n + sh.zhi_n) * v.ny * v.nx;
    sh.n2 = nzi * nyb * v.nx;
    sh.n3 = nzi * nyi * nxb;
    return sh;
}

// i-th shell voxel -> (buffer plane, y, x)
__device__ __forceinline__ void shell_voxel(const Shell& sh, long long i, int& zb, int& y, int& x) {
    const int nx = sh.v.nx, ny = sh.v.ny;
    if (i < sh.n1) {
        const long long pl = (long long)ny * nx;
        const int j = (int)(i / pl);
        const long long r = i - (long long)j * pl;
        zb = j < sh.zlo_n ? sh.v.zc0 + j : sh.zhi0 + (j - sh.zlo_n);
        y = (int)(r / nx);
        x = (int)(r - (long long)y * nx);
        return;
    }
    i -= sh.n1;
    if (i < sh.n2) {
        const int nyb = sh.yb_lo + (ny - sh.yb_hi0);
        const long long per = (long long)nyb * nx;
        const int j = (int)(i / per);
        const long long r = i - (long long)j * per;
        const int yy = (int)(r / nx);
        zb = sh.zi0 + j;
        y = yy < sh.yb_lo ? yy : sh.yb_hi0 + (yy - sh.yb_lo);
        x = (int)(r - (long long)yy * nx);
        return;
    }
    i -= sh.n2;
    const int nxb = sh.xb_lo + (nx - sh.xb_hi0);
    const int nyi = sh.yb_hi0 - sh.yb_lo;
    const long long per = (long long)nyi * nxb;
    const int j = (int)(i / per);
    const long long r = i - (long long)j * per;
    const int yy = (int)(r / nxb);
    const int xx = (int)(r - (long long)yy * nxb);
    zb = sh.zi0 + j;
    y = sh.yb_lo + yy;
    x = xx < sh.xb_lo ? xx : sh.xb_hi0 + (xx - sh.xb_lo);
}

struct GlobalLoad3 {
    const float* p; long long plane; int nx;
    __device__ __forceinline__ float operator()(int dz, int dy, int dx) const {
        return __ldg(p + dz * plane + (long long)dy * nx + dx);
    }
};

struct ShellHessian {
    float zz, zy, zx, yy, yx, xx;
};
__device__ __forceinline__ ShellHessian shell_hessian(const float* __restrict__ g, const Shell& sh, const nb::Spacing3& sp,
                                                      int zb, int y, int x, long long& idx) {
    const long long plane = (long long)sh.v.ny * sh.v.nx;
    idx = (long long)zb * plane + (long long)y * sh.v.nx + x;
    GlobalLoad3 L{g + idx, plane, sh.v.nx};
    const int n[3] = {sh.v.nz_glob, sh.v.ny, sh.v.nx};
    ShellHessian h;
    nb::hessian3(L, zb + sh.v.zg_off, y, x, n, sp, h.zz, h.zy, h.zx, h.yy, h.yx, h.xx);
    return h;
}

__global__ void __launch_bounds__(256)
shell_stats_kernel(const float* __restrict__ g, Shell sh, nb::Spacing3 sp, StatsParams p, const double* spd, int gate) {
    // gate: -1 always; 1 = redo pass (only when sp[UNSAFE]); 2 = only when sp[AMBIG] and not sp[UNSAFE]
    if (gate == 1 && spd[NB200_SP_UNSAFE] == 0.0) return;
    if (gate == 2 && (spd[NB200_SP_UNSAFE] != 0.0 || spd[NB200_SP_AMBIG] == 0.0)) return;
    float m_abs = 0.0f, m_frob = 0.0f, g_min = INFINITY, g_max = 0.0f;
    const long long total = sh.n1 + sh.n2 + sh.n3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int zb, y, x;
        shell_voxel(sh, i, zb, y, x);
        long long idx;
        const ShellHessian h = shell_hessian(g, sh, sp, zb, y, x, idx);
        const float fs = nb::frob_sq3(h.zz, h.zy, h.zx, h.yy, h.yx, h.xx);
        m_abs = fmaxf(m_abs, fmaxf(fmaxf(fmaxf(fabsf(h.zz), fabsf(h.zy)), fmaxf(fabsf(h.zx), fabsf(h.yy))),
                                   fmaxf(fabsf(h.yx), fabsf(h.xx))));
        m_frob = fmaxf(m_frob, fs);
        {
            const float c = fabsf(__ldg(g + idx));
            g_max = fmaxf(g_max, c);
            g_min = fminf(g_min, c > 0.0f ? c : INFINITY);
        }
        if (p.code) p.code[idx] = voxel_code(fs, h.zz, h.yy, h.xx);
        const int zg = zb + sh.v.zg_off;
        if (p.frob_samples && zg % p.sz == 0 && y % p.sy == 0 && x % p.sx == 0)
            p.frob_samples[((long long)((zg - p.g_first) / p.sz) * p.ly_n + y / p.sy) * p.lx_n + x / p.sx] = sqrtf(fs);
    }
    for (int o = 16; o > 0; o >>= 1) {
        m_abs = fmaxf(m_abs, __shfl_xor_sync(0xffffffffu, m_abs, o));
        m_frob = fmaxf(m_frob, __shfl_xor_sync(0xffffffffu, m_frob, o));
        g_min = fminf(g_min, __shfl_xor_sync(0xffffffffu, g_min, o));
        g_max = fmaxf(g_max, __shfl_xor_sync(0xffffffffu, g_max, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_ABS_BITS], (unsigned long long)nb::f2u(m_abs));
        atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_FROBSQ_BITS], (unsigned long long)nb::f2u(m_frob));
        store_range(p.hstats, g_min, g_max);
    }
}

// redo pass: forget the maxima of the unsafe fast pass (the range words stay)
__global__ void hstats_redo_reset_kernel(long long* hstats, const double* sp) {
    if (threadIdx.x == 0 && sp[NB200_SP_UNSAFE] != 0.0) {
        hstats[NB200_HS_MAX_ABS_BITS] = 0;
        hstats[NB200_HS_MAX_FROBSQ_BITS] = 0;
    }
}

__global__ void __launch_bounds__(256)
shell_dump_kernel(const float* __restrict__ g, Shell sh, nb::Spacing3 sp, DumpParams p) {
    const long long total = sh.n1 + sh.n2 + sh.n3;
    const long long vol = (long long)sh.v.nz_buf * sh.v.ny * sh.v.nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int zb, y, x;
        shell_voxel(sh, i, zb, y, x);
        long long idx;
        const ShellHessian h = shell_hessian(g, sh, sp, zb, y, x, idx);
        p.out[idx] = h.zz; p.out[vol + idx] = h.zy; p.out[2 * vol + idx] = h.zx;
        p.out[3 * vol + idx] = h.yy; p.out[4 * vol + idx] = h.yx; p.out[5 * vol + idx] = h.xx;
    }
}

__global__ void __launch_bounds__(256)
shell_frangi_kernel(const float* __restrict__ g, Shell sh, nb::Spacing3 sp, FrangiParams p) {
    if (p.spd[NB200_SP_SKIP] != 0.0) return;
    const float gamma_sq = (float)p.spd[NB200_SP_GAMMA_SQ];
    const float fs_min = (float)p.spd[NB200_SP_FROBSQ_MIN];
    const ZeroTests zt = zero_tests_from(p.spd);
    const long long total = sh.n1 + sh.n2 + sh.n3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int zb, y, x;
        shell_voxel(sh, i, zb, y, x);
        const long long plane = (long long)sh.v.ny * sh.v.nx;
        const long long at = (long long)zb * plane + (long long)y * sh.v.nx + x;
        const float prev = p.acc[at];
        if (prev < 0.0f) continue;
        long long idx;
        const ShellHessian h = shell_hessian(g, sh, sp, zb, y, x, idx);
        const float fs = nb::frob_sq3(h.zz, h.zy, h.zx, h.yy, h.yx, h.xx);
        if (!(fs >= fs_min)) { p.acc[at] = -1.0f; continue; }
        if (nb::pd_reject_diag(h.zz, h.yy, h.xx, zt.tau1)) continue;
        if (nb::pd_reject_full(h.zz, h.zy, h.zx, h.yy, h.yx, h.xx, zt.tau2, zt.tau3)) continue;
        const float vv = eig_vesselness(h.zz, h.zy, h.zx, h.yy, h.yx, h.xx, p.alpha_sq, p.beta_sq, gamma_sq);
        if (vv > prev) p.acc[at] = vv;
    }
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

}  // namespace
namespace nb {
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
}  // namespace nb
namespace {
using nb::check_vol;

hm::Divs divs_from(const float* s) {
    hm::Divs dv;
    for (int a = 0; a < 3; ++a) {
        dv.d2[a] = s[2 * a + 1];
        dv.r2[a] = 1.0f / s[2 * a + 1];
    }
    return dv;
}
nb::Spacing3 spacing3_from(const float* s) {
    nb::Spacing3 sp;
    for (int a = 0; a < 3; ++a) {
        sp.h1[a] = s[2 * a];
        sp.h2[a] = s[2 * a + 1];
        sp.r1[a] = 1.0f / s[2 * a];
        sp.r2[a] = 1.0f / s[2 * a + 1];
    }
    return sp;
}

// ---- TMA descriptor of a (nz_buf, ny, nx) float32 volume, box = one staged plane tile ------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace
namespace nb {
// returns true when `map` describes g for TMA; false -> the kernel uses plain loads
bool make_plane_map(const float* g, const nb200_vol& v, CUtensorMap* map) {
    memset(map, 0, sizeof(*map));
    static const bool disabled = getenv("NB200_NO_TMA") != nullptr;
    if (disabled || v.nx % 4 != 0 || (reinterpret_cast<unsigned long long>(g) & 15ull) != 0) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)v.nx, (cuuint64_t)v.ny, (cuuint64_t)v.nz_buf};
    const cuuint64_t strides[2] = {(cuuint64_t)v.nx * 4ull, (cuuint64_t)v.nx * (cuuint64_t)v.ny * 4ull};
    const cuuint32_t box[3] = {(cuuint32_t)hm::PITCH, (cuuint32_t)hm::GR, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
}  // namespace nb
namespace {
using nb::make_plane_map;
using nb::MarchPlan;
using nb::plan_march;

template <class Epi>
constexpr size_t smem_bytes_for() { return sizeof(hm::Smem); }
template <>
constexpr size_t smem_bytes_for<FrangiEpi>() { return sizeof(hm::Smem) + sizeof(FrangiQueue); }

}  // namespace
namespace nb {
// Interior of the compute window; Z chunks sized for several waves of 2 CTAs/SM, no shorter than 32 planes
MarchPlan plan_march(const nb200_vol& v) {
    MarchPlan m;
    const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off;
    m.zi0 = max(g0, 2);
    m.zi1 = min(g1, v.nz_glob - 2);
    m.zchunk = 0;
    m.n_ctas = 0;
    if (m.zi1 <= m.zi0 || v.ny < 5 || v.nx < 12) return m;
    const long long tiles = (long long)((v.nx + hm::TX - 1) / hm::TX) * ((v.ny - 4 + hm::TYO - 1) / hm::TYO);
    const int nz = m.zi1 - m.zi0;
    const long long want = 16LL * nb::sm_count();
    long long chunks = (want + tiles - 1) / tiles;
    if (chunks < 1) chunks = 1;
    int zchunk = (int)((nz + chunks - 1) / chunks);
    if (zchunk < 32) zchunk = nz < 32 ? nz : 32;
    chunks = (nz + zchunk - 1) / zchunk;
    m.zchunk = zchunk;
    m.n_ctas = tiles * chunks;
    return m;
}
}  // namespace nb
namespace {

template <int MODE, class Epi, class Params>
int launch_one(const float* g, const nb200_vol& v, const CUtensorMap& map, bool use_tma, const MarchPlan& m,
               const float* spacing, const double* sp, int run_if_unsafe, int check_skip, const Params& p,
               cudaStream_t st) {
    auto kernel = march_kernel<MODE, Epi, Params>;
    constexpr size_t smem = smem_bytes_for<Epi>();
    static bool smem_set = false;       // one flag per <MODE, Epi> instantiation
    if (!smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            nb::set_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
            return NB200_ERR_CUDA;
        }
        smem_set = true;
    }
    kernel<<<(unsigned)m.n_ctas, hm::NT, smem, st>>>(map, g, v, divs_from(spacing), sp, run_if_unsafe, m.zi0, m.zi1,
                                                     m.zchunk, check_skip, use_tma ? 1 : 0, p);
    return NB200_OK;
}

// interior march: FAST mode enqueues the fast kernel and its IEEE twin; the device flag sp[UNSAFE] picks the
// one that runs
// twin: 0 = fast launch (runs unless sp[UNSAFE]) + IEEE twin (runs if sp[UNSAFE]);  1 = fast launch only,
// unconditional (K2's first pass, which also measures the value range the flag is derived from);
// 2 = exact redo pass, runs if sp[UNSAFE];  3 = exact pass that runs if sp[AMBIG] (and not UNSAFE)
template <class Epi, class Params>
int launch_march(const float* g, const nb200_vol& v, const float* spacing, int div_mode, const double* sp,
                 int check_skip, const Params& p, cudaStream_t st, const char* what, int twin = 0) {
    const MarchPlan m = plan_march(v);
    if (m.n_ctas == 0) return NB200_OK;
    CUtensorMap map;
    const bool use_tma = make_plane_map(g, v, &map);
    int rc;
    if (twin == 2 || twin == 3) {
        // gated passes: 2 = exact redo when sp[UNSAFE] (IEEE division replaces the fast sequence), 3 = exact pass when
        // the emptiness of the mask is undecided (sp[AMBIG], UNSAFE clear: the native mode is valid)
        const int gate = twin == 2 ? 1 : 2;
        if (sp == nullptr) return NB200_OK;
        if (div_mode == hm::DIV_POW2) rc = launch_one<hm::DIV_POW2, Epi>(g, v, map, use_tma, m, spacing, sp, gate, check_skip, p, st);
        else if (div_mode == hm::DIV_FAST && twin == 3) rc = launch_one<hm::DIV_FAST, Epi>(g, v, map, use_tma, m, spacing, sp, gate, check_skip, p, st);
        else rc = launch_one<hm::DIV_IEEE, Epi>(g, v, map, use_tma, m, spacing, sp, gate, check_skip, p, st);
    }
    else if (div_mode == hm::DIV_POW2) rc = launch_one<hm::DIV_POW2, Epi>(g, v, map, use_tma, m, spacing, sp, -1, check_skip, p, st);
    else if (div_mode == hm::DIV_IEEE || sp == nullptr) rc = launch_one<hm::DIV_IEEE, Epi>(g, v, map, use_tma, m, spacing, sp, -1, check_skip, p, st);
    else if (twin == 1) rc = launch_one<hm::DIV_FAST, Epi>(g, v, map, use_tma, m, spacing, sp, -1, check_skip, p, st);
    else {
        rc = launch_one<hm::DIV_FAST, Epi>(g, v, map, use_tma, m, spacing, sp, 0, check_skip, p, st);
        if (rc == NB200_OK) rc = launch_one<hm::DIV_IEEE, Epi>(g, v, map, use_tma, m, spacing, sp, 1, check_skip, p, st);
    }
    if (rc) return rc;
    return nb::check_launch(what);
}

unsigned shell_grid(const Shell& sh) {
    return nb::grid_for(sh.n1 + sh.n2 + sh.n3, 256, 8);
}

}  // namespace

namespace nb {
int launch_shell_stats(const float* g, const nb200_vol& v, const float* spacing, int sz, int sy, int sx,
                       float* frob_samples, long long* hstats, float* code, const double* sp, int gate,
                       cudaStream_t st) {
    StatsParams p;
    p.sz = sz; p.sy = sy; p.sx = sx;
    const int g0 = v.zc0 + v.zg_off;
    p.g_first = ((g0 + sz - 1) / sz) * sz;
    p.ly_n = (v.ny + sy - 1) / sy;
    p.lx_n = (v.nx + sx - 1) / sx;
    p.frob_samples = frob_samples;
    p.hstats = hstats;
    p.code = code;
    p.code_vec_ok = 0;
    const Shell sh = make_shell(v);
    if (sh.n1 + sh.n2 + sh.n3 <= 0) return NB200_OK;
    shell_stats_kernel<<<shell_grid(sh), 256, 0, st>>>(g, sh, spacing3_from(spacing), p, sp, sp ? gate : -1);
    return nb::check_launch("hessian_stats(shell)");
}

int launch_shell_frangi(const float* g, float* acc, const nb200_vol& v, const float* spacing, float alpha_sq,
                        float beta_sq, const double* sp, cudaStream_t st) {
    FrangiParams p;
    p.acc = acc;
    p.alpha_sq = alpha_sq;
    p.beta_sq = beta_sq;
    p.spd = sp;
    p.acc_vec_ok = 0;
    p.debug = 0;
    const Shell sh = make_shell(v);
    if (sh.n1 + sh.n2 + sh.n3 <= 0) return NB200_OK;
    shell_frangi_kernel<<<shell_grid(sh), 256, 0, st>>>(g, sh, spacing3_from(spacing), p);
    return nb::check_launch("frangi_accumulate(shell)");
}
}  // namespace nb

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
    // outside [2^-12, 2^12] the value-range argument behind sp[UNSAFE] (finalize_max_abs_kernel) does not hold
    if (!(d >= 0.000244140625f && d <= 4096.0f)) { *mode_out = hm::DIV_IEEE; return NB200_OK; }
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
    return nb200_hessian_stats_code(gauss, vol, spacing, div_mode, sp, sz, sy, sx, frob_samples, hstats, nullptr, stream);
}

namespace {
// pass 0 = first (fast division unconditional when div_mode is FAST), pass 1 = exact redo if sp[UNSAFE] (IEEE division
// when the mode was FAST), pass 2 = exact statistics if sp[AMBIG] (max frob_sq decides whether the mask is empty)
int hessian_stats_impl(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode, const double* sp,
                       int sz, int sy, int sx, float* frob_samples, long long* hstats, float* code, void* stream,
                       int pass) {
    NB_REQUIRE(gauss && vol && spacing && hstats && sz > 0 && sy > 0 && sx > 0, NB200_ERR_ARG,
               "nb200_hessian_stats: bad argument");
    NB_REQUIRE(code != gauss, NB200_ERR_ARG, "nb200_hessian_stats: code must not alias the blurred volume");
    NB_REQUIRE(div_mode >= 0 && div_mode <= 2, NB200_ERR_ARG, "nb200_hessian_stats: div_mode %d", div_mode);
    const nb200_vol v = *vol;
    int rc = check_vol(v, "nb200_hessian_stats");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    if (pass != 0 && sp == nullptr) return NB200_OK;   // nothing to gate on
    StatsParams p;
    p.sz = sz; p.sy = sy; p.sx = sx;
    const int g0 = v.zc0 + v.zg_off;
    p.g_first = ((g0 + sz - 1) / sz) * sz;
    p.ly_n = (v.ny + sy - 1) / sy;
    p.lx_n = (v.nx + sx - 1) / sx;
    p.frob_samples = frob_samples;
    p.hstats = hstats;
    p.code = code;
    p.code_vec_ok = (v.nx % 4 == 0) && ((reinterpret_cast<unsigned long long>(code) & 15ull) == 0);
    cudaStream_t st = nb::as_stream(stream);
    if (pass == 1) {
        hstats_redo_reset_kernel<<<1, 32, 0, st>>>(hstats, sp);
        rc = nb::check_launch("hessian_stats(redo reset)");
        if (rc) return rc;
    }
    rc = launch_march<StatsEpi>(gauss, v, spacing, div_mode, sp, 0, p, st, "hessian_stats", pass + 1);
    if (rc) return rc;
    return nb::launch_shell_stats(gauss, v, spacing, sz, sy, sx, frob_samples, hstats, code, sp,
                                  pass == 0 ? -1 : pass, st);
}
}  // namespace


int nb200_hessian_stats_code(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                             float* code, void* stream) {
    return hessian_stats_impl(gauss, vol, spacing, div_mode, sp, sz, sy, sx, frob_samples, hstats, code, stream, 0);
}

int nb200_hessian_stats_redo(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                             float* code, void* stream) {
    return hessian_stats_impl(gauss, vol, spacing, div_mode, sp, sz, sy, sx, frob_samples, hstats, code, stream, 1);
}

int nb200_hessian_stats_ambig(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                              const double* sp, int sz, int sy, int sx, float* frob_samples, long long* hstats,
                              void* stream) {
    return hessian_stats_impl(gauss, vol, spacing, div_mode, sp, sz, sy, sx, frob_samples, hstats, nullptr, stream, 2);
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
    const MarchPlan m = plan_march(v);
    if (m.n_ctas > 0) {
        CUtensorMap map;
        const bool use_tma = make_plane_map(gauss, v, &map);
        if (div_mode == hm::DIV_POW2) rc = launch_one<hm::DIV_POW2, DumpEpi>(gauss, v, map, use_tma, m, spacing, nullptr, -1, 0, p, st);
        else if (div_mode == hm::DIV_FAST) rc = launch_one<hm::DIV_FAST, DumpEpi>(gauss, v, map, use_tma, m, spacing, nullptr, -1, 0, p, st);
        else rc = launch_one<hm::DIV_IEEE, DumpEpi>(gauss, v, map, use_tma, m, spacing, nullptr, -1, 0, p, st);
        if (rc) return rc;
        rc = nb::check_launch("hessian_components");
        if (rc) return rc;
    }
    const Shell sh = make_shell(v);
    if (sh.n1 + sh.n2 + sh.n3 > 0) {
        shell_dump_kernel<<<shell_grid(sh), 256, 0, st>>>(gauss, sh, spacing3_from(spacing), p);
        rc = nb::check_launch("hessian_components(shell)");
    }
    return rc;
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
    cudaStream_t st = nb::as_stream(stream);
    rc = launch_march<FrangiEpi>(gauss, v, spacing, div_mode, sp, 1, p, st, "frangi_accumulate");
    if (rc) return rc;
    // the border shell runs after the interior on the same stream (the interior rewrites unchanged shell
    // values inside partially valid groups of four)
    const Shell sh = make_shell(v);
    if (sh.n1 + sh.n2 + sh.n3 > 0) {
        shell_frangi_kernel<<<shell_grid(sh), 256, 0, st>>>(gauss, sh, spacing3_from(spacing), p);
        rc = nb::check_launch("frangi_accumulate(shell)");
    }
    return rc;
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
