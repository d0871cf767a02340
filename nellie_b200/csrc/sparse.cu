// K3, sparse form — Frobenius mask + 3x3 eigenvalues + vesselness + max/AND accumulate
// (nellie/segmentation/filtering.py:407-444, :651-767, :842-851) from the per-voxel record that the Hessian
// statistics kernel (K2, frangi.cu: voxel_code) leaves behind:
//     |code| = frob_sq, bit-identical to the value the mask tests;  sign bit = "response provably zero".
//
// The dense march of frangi.cu evaluates every Hessian a second time although only voxels that are alive,
// pass this sigma's mask and are not provably zero (a few percent of a frame) need eigenvalues.  Here the
// dense part is a pure stream — read code (4 B) and acc (4 B), write acc where a voxel dies — and the
// survivors go through two shared-memory work queues that persist across the bricks a CTA visits:
//   RAW (packed z,y,x)         filled by the stream; drained 256 at a time: every thread evaluates ONE
//                              Hessian with the generic per-voxel rules of hessian.cuh (one-sided differences
//                              at the frame border, IEEE division: bit-identical to the march) from global
//                              memory, which the brick order keeps in L1/L2, then the full "provably zero"
//                              test (pd_reject_full) with margins relative to the voxel's own norm;
//   RDY (six entries + index)  survivors; drained 256 at a time by the eigen-solver + vesselness.
// Both expensive phases run with full warps whatever the shape of the mask.
// Algorithmic HBM traffic: 12 B/voxel (code R, acc R, acc W) + the blurred neighbourhoods of the candidates.
#include <stdlib.h>

#include "common.cuh"
#include "hessian.cuh"

namespace {

constexpr int NT = 256;
constexpr int BX = 128, BY = 8, BZ = 16;            // brick; one plane of it = 1024 voxels = 4 per thread
constexpr int PPG = 2;                              // planes streamed per barrier
constexpr int RAW_CAP = NT + PPG * BX * BY;         // < 256 left over + one group pushed
constexpr int RDY_CAP = 2 * NT;                     // < 256 left over + at most 256 pushed per RAW round

struct Queues {
    unsigned long long raw[RAW_CAP];
    float rdy[8][RDY_CAP];                          // zz, zy, zx, yy, yx, xx, index lo, index hi
    int n_raw, n_rdy;
};

struct SparseParams {
    const float* g;
    const float* code;
    float* acc;
    nb200_vol v;
    nb::Spacing3 sp;
    float alpha_sq, beta_sq;
    const double* spd;
    int vec_ok;
    int div_mode;
    int nbx, nby, nbz;
    int bz;                 // brick depth in planes
    int only_if_unsafe;     // fallback of nb200_frangi_fast: run only when sp[UNSAFE] is set
};

struct GlobalLoad3 {
    const float* p; long long plane; int nx;
    __device__ __forceinline__ float operator()(int dz, int dy, int dx) const {
        return __ldg(p + dz * plane + (long long)dy * nx + dx);
    }
};

__device__ __noinline__ float eig_vesselness(float a00, float a01, float a02, float a11, float a12, float a22,
                                             float alpha_sq, float beta_sq, float gamma_sq) {
    float l1, l2, l3;
    nb::eig3_sym<2>(a00, a01, a02, a11, a12, a22, l1, l2, l3);
    if (l3 > 0.0f || l2 > 0.0f) return 0.0f;           // filtering.py:759-761 zeroes these responses
    return nb::vesselness3(l1, l2, l3, alpha_sq, beta_sq, gamma_sq);
}

// (a - b) / fl32(2h) of one axis
template <int MODE>
__device__ __forceinline__ float cdiff(float a, float b, float d, float r) {
    const float n = a - b;
    if (MODE == 2) return n * r;
    if (MODE == 1) {
        const float q0 = n * r;
        const float e = fmaf(-q0, d, n);
        return fmaf(e, r, q0);
    }
    return n / d;
}

template <int MODE>
__device__ __forceinline__ void candidate_hessian(const SparseParams& p, long long at, long long plane, int zb, int y, int x,
                                                  float (&h)[6]) {
    const int zg = zb + p.v.zg_off;
    const int nx = p.v.nx;
    if (zg >= 2 && zg < p.v.nz_glob - 2 && y >= 2 && y < p.v.ny - 2 && x >= 2 && x < nx - 2) {
        // interior: central differences only (19 distinct samples, fixed offsets); same operations in the same
        // order as nb::hessian3 / numpy.gradient(numpy.gradient(g))
        const float* c = p.g + at;
        const long long P = plane;
        const float dz = p.sp.h2[0], dy = p.sp.h2[1], dx = p.sp.h2[2];
        const float rz = p.sp.r2[0], ry = p.sp.r2[1], rx = p.sp.r2[2];
        const float g0 = __ldg(c);
        const float zp2 = __ldg(c + 2 * P), zm2 = __ldg(c - 2 * P);
        const float yp2 = __ldg(c + 2 * nx), ym2 = __ldg(c - 2 * nx);
        const float xp2 = __ldg(c + 2), xm2 = __ldg(c - 2);
        const float zpyp = __ldg(c + P + nx), zpym = __ldg(c + P - nx), zmyp = __ldg(c - P + nx), zmym = __ldg(c - P - nx);
        const float zpxp = __ldg(c + P + 1), zpxm = __ldg(c + P - 1), zmxp = __ldg(c - P + 1), zmxm = __ldg(c - P - 1);
        const float ypxp = __ldg(c + nx + 1), ypxm = __ldg(c + nx - 1), ymxp = __ldg(c - nx + 1), ymxm = __ldg(c - nx - 1);
        // d0 d0: gz(z+1) - gz(z-1)
        h[0] = cdiff<MODE>(cdiff<MODE>(zp2, g0, dz, rz), cdiff<MODE>(g0, zm2, dz, rz), dz, rz);
        // d1 d0: gz(y+1) - gz(y-1)
        h[1] = cdiff<MODE>(cdiff<MODE>(zpyp, zmyp, dz, rz), cdiff<MODE>(zpym, zmym, dz, rz), dy, ry);
        // d2 d0: gz(x+1) - gz(x-1)
        h[2] = cdiff<MODE>(cdiff<MODE>(zpxp, zmxp, dz, rz), cdiff<MODE>(zpxm, zmxm, dz, rz), dx, rx);
        // d1 d1
        h[3] = cdiff<MODE>(cdiff<MODE>(yp2, g0, dy, ry), cdiff<MODE>(g0, ym2, dy, ry), dy, ry);
        // d2 d1: gy(x+1) - gy(x-1)
        h[4] = cdiff<MODE>(cdiff<MODE>(ypxp, ymxp, dy, ry), cdiff<MODE>(ypxm, ymxm, dy, ry), dx, rx);
        // d2 d2
        h[5] = cdiff<MODE>(cdiff<MODE>(xp2, g0, dx, rx), cdiff<MODE>(g0, xm2, dx, rx), dx, rx);
        return;
    }
    GlobalLoad3 L{p.g + at, plane, nx};
    const int n3[3] = {p.v.nz_glob, p.v.ny, nx};
    nb::hessian3<GlobalLoad3, MODE>(L, zg, y, x, n3, p.sp, h[0], h[1], h[2], h[3], h[4], h[5]);
}

__global__ void __launch_bounds__(NT, 3)
frangi_sparse_kernel(const SparseParams p) {
    if (p.spd[NB200_SP_SKIP] != 0.0) return;            // empty mask: the sigma contributes nothing (:843-844)
    if (p.only_if_unsafe && p.spd[NB200_SP_UNSAFE] == 0.0) return;
    __shared__ Queues q;
    const float gamma_sq = (float)p.spd[NB200_SP_GAMMA_SQ];
    const float fs_min = (float)p.spd[NB200_SP_FROBSQ_MIN];   // mask <=> frob_sq >= fs_min (finalize_frob_kernel)
    const int div_mode = (p.div_mode == NB200_DIV_FAST && p.spd[NB200_SP_UNSAFE] != 0.0) ? NB200_DIV_IEEE : p.div_mode;
    const nb200_vol v = p.v;
    const long long plane = (long long)v.ny * v.nx;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (threadIdx.x == 0) { q.n_raw = 0; q.n_rdy = 0; }
    __syncthreads();

    // RAW[n-m, n): Hessian + full zero test; survivors -> RDY.  Called by all threads (CTA-uniform n, m).
    auto round_a = [&](int n, int m) {
        const int e = n - m + (int)threadIdx.x;
        bool keep = false;
        float h[6];
        long long at = 0;
        if (e < n) {
            const unsigned long long w = q.raw[e];
            const int zb = (int)(w >> 44), y = (int)((w >> 22) & 0x3fffffu), x = (int)(w & 0x3fffffu);
            at = (long long)zb * plane + (long long)y * v.nx + x;
            // division: the verified constant-divisor sequence unless K2 flagged the value range as unsafe
            if (div_mode == NB200_DIV_POW2) candidate_hessian<2>(p, at, plane, zb, y, x, h);
            else if (div_mode == NB200_DIV_FAST) candidate_hessian<1>(p, at, plane, zb, y, x, h);
            else candidate_hessian<0>(p, at, plane, zb, y, x, h);
            // margins relative to this voxel's own Frobenius norm (F^2 = frob_sq * 1.001 >= ||H||_F^2); outside the
            // range pd_margins accepts nothing is rejected
            const float fs = nb::frob_sq3(h[0], h[1], h[2], h[3], h[4], h[5]);
            float tau2 = INFINITY, tau3 = INFINITY;
            if (fs > 1e-20f && fs < 1e20f) {
                const float f2 = 1.001f * fs;
                tau2 = 1e-5f * f2;
                tau3 = 1e-4f * (f2 * sqrtf(f2));
            }
            keep = !nb::pd_reject_full(h[0], h[1], h[2], h[3], h[4], h[5], tau2, tau3);
        }
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if (bits != 0u) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&q.n_rdy, __popc(bits));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                const int o = base + __popc(bits & lt);
#pragma unroll
                for (int j = 0; j < 6; ++j) q.rdy[j][o] = h[j];
                q.rdy[6][o] = __uint_as_float((unsigned)(at & 0xffffffffll));
                q.rdy[7][o] = __uint_as_float((unsigned)(at >> 32));
            }
        }
    };
    // RDY[n-m, n): eigenvalues + vesselness + running max
    auto round_b = [&](int n, int m) {
        const int e = n - m + (int)threadIdx.x;
        if (e < n) {
            const long long at = (long long)__float_as_uint(q.rdy[6][e]) | ((long long)__float_as_uint(q.rdy[7][e]) << 32);
            const float cur = p.acc[at];                 // issued before the solve: latency hidden
            const float vv = eig_vesselness(q.rdy[0][e], q.rdy[1][e], q.rdy[2][e], q.rdy[3][e], q.rdy[4][e], q.rdy[5][e],
                                            p.alpha_sq, p.beta_sq, gamma_sq);
            if (vv > cur) p.acc[at] = vv;                // acc >= 0 here: a zero response changes nothing
        }
    };
    // drain the queues down to < 256 entries each (all = to empty); called by every thread of the CTA
    auto drain = [&](bool all) {
        while (true) {
            __syncthreads();                             // pushes / counter updates of the previous step are visible
            const int n = *reinterpret_cast<volatile int*>(&q.n_raw);
            const int nr = *reinterpret_cast<volatile int*>(&q.n_rdy);
            __syncthreads();                             // everyone has read the same counters before anyone changes them
            const bool do_a = n >= NT || (all && n > 0);
            const bool do_b = !do_a && (nr >= NT || (all && nr > 0));
            if (!do_a && !do_b) break;
            if (do_a) {
                const int m = n < NT ? n : NT;
                if (threadIdx.x == 0) q.n_raw = n - m;
                round_a(n, m);
            } else {
                const int m = nr < NT ? nr : NT;
                if (threadIdx.x == 0) q.n_rdy = nr - m;
                round_b(nr, m);
            }
            // after a RAW round RDY may hold up to 511 entries; the next iteration sees n_rdy >= 256 only if RAW is
            // below 256, so solve right away
            if (do_a) {
                __syncthreads();
                const int nr2 = *reinterpret_cast<volatile int*>(&q.n_rdy);
                __syncthreads();
                if (nr2 >= NT) {
                    if (threadIdx.x == 0) q.n_rdy = nr2 - NT;
                    round_b(nr2, NT);
                }
            }
        }
    };

    const int tx = 4 * lane, ty = threadIdx.x >> 5;
    const long long nbricks = (long long)p.nbx * p.nby * p.nbz;
    for (long long b = blockIdx.x; b < nbricks; b += gridDim.x) {
        long long r = b;
        const int bx = (int)(r % p.nbx); r /= p.nbx;
        const int by = (int)(r % p.nby); r /= p.nby;
        const int x = bx * BX + tx, y = by * BY + ty;
        const int z0 = v.zc0 + (int)r * BZ;
        const int z1 = min(z0 + BZ, v.zc1);
        const bool row_in = y < v.ny && x < v.nx;
        for (int zg0 = z0; zg0 < z1; zg0 += PPG) {
            float4 c[PPG], a[PPG];
            bool in[PPG];
#pragma unroll
            for (int i = 0; i < PPG; ++i) {
                const int zb = zg0 + i;
                in[i] = row_in && zb < z1;
                c[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                a[i] = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
                if (in[i]) {
                    const long long at = (long long)zb * plane + (long long)y * v.nx + x;
                    if (p.vec_ok) {
                        c[i] = __ldg(reinterpret_cast<const float4*>(p.code + at));
                        a[i] = *reinterpret_cast<const float4*>(p.acc + at);
                    } else {
                        float* cc = &c[i].x; float* aa = &a[i].x;
                        for (int k = 0; k < 4; ++k)
                            if (x + k < v.nx) { cc[k] = __ldg(p.code + at + k); aa[k] = p.acc[at + k]; }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PPG; ++i) {
                const int zb = zg0 + i;
                const float* cc = &c[i].x;
                float* aa = &a[i].x;
                unsigned cand = 0, kill = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool alive = aa[k] >= 0.0f;
                    const bool pass = fabsf(cc[k]) >= fs_min;            // NaN fails, like the reference's comparison
                    if (alive && !pass) { kill |= 1u << k; aa[k] = -1.0f; }   // dead voxels stay dead (AND of masks)
                    if (alive && pass && !(__float_as_uint(cc[k]) >> 31)) cand |= 1u << k;
                }
                if (kill) {
                    const long long at = (long long)zb * plane + (long long)y * v.nx + x;
                    if (p.vec_ok) *reinterpret_cast<float4*>(p.acc + at) = a[i];
                    else for (int k = 0; k < 4; ++k) if ((kill >> k) & 1u) p.acc[at + k] = -1.0f;
                }
                // warp-aggregated push of the candidates
                if (__any_sync(0xffffffffu, cand != 0u)) {
                    const unsigned b0 = __ballot_sync(0xffffffffu, cand & 1u), b1 = __ballot_sync(0xffffffffu, cand & 2u);
                    const unsigned b2 = __ballot_sync(0xffffffffu, cand & 4u), b3 = __ballot_sync(0xffffffffu, cand & 8u);
                    const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&q.n_raw, n0 + n1 + n2 + n3);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (cand & 1u) q.raw[base + __popc(b0 & lt)] = ((unsigned long long)zb << 44) | ((unsigned long long)y << 22) | (unsigned)(x);
                    if (cand & 2u) q.raw[base + n0 + __popc(b1 & lt)] = ((unsigned long long)zb << 44) | ((unsigned long long)y << 22) | (unsigned)(x + 1);
                    if (cand & 4u) q.raw[base + n0 + n1 + __popc(b2 & lt)] = ((unsigned long long)zb << 44) | ((unsigned long long)y << 22) | (unsigned)(x + 2);
                    if (cand & 8u) q.raw[base + n0 + n1 + n2 + __popc(b3 & lt)] = ((unsigned long long)zb << 44) | ((unsigned long long)y << 22) | (unsigned)(x + 3);
                }
            }
            drain(false);
        }
    }
    drain(true);
}

// --------------------------------------------------------------------------------------------
// Two-kernel form (used when the caller lends a candidate list: one uint32 per voxel of capacity).
//   stream:  code + acc -> deaths written back, candidates appended to a global list.  No barrier anywhere:
//            every warp stages its candidates in its own shared-memory slice and flushes 128 at a time with
//            one global atomicAdd, so the kernel is a pure HBM stream (12 B/voxel) at high occupancy.
//   solve:   warps walk the list 32 candidates at a time: Hessian + full zero test, survivors compacted into a
//            per-warp queue, eigenvalues + vesselness for full warps of survivors.  No CTA barrier either.
// The list order depends on scheduling; results do not (every voxel is independent and owns its acc word).
// --------------------------------------------------------------------------------------------
constexpr int WSTAGE = 256;                 // per-warp staging slots (flush at >= 128, at most 128 pushed per step)

__global__ void __launch_bounds__(NT)
sparse_stream_kernel(const SparseParams p, unsigned* __restrict__ list, unsigned long long* __restrict__ counter) {
    if (p.spd[NB200_SP_SKIP] != 0.0) return;
    if (p.only_if_unsafe && p.spd[NB200_SP_UNSAFE] == 0.0) return;
    __shared__ unsigned stage[NT / 32][WSTAGE];
    const float fs_min = (float)p.spd[NB200_SP_FROBSQ_MIN];
    const nb200_vol v = p.v;
    const long long plane = (long long)v.ny * v.nx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned* mine = stage[warp];
    int n_st = 0;                                        // warp-uniform
    auto flush = [&](int keep_below) {
        // write stage[0, n_st) out when it holds at least keep_below entries
        if (n_st >= keep_below && n_st > 0) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)n_st);
            base = __shfl_sync(0xffffffffu, base, 0);
            __syncwarp();
            for (int i = lane; i < n_st; i += 32) list[base + i] = mine[i];
            __syncwarp();
            n_st = 0;
        }
    };
    const int tx = 4 * lane;
    const long long nbricks = (long long)p.nbx * p.nby * p.nbz;
    for (long long b = blockIdx.x; b < nbricks; b += gridDim.x) {
        long long r = b;
        const int bx = (int)(r % p.nbx); r /= p.nbx;
        const int by = (int)(r % p.nby); r /= p.nby;
        const int x = bx * BX + tx, y = by * BY + warp;   // one row of the brick per warp
        const int z0 = v.zc0 + (int)r * p.bz;
        const int z1 = min(z0 + p.bz, v.zc1);
        const bool row_in = y < v.ny && x < v.nx;
        constexpr int G = 4;                             // planes in flight per thread
        for (int zg0 = z0; zg0 < z1; zg0 += G) {
            float4 c[G], a[G];
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int zb = zg0 + i;
                c[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                a[i] = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
                if (row_in && zb < z1) {
                    const long long at = (long long)zb * plane + (long long)y * v.nx + x;
                    if (p.vec_ok) {
                        c[i] = __ldg(reinterpret_cast<const float4*>(p.code + at));
                        a[i] = *reinterpret_cast<const float4*>(p.acc + at);
                    } else {
                        float* cc = &c[i].x; float* aa = &a[i].x;
                        for (int k = 0; k < 4; ++k)
                            if (x + k < v.nx) { cc[k] = __ldg(p.code + at + k); aa[k] = p.acc[at + k]; }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int zb = zg0 + i;
                const float* cc = &c[i].x;
                float* aa = &a[i].x;
                unsigned cand = 0, kill = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool alive = aa[k] >= 0.0f;
                    const bool pass = fabsf(cc[k]) >= fs_min;            // NaN fails, like the reference's comparison
                    if (alive && !pass) { kill |= 1u << k; aa[k] = -1.0f; }   // dead voxels stay dead (AND of masks)
                    if (alive && pass && !(__float_as_uint(cc[k]) >> 31)) cand |= 1u << k;
                }
                const long long at = (long long)zb * plane + (long long)y * v.nx + x;
                if (kill) {
                    if (p.vec_ok) *reinterpret_cast<float4*>(p.acc + at) = a[i];
                    else for (int k = 0; k < 4; ++k) if ((kill >> k) & 1u) p.acc[at + k] = -1.0f;
                }
                if (__any_sync(0xffffffffu, cand != 0u)) {
                    const unsigned b0 = __ballot_sync(0xffffffffu, cand & 1u), b1 = __ballot_sync(0xffffffffu, cand & 2u);
                    const unsigned b2 = __ballot_sync(0xffffffffu, cand & 4u), b3 = __ballot_sync(0xffffffffu, cand & 8u);
                    const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
                    const unsigned at32 = (unsigned)at;
                    if (cand & 1u) mine[n_st + __popc(b0 & lt)] = at32;
                    if (cand & 2u) mine[n_st + n0 + __popc(b1 & lt)] = at32 + 1u;
                    if (cand & 4u) mine[n_st + n0 + n1 + __popc(b2 & lt)] = at32 + 2u;
                    if (cand & 8u) mine[n_st + n0 + n1 + n2 + __popc(b3 & lt)] = at32 + 3u;
                    n_st += n0 + n1 + n2 + n3;
                    flush(WSTAGE - 128);
                }
            }
        }
    }
    flush(1);
}

constexpr int WQ = 64;                      // per-warp survivor queue (< 32 left over + <= 32 pushed)

__global__ void __launch_bounds__(NT, 4)
sparse_solve_kernel(const SparseParams p, const unsigned* __restrict__ list, const unsigned long long* __restrict__ counter) {
    if (p.spd[NB200_SP_SKIP] != 0.0) return;
    if (p.only_if_unsafe && p.spd[NB200_SP_UNSAFE] == 0.0) return;
    __shared__ float wq[NT / 32][7][WQ];
    const float gamma_sq = (float)p.spd[NB200_SP_GAMMA_SQ];
    const int div_mode = (p.div_mode == NB200_DIV_FAST && p.spd[NB200_SP_UNSAFE] != 0.0) ? NB200_DIV_IEEE : p.div_mode;
    const nb200_vol v = p.v;
    const unsigned plane = (unsigned)((long long)v.ny * v.nx), nx = (unsigned)v.nx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    float (*q)[WQ] = wq[warp];
    int n_q = 0;                                         // warp-uniform
    const unsigned long long n = *counter;
    const unsigned long long wstride = (unsigned long long)gridDim.x * (NT / 32) * 32ull;
    auto solve32 = [&](int cnt) {                        // the top `cnt` (<= 32) entries of this warp's queue
        __syncwarp();
        const int e = n_q - cnt + lane;
        if (lane < cnt) {
            const unsigned at = __float_as_uint(q[6][e]);
            const float cur = p.acc[at];
            const float vv = eig_vesselness(q[0][e], q[1][e], q[2][e], q[3][e], q[4][e], q[5][e], p.alpha_sq, p.beta_sq, gamma_sq);
            if (vv > cur) p.acc[at] = vv;                // acc >= 0 here: a zero response changes nothing
        }
        n_q -= cnt;
        __syncwarp();
    };
    const unsigned long long first = ((unsigned long long)blockIdx.x * (NT / 32) + warp) * 32ull + lane;
    unsigned at_next = first < n ? __ldg(list + first) : 0u;         // list entries are requested one chunk ahead
    for (unsigned long long base = first - lane; base < n; base += wstride) {
        const unsigned long long i = base + lane;
        bool keep = false;
        float h[6];
        const unsigned at = at_next;
        at_next = i + wstride < n ? __ldg(list + i + wstride) : 0u;
        if (i < n) {
            const unsigned zb = at / plane, rem = at - zb * plane;
            const unsigned y = rem / nx, x = rem - y * nx;
            if (div_mode == NB200_DIV_POW2) candidate_hessian<2>(p, (long long)at, (long long)plane, (int)zb, (int)y, (int)x, h);
            else if (div_mode == NB200_DIV_FAST) candidate_hessian<1>(p, (long long)at, (long long)plane, (int)zb, (int)y, (int)x, h);
            else candidate_hessian<0>(p, (long long)at, (long long)plane, (int)zb, (int)y, (int)x, h);
            const float fs = nb::frob_sq3(h[0], h[1], h[2], h[3], h[4], h[5]);
            float tau2 = INFINITY, tau3 = INFINITY;
            if (fs > 1e-20f && fs < 1e20f) {
                const float f2 = 1.001f * fs;
                tau2 = 1e-5f * f2;
                tau3 = 1e-4f * (f2 * sqrtf(f2));
            }
            keep = !nb::pd_reject_full(h[0], h[1], h[2], h[3], h[4], h[5], tau2, tau3);
        }
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if (bits != 0u) {
            if (keep) {
                const int o = n_q + __popc(bits & lt);
#pragma unroll
                for (int j = 0; j < 6; ++j) q[j][o] = h[j];
                q[6][o] = __uint_as_float(at);
            }
            n_q += __popc(bits);
            if (n_q >= 32) solve32(32);
        }
    }
    if (n_q > 0) solve32(n_q);
}

}  // namespace

namespace {
int frangi_sparse_impl(const float* gauss, const float* code, float* acc, const nb200_vol* vol,
                       const float* spacing, int div_mode, float alpha_sq, float beta_sq, const double* sp,
                       unsigned* list, long long list_capacity, unsigned long long* counter,
                       void* stream, int only_if_unsafe) {
    NB_REQUIRE(gauss && code && acc && vol && spacing && sp, NB200_ERR_ARG, "nb200_frangi_sparse: null argument");
    NB_REQUIRE(div_mode >= 0 && div_mode <= 2, NB200_ERR_ARG, "nb200_frangi_sparse: div_mode %d", div_mode);
    const nb200_vol v = *vol;
    NB_REQUIRE(v.ny >= 2 && v.nx >= 2 && v.nz_glob >= 2, NB200_ERR_ARG,
               "nb200_frangi_sparse: every axis needs >= 2 samples (numpy.gradient)");
    NB_REQUIRE(v.zc0 >= 0 && v.zc1 <= v.nz_buf && v.zc0 <= v.zc1, NB200_ERR_ARG, "nb200_frangi_sparse: bad Z window");
    NB_REQUIRE(v.ny < (1 << 22) && v.nx < (1 << 22) && v.nz_buf < (1 << 20), NB200_ERR_UNSUPPORTED,
               "nb200_frangi_sparse: extent too large for the packed queue entries");
    {
        const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off;
        const int need_lo = g0 - 2 < 0 ? 0 : g0 - 2, need_hi = g1 + 1 >= v.nz_glob ? v.nz_glob - 1 : g1 + 1;
        NB_REQUIRE(g0 >= 0 && g1 <= v.nz_glob && need_lo - v.zg_off >= 0 && need_hi - v.zg_off < v.nz_buf,
                   NB200_ERR_ARG, "nb200_frangi_sparse: Z halo of 2 planes missing");
    }
    if (v.zc0 == v.zc1) return NB200_OK;
    SparseParams p;
    p.g = gauss; p.code = code; p.acc = acc; p.v = v;
    for (int a = 0; a < 3; ++a) {
        p.sp.h1[a] = spacing[2 * a]; p.sp.h2[a] = spacing[2 * a + 1];
        p.sp.r1[a] = 1.0f / spacing[2 * a]; p.sp.r2[a] = 1.0f / spacing[2 * a + 1];
    }
    p.div_mode = div_mode;
    p.only_if_unsafe = only_if_unsafe;
    p.alpha_sq = alpha_sq; p.beta_sq = beta_sq; p.spd = sp;
    p.vec_ok = (v.nx % 4 == 0) && (((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(code)) & 15) == 0);
    p.nbx = (v.nx + BX - 1) / BX;
    p.nby = (v.ny + BY - 1) / BY;
    p.bz = BZ;
    p.nbz = (v.zc1 - v.zc0 + BZ - 1) / BZ;
    long long nbricks = (long long)p.nbx * p.nby * p.nbz;
    const long long n_own = (long long)(v.zc1 - v.zc0) * v.ny * v.nx;
    const long long n_buf = (long long)v.nz_buf * v.ny * v.nx;
    if (list != nullptr && counter != nullptr && list_capacity >= n_own && n_buf < (1LL << 32)) {
        cudaStream_t st = nb::as_stream(stream);
        cudaError_t ce = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
        NB_REQUIRE(ce == cudaSuccess, NB200_ERR_CUDA, "nb200_frangi_sparse: %s", cudaGetErrorString(ce));
        // few resident bricks: the list then walks the volume in compact order and the solve kernel's stencil
        // reads stay inside L2 (with 8 CTAs/SM the candidates of > 1000 bricks interleave and DRAM reads triple)
        static const int per_sm = getenv("NB200_STREAM_CTAS") ? atoi(getenv("NB200_STREAM_CTAS")) : 3;
        static const int bz_s = getenv("NB200_STREAM_BZ") ? atoi(getenv("NB200_STREAM_BZ")) : 16;
        p.bz = bz_s > 0 ? bz_s : 16;
        p.nbz = (v.zc1 - v.zc0 + p.bz - 1) / p.bz;
        nbricks = (long long)p.nbx * p.nby * p.nbz;
        const long long cap_s = (long long)(per_sm > 0 ? per_sm : 3) * nb::sm_count();
        sparse_stream_kernel<<<(unsigned)(nbricks < cap_s ? nbricks : cap_s), NT, 0, st>>>(p, list, counter);
        int rc = nb::check_launch("frangi_sparse(stream)");
        if (rc) return rc;
        sparse_solve_kernel<<<(unsigned)(4 * nb::sm_count()), NT, 0, st>>>(p, list, counter);
        return nb::check_launch("frangi_sparse(solve)");
    }
    p.bz = BZ;
    p.nbz = (v.zc1 - v.zc0 + BZ - 1) / BZ;
    nbricks = (long long)p.nbx * p.nby * p.nbz;
    const long long cap = 3LL * nb::sm_count();
    frangi_sparse_kernel<<<(unsigned)(nbricks < cap ? nbricks : cap), NT, 0, nb::as_stream(stream)>>>(p);
    return nb::check_launch("frangi_sparse");
}
}  // namespace

extern "C" int nb200_frangi_sparse(const float* gauss, const float* code, float* acc, const nb200_vol* vol,
                                   const float* spacing, int div_mode, float alpha_sq, float beta_sq, const double* sp,
                                   unsigned* list, long long list_capacity, unsigned long long* counter,
                                   void* stream) {
    return frangi_sparse_impl(gauss, code, acc, vol, spacing, div_mode, alpha_sq, beta_sq, sp, list, list_capacity, counter,
                              stream, 0);
}

extern "C" int nb200_frangi_sparse_gated(const float* gauss, const float* code, float* acc, const nb200_vol* vol,
                                         const float* spacing, int div_mode, float alpha_sq, float beta_sq,
                                         const double* sp, unsigned* list, long long list_capacity,
                                         unsigned long long* counter, void* stream) {
    return frangi_sparse_impl(gauss, code, acc, vol, spacing, div_mode, alpha_sq, beta_sq, sp, list, list_capacity, counter,
                              stream, 1);
}
