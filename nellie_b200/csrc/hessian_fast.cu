// F4-F9, fast form — Hessian statistics (K2) and mask + eigenvalues + vesselness + max/AND (K3) of
// nellie/segmentation/filtering.py:446-562, :407-444, :651-767, :842-851 as ONE kind of kernel: a barrier-free
// Z-march over TMA-staged planes that classifies every voxel with an APPROXIMATE Hessian whose distance from the
// reference's float32 Hessian is bounded by a proven constant, and evaluates the reference's exact arithmetic only
// where the approximation cannot decide.  Results are bit-identical to the exact kernels of frangi.cu / sparse.cu
// (which remain the fallback for odd shapes, exotic value ranges and as the reference form in the tests).
//
// Why: numpy.gradient(numpy.gradient(g)) costs 18 correctly rounded divisions per voxel; the exact march of
// frangi.cu shares them down to 9.3 and still needs ~90 instructions per voxel (issue-bound at 28 % of the HBM
// roofline).  The six second differences themselves are 11 subtractions:
//     N_ab(p) = g(p+ea+eb) - g(p-ea+eb) - g(p+ea-eb) + g(p-ea-eb),        H*_ab = N_ab / (d_a d_b),  d = fl32(2h)
// and with u = 2^-24, G >= |g| everywhere (measured by the statistics pass) the reference value H_ab (two rounded
// subtractions and two correctly rounded divisions per first derivative level) satisfies |H_ab - H*_ab| <= 16.1 u G
// / (d_a d_b), the float32 evaluation of N_ab (three rounded subtractions) |N~_ab - N_ab| <= 8.01 u G, so
//     |N~_ab s_ab - H_ab| <= delta_ab := 40 u G s_ab,      s_ab = RN(1 / (d_a d_b))             (29 u G s proven)
// provided nothing over/underflows, which the value-range flag sp[UNSAFE] of round 1 already guarantees.  From that:
//   * max|H| (statistics): the true maximum is attained in a sub-chunk whose approximate maximum is within
//     2 delta of the approximate global maximum; only those sub-chunks are re-evaluated exactly.
//   * mask frob_sq >= fs_min: |Q(H~) - Q(H)| <= 6 delta sqrt(Q~) + 9 delta^2 (Q = sum of weighted squares, weights
//     sum to 9); nb200_finalize_frob_fast turns that into two thresholds FS_LO < fs_min < FS_HI; voxels in between
//     ("uncertain", ~1e-4 of a frame) are decided with the exact frob_sq.
//   * provably-zero response (sum of two diagonal entries clearly positive => lambda_2 or lambda_3 > 0,
//     devmath.cuh pd_reject_*): the approximate pair sum minus 2.5 delta must exceed 1.4e-5 ||H~||_F, which implies
//     the exact test of round 1 (1e-5 ||H||_F) because ||H||_F <= 1.376 ||H~||_F for voxels above FS_HI.
//   * everything else (alive, passes, not provably zero: the candidates, 5-9 % of a phantom frame) gets the
//     reference's exact Hessian from the staged planes (19 shared-memory reads), the exact tests of sparse.cu and, for
//     the survivors, the float64 eigenvalues + bit-exact vesselness of devmath.cuh — inside the same kernel, in
//     per-warp queues that are drained 32 entries at a time.
//
// Kernel structure (both passes): CTA = 7 warps, tile 128 x 14 columns, march along Z.  Planes (136 x 18 floats,
// halo 4 / 2) arrive by cp.async.bulk.tensor.3d into a ring of D slots with a full[] / empty[] mbarrier pair per
// slot; the warp (p mod 7) issues plane p, every warp releases a plane when it no longer needs it.  There is no
// __syncthreads() after the barrier initialisation: warps only meet through the mbarriers.  A thread owns 2 rows x 4
// X-consecutive voxels; per plane it reads 12 LDS.128 + 12 narrow loads, carries four float4 of first differences in
// registers, and runs the subtractions on the packed FADD2 pipe.
#include <stdlib.h>

#include "common.cuh"
#include "hessian.cuh"
#include "hessian_march.cuh"
#include "march_host.cuh"

#ifndef NB200_STATS_CTAS
#define NB200_STATS_CTAS 2
#define NB200_STATS_D 8
#endif
#ifndef NB200_MARCH_UNROLL
#define NB200_MARCH_UNROLL 1
#endif
#ifndef NB200_SLEEP_PRODUCER
#define NB200_SLEEP_PRODUCER 20000
#endif
#ifndef NB200_SLEEP_CONSUMER
#define NB200_SLEEP_CONSUMER 4000
#endif

namespace {
namespace hf {

constexpr int TX = 128, RW = 2, NW = 8, NT = NW * 32, TYO = NW * RW, GR = TYO + 4, PITCH = TX + 8;
constexpr int SLOT = GR * PITCH;        // 136 x 20 floats = 10880 bytes: a multiple of 128, so every slot is TMA-aligned
static_assert((SLOT * 4) % 128 == 0, "ring slots must stay 128-byte aligned");
constexpr unsigned BOX_BYTES = GR * PITCH * 4;
constexpr int MARCH_UNROLL = NB200_MARCH_UNROLL;
constexpr int ZR = 32;                  // planes per max|H| record of the statistics pass
constexpr unsigned WL_CAP = 16384;      // work-list entries (8 words each)
constexpr int WL_HDR = 8;
constexpr float U24 = 5.9604644775390625e-08f;   // 2^-24

struct Consts {
    float d2[3], r2[3];      // fl32(2h) and RN(1/fl32(2h)) for Z, Y, X (exact path)
    float s[6];              // s_ab = RN(1/(d_a d_b)) for zz, zy, zx, yy, yx, xx
    float smax;
    int iso;                 // all six scales equal: the approximate tests run in numerator units
};

template <int D>
struct Ring {
    float g[D][SLOT];                   // TMA destinations, 128-byte aligned
    unsigned long long full[D];         // "plane landed" mbarriers (TMA complete_tx)
    unsigned released[D];               // warps that no longer need the plane in this slot
};

// ---- mbarrier / TMA helpers on shared-window addresses (bounded wait: a protocol bug traps instead of hanging) ------
// A try_wait without a time hint returns almost at once on this part when the phase is not complete: the bare poll loops
// issued 22 % of all instructions of the K3 kernel (ncu: 1 active thread in the producer's loop), and NANOSLEEP 200
// between polls slept only ~20 ns.  The suspend-time hint (NS, nanoseconds) parks the thread in hardware instead.
template <int NS>
__device__ __forceinline__ void mbar_wait_a(unsigned a, unsigned parity) {
    for (unsigned it = 0;; ++it) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity), "r"((unsigned)NS)       // suspend-time hint in ns: the hardware parks the thread
            : "memory");
        if (ok) return;
        if (it > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_plane_a(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BOX_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- packed float32x2 helpers ------------------------------------------------------------------------------------
struct V4 {
    float2 lo, hi;
};
__device__ __forceinline__ V4 ld4(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    V4 v;
    v.lo = make_float2(t.x, t.y);
    v.hi = make_float2(t.z, t.w);
    return v;
}
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ V4 sub(const V4& a, const V4& b) {
    V4 r;
    r.lo = sub2(a.lo, b.lo);
    r.hi = sub2(a.hi, b.hi);
    return r;
}
// central first difference along X of 4 consecutive values (l = value left of v.lo.x, r = value right of v.hi.y)
__device__ __forceinline__ V4 dxrow(const V4& v, float l, float r) {
    V4 d;
    d.lo = make_float2(v.lo.y - l, v.hi.x - v.lo.x);
    d.hi = make_float2(v.hi.y - v.lo.y, r - v.hi.x);
    return d;
}
__device__ __forceinline__ float max3abs(float m, float a, float b) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- the reference's exact interior Hessian (central differences only) -------------------------------------------
// (a - b) / fl32(2h): MODE 1 = verified constant-divisor sequence, MODE 2 = exact reciprocal product, 0 = IEEE
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
// LD(dz, dy, dx) returns the blurred value at an offset; same operations in the same order as nb::hessian3 /
// numpy.gradient(numpy.gradient(g)) for a voxel at least two samples away from every frame border.
template <int MODE, class LD>
__device__ __forceinline__ void hess_interior(const LD& L, const Consts& k, float (&h)[6]) {
    const float dz = k.d2[0], dy = k.d2[1], dx = k.d2[2], rz = k.r2[0], ry = k.r2[1], rx = k.r2[2];
    const float g0 = L(0, 0, 0);
    const float zp2 = L(2, 0, 0), zm2 = L(-2, 0, 0), yp2 = L(0, 2, 0), ym2 = L(0, -2, 0), xp2 = L(0, 0, 2), xm2 = L(0, 0, -2);
    const float zpyp = L(1, 1, 0), zpym = L(1, -1, 0), zmyp = L(-1, 1, 0), zmym = L(-1, -1, 0);
    const float zpxp = L(1, 0, 1), zpxm = L(1, 0, -1), zmxp = L(-1, 0, 1), zmxm = L(-1, 0, -1);
    const float ypxp = L(0, 1, 1), ypxm = L(0, 1, -1), ymxp = L(0, -1, 1), ymxm = L(0, -1, -1);
    h[0] = cdiff<MODE>(cdiff<MODE>(zp2, g0, dz, rz), cdiff<MODE>(g0, zm2, dz, rz), dz, rz);            // d0 d0
    h[1] = cdiff<MODE>(cdiff<MODE>(zpyp, zmyp, dz, rz), cdiff<MODE>(zpym, zmym, dz, rz), dy, ry);      // d1 d0
    h[2] = cdiff<MODE>(cdiff<MODE>(zpxp, zmxp, dz, rz), cdiff<MODE>(zpxm, zmxm, dz, rz), dx, rx);      // d2 d0
    h[3] = cdiff<MODE>(cdiff<MODE>(yp2, g0, dy, ry), cdiff<MODE>(g0, ym2, dy, ry), dy, ry);            // d1 d1
    h[4] = cdiff<MODE>(cdiff<MODE>(ypxp, ymxp, dy, ry), cdiff<MODE>(ypxm, ymxm, dy, ry), dx, rx);      // d2 d1
    h[5] = cdiff<MODE>(cdiff<MODE>(xp2, g0, dx, rx), cdiff<MODE>(g0, xm2, dx, rx), dx, rx);            // d2 d2
}

// staged planes: pl[j] = base of the ring slot holding plane o-2+j, `off` = row * PITCH + column of the voxel
struct SmemLoad {
    const float* pl[5];
    int off;
    __device__ __forceinline__ float operator()(int dz, int dy, int dx) const { return pl[2 + dz][off + dy * PITCH + dx]; }
};
struct GlobalLoad {
    const float* p;
    long long plane;
    int nx;
    __device__ __forceinline__ float operator()(int dz, int dy, int dx) const {
        return __ldg(p + dz * plane + (long long)dy * nx + dx);
    }
};

struct Geo {
    nb200_vol v;
    int x0, y0;
};

// approximate second differences (numerator units unless scaled) of a thread's 2 rows x 4 voxels at one plane
struct Num {
    V4 zz[RW], zy[RW], zx[RW], yy[RW], yx[RW], xx[RW];
};

// per-step view handed to the epilogues
struct StepCtx {
    const float* ring;       // base of slot 0
    int sl[5];               // ring slots of planes o-2 .. o+2
    int off0;                // row * PITCH + column of the thread's first voxel inside a slot
    int o;                   // global plane index of the output plane
};

// ------------------------------------------------------------------------------------------------------------------
// The march over global planes [zs, ze) (all interior).  D = ring depth (power of two), K = planes a released slot
// lags behind (the epilogue may still read planes o-2-K .. o+2 during step o).
// Warp NW (the eighth) is the producer: one lane issues plane p as soon as its slot is free.  Warps 0..NW-1 consume.
// Epi provides: prefetch(o), track(V4) [range of staged values], put<C>(row, V4) [second difference C of a row, as
// soon as it is known], step(ctx) [end of the plane], finish().
// The loop is deliberately NOT unrolled over the ring: the unrolled form (immediate shared-memory offsets) was 245 KB
// of code for K3 and ran 30x slower — stall_no_instruction, the 32 KB instruction cache thrashed by divergent warps.
// ------------------------------------------------------------------------------------------------------------------
// Ring protocol.  full[slot] is an mbarrier completed by the TMA transaction bytes of the plane.  There is no "empty"
// barrier and no producer warp: every warp counts itself off a plane in released[slot] (shared-memory atomic), and the
// warp that completes the count issues the TMA of the next plane for that slot at once.  (A dedicated producer lane
// polling an empty-mbarrier issued 14 % of the kernel's instructions, and the eighth warp slot of the CTA was wasted.)
template <int D>
__device__ __forceinline__ void ring_init(Ring<D>& rg, const CUtensorMap* map, const Geo& q, int zs, int ze) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            hm::mbar_init(&rg.full[k], 1);
            rg.released[k] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // first D planes: their slots are free
        const int n_pl = ze - zs + 4;
        const unsigned sb = hm::smem_u32(&rg);
        const unsigned sb_full = sb + (unsigned)(sizeof(float) * D * SLOT);
        for (int p = 0; p < D && p < n_pl; ++p)
            tma_plane_a(sb + (unsigned)p * (unsigned)(SLOT * 4), map, sb_full + 8u * (unsigned)p, q.x0 - 4, q.y0 - 2,
                        zs - 2 - q.v.zg_off + p);
    }
    __syncthreads();                              // the only CTA-wide barrier of the kernel
}

template <int D, int K, class Epi>
__device__ __forceinline__ void consume(Ring<D>& rg, const CUtensorMap* map, const Geo& q, int zs, int ze, Epi& epi) {
    static_assert(D >= 6 + K, "ring too shallow");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_out = ze - zs, n_pl = n_out + 4;
    const int tx0 = q.x0 - 4, ty0 = q.y0 - 2, zb0 = zs - 2 - q.v.zg_off;
    const int off0 = (RW * warp + 2) * PITCH + 4 + 4 * lane;
    const float* ring = &rg.g[0][0];
    const unsigned sb = hm::smem_u32(&rg);
    const unsigned sb_full = sb + (unsigned)(sizeof(float) * D * SLOT);
    auto wait_full = [&](int p) { mbar_wait_a<NB200_SLEEP_CONSUMER>(sb_full + 8u * ((unsigned)p % D), ((unsigned)p / D) & 1u); };
    V4 a0[RW], a1[RW], dyP[RW], dxP[RW], dyQ[RW], dxQ[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) { dyQ[r] = V4(); dxQ[r] = V4(); }
    {
        // carries of the first step: a_o, a_{o+1}, first differences of plane o-1 (planes 0..3 sit in slots 0..3)
#pragma unroll
        for (int j = 0; j < 4; ++j) wait_full(j);
        const float* Q0 = ring + off0;
        const float* Q1 = ring + SLOT + off0;
        const float* Q2 = ring + 2 * SLOT + off0;
        const float* Q3 = ring + 3 * SLOT + off0;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const V4 f0 = ld4(Q0 + r * PITCH), f1 = ld4(Q1 + r * PITCH), f2 = ld4(Q2 + r * PITCH), f3 = ld4(Q3 + r * PITCH);
            epi.track(f0); epi.track(f1); epi.track(f2); epi.track(f3);
            a0[r] = sub(f2, f0);
            a1[r] = sub(f3, f1);
            const V4 up = ld4(Q1 + (r + 1) * PITCH), dn = ld4(Q1 + (r - 1) * PITCH);
            epi.track(up); epi.track(dn);       // rows y-1 / y+2 of the plane below the chunk (slab halo planes)
            dyP[r] = sub(up, dn);
            dxP[r] = dxrow(f1, Q1[r * PITCH - 1], Q1[r * PITCH + 4]);
        }
    }
    const float* base = ring + off0;
#pragma unroll MARCH_UNROLL
    for (int i = 0; i < n_out; ++i) {
        epi.prefetch(zs + i);
        StepCtx cx;
        cx.ring = ring;
        cx.off0 = off0;
        cx.o = zs + i;
#pragma unroll
        for (int t = 0; t < 5; ++t) cx.sl[t] = (int)((unsigned)(i + t) % D);
        wait_full(i + 4);
        const float* P0 = base + cx.sl[2] * SLOT;     // plane o
        const float* P1 = base + cx.sl[3] * SLOT;     // plane o+1
        const float* P2 = base + cx.sl[4] * SLOT;     // plane o+2
        // ---- plane o: rows -2 .. 3 ----
        const V4 m2 = ld4(P0 - 2 * PITCH), m1 = ld4(P0 - PITCH), r0 = ld4(P0), r1 = ld4(P0 + PITCH);
        const V4 p1 = ld4(P0 + 2 * PITCH), p2 = ld4(P0 + 3 * PITCH);
        const V4 dym = sub(r0, m2), dy0 = sub(r1, m1), dy1 = sub(p1, r0), dy2 = sub(p2, r1);
        epi.template put<3>(0, sub(dy1, dym));
        epi.template put<3>(1, sub(dy2, dy0));
        const float2 L0 = ld2(P0 - 2), R0 = ld2(P0 + 4), L1 = ld2(P0 + PITCH - 2), R1 = ld2(P0 + PITCH + 4);
        const V4 dxm = dxrow(m1, P0[-PITCH - 1], P0[-PITCH + 4]);
        const V4 dx0 = dxrow(r0, L0.y, R0.x), dx1 = dxrow(r1, L1.y, R1.x);
        const V4 dx2 = dxrow(p1, P0[2 * PITCH - 1], P0[2 * PITCH + 4]);
        epi.template put<4>(0, sub(dx1, dxm));
        epi.template put<4>(1, sub(dx2, dx0));
        {
            const float2 eA = sub2(r0.lo, L0), eB = sub2(r0.hi, r0.lo), eC = sub2(R0, r0.hi);
            V4 xx;
            xx.lo = sub2(eB, eA);
            xx.hi = sub2(eC, eB);
            epi.template put<5>(0, xx);
        }
        {
            const float2 eA = sub2(r1.lo, L1), eB = sub2(r1.hi, r1.lo), eC = sub2(R1, r1.hi);
            V4 xx;
            xx.lo = sub2(eB, eA);
            xx.hi = sub2(eC, eB);
            epi.template put<5>(1, xx);
        }
        // ---- plane o+2: own rows (zz through a_p = g(p) - g(p-2)) ----
        {
            const V4 t0 = ld4(P2), t1 = ld4(P2 + PITCH);
            epi.track(t0);
            epi.track(t1);
            const V4 a20 = sub(t0, r0), a21 = sub(t1, r1);
            epi.template put<0>(0, sub(a20, a0[0]));
            epi.template put<0>(1, sub(a21, a0[1]));
            a0[0] = a1[0]; a0[1] = a1[1];
            a1[0] = a20;   a1[1] = a21;
        }
        if (!Epi::kLagZ) {
            // ---- plane o+1: first differences of the own rows; mixed Z derivatives against the carried plane o-1 ----
            const V4 qm = ld4(P1 - PITCH), q0 = ld4(P1), q1 = ld4(P1 + PITCH), qp = ld4(P1 + 2 * PITCH);
            if (i == n_out - 1) { epi.track(qm); epi.track(qp); }   // plane above the chunk: rows owned by nobody else
            const V4 dyN0 = sub(q1, qm), dyN1 = sub(qp, q0);
            const V4 dxN0 = dxrow(q0, P1[-1], P1[4]), dxN1 = dxrow(q1, P1[PITCH - 1], P1[PITCH + 4]);
            epi.template put<1>(0, sub(dyN0, dyP[0]));
            epi.template put<1>(1, sub(dyN1, dyP[1]));
            epi.template put<2>(0, sub(dxN0, dxP[0]));
            epi.template put<2>(1, sub(dxN1, dxP[1]));
            dyP[0] = dy0; dyP[1] = dy1;
            dxP[0] = dx0; dxP[1] = dx1;
        } else {
            // ---- lagged form (epilogues that take every component on its own, i.e. the statistics): the mixed Z
            //      derivatives of plane o-1 from the first differences of planes o (just computed) and o-2 (carried).
            //      No load from plane o+1: a third of the shared-memory wavefronts of the step. ----
            if (i > 0) {
                epi.template put<1>(0, sub(dy0, dyQ[0]));
                epi.template put<1>(1, sub(dy1, dyQ[1]));
                epi.template put<2>(0, sub(dx0, dxQ[0]));
                epi.template put<2>(1, sub(dx1, dxQ[1]));
            }
            dyQ[0] = dyP[0]; dyQ[1] = dyP[1]; dxQ[0] = dxP[0]; dxQ[1] = dxP[1];
            dyP[0] = dy0; dyP[1] = dy1; dxP[0] = dx0; dxP[1] = dx1;
            if (i == n_out - 1) {                           // last plane of the chunk: its own mixed derivatives now
                const V4 qm = ld4(P1 - PITCH), q0 = ld4(P1), q1 = ld4(P1 + PITCH), qp = ld4(P1 + 2 * PITCH);
                epi.track(qm); epi.track(qp);
                const V4 dyN0 = sub(q1, qm), dyN1 = sub(qp, q0);
                const V4 dxN0 = dxrow(q0, P1[-1], P1[4]), dxN1 = dxrow(q1, P1[PITCH - 1], P1[PITCH + 4]);
                epi.template put<1>(0, sub(dyN0, dyQ[0]));
                epi.template put<1>(1, sub(dyN1, dyQ[1]));
                epi.template put<2>(0, sub(dxN0, dxQ[0]));
                epi.template put<2>(1, sub(dxN1, dxQ[1]));
            }
        }
        epi.step(cx);
        // ---- release plane i-K; the last warp to let go of it refills the slot with plane i-K+D ----
        if (i >= K) {
            __syncwarp();
            if (lane == 0) {
                const int pr = i - K;
                const unsigned slot = (unsigned)pr % D;
                __threadfence_block();                       // this warp's reads of the slot are done before the count
                if (atomicAdd(&rg.released[slot], 1u) == NW - 1) {
                    atomicExch(&rg.released[slot], 0u);      // next use of the counter comes after the refill landed
                    if (pr + D < n_pl)
                        tma_plane_a(sb + slot * (unsigned)(SLOT * 4), map, sb_full + 8u * slot, tx0, ty0, zb0 + pr + D);
                }
            }
        }
    }
    epi.finish();
}

// ==================================================================================================================
// pass A: statistics
// ==================================================================================================================
// sqrt(frob_sq) of the reference's exact Hessian at the lattice points of a thread's 2 x 4 voxels (out of line: rare)
template <int MODE>
__device__ __noinline__ void lattice_samples(const float* ring, int s0, int s1, int s2, int s3, int s4, int off0,
                                             const Consts& k, int ylat0, int ylat1, int xmask, long long zrow, int xlat,
                                             int lx_n, float* __restrict__ frob_samples) {
    SmemLoad L;
    L.pl[0] = ring + s0 * SLOT; L.pl[1] = ring + s1 * SLOT; L.pl[2] = ring + s2 * SLOT;
    L.pl[3] = ring + s3 * SLOT; L.pl[4] = ring + s4 * SLOT;
    for (int r = 0; r < RW; ++r) {
        const int yl = r ? ylat1 : ylat0;
        if (yl < 0) continue;
        long long idx = (zrow + yl) * lx_n + xlat;
        for (int kk = 0; kk < 4; ++kk) {
            if (!(xmask >> kk & 1)) continue;
            L.off = off0 + r * PITCH + kk;
            float h[6];
            hess_interior<MODE>(L, k, h);
            frob_samples[idx++] = sqrtf(nb::frob_sq3(h[0], h[1], h[2], h[3], h[4], h[5]));
        }
    }
}
struct StatsFastParams {
    int sz, sy, sx, g_first, ly_n, lx_n;
    float* frob_samples;
    long long* hstats;
    unsigned* wl;            // work list: [0] count, then WL_HDR.. entries of 8 words
};

// one record per warp and ZR planes: fold the warp's approximate maximum into the global one and append the sub-chunk
// to the work list when it is within 2^-10 of the running maximum (out of line: once per 32 planes)
__device__ __noinline__ void record_max(float mx, bool nan, long long* hstats, unsigned* wl, int x0, int y, int z0, int z1) {
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const bool any_nan = __any_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) != 0) return;
    if (any_nan) {
        atomicMax((unsigned long long*)&hstats[NB200_HS_FALLBACK], 1ull);
    } else if (mx > 0.0f) {
        const unsigned long long old =
            atomicMax((unsigned long long*)&hstats[NB200_HS_APPROX_MAX_BITS], (unsigned long long)__float_as_uint(mx));
        const float run = fmaxf(__uint_as_float((unsigned)old), mx);
        if (mx >= run * 0.9990234375f) {               // within 2^-10 of the running maximum
            const unsigned e = atomicAdd(&wl[0], 1u);
            if (e < WL_CAP) {
                unsigned* w = wl + WL_HDR + 8ull * e;
                w[0] = (unsigned)x0;
                w[1] = (unsigned)y;
                w[2] = (unsigned)z0;
                w[3] = (unsigned)z1;
                w[4] = __float_as_uint(mx);
            }
        }
    }
}

template <int MODE>
struct StatsFastEpi {
    static constexpr bool kLagZ = true;      // components are folded one by one: the march may emit them plane-shifted
    const StatsFastParams& p;
    const Consts& k;
    const Geo& q;
    float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g_max = 0.0f;
    unsigned g_minb = 0xffffffffu;      // (bits of the smallest non-zero |g|) - 1
    bool ok[RW];
    int xmask = 0, xlat = 0, ylat[RW];
    long long zrow = -1;
    int zmod = -1, zq = 0;
    int rec_planes = 0, rec_z0, ze;
    int lane, warp;
    __device__ StatsFastEpi(const StatsFastParams& p_, const Consts& k_, const Geo& q_, int zs, int ze_)
        : p(p_), k(k_), q(q_), rec_z0(zs), ze(ze_) {
        // lattice plane counters positioned one plane before zs
        zmod = (zs - 1) % p.sz;
        zq = ((zs - 1) - zmod - p.g_first) / p.sz;      // may be negative before the first lattice plane (exact division)
        if (((zs - 1) - zmod - p.g_first) < 0) zq = -(((p.g_first - ((zs - 1) - zmod))) / p.sz);
        lane = threadIdx.x & 31;
        warp = threadIdx.x >> 5;
        const int x = q.x0 + 4 * lane;
        const int xhi = 4 * ((q.v.nx - 2) / 4);
        const bool col_valid = x >= 4 && x < xhi;
        for (int kk = 0; kk < 4; ++kk) xmask |= ((x + kk) % p.sx == 0) ? (1 << kk) : 0;
        xlat = (x + p.sx - 1) / p.sx;
        for (int r = 0; r < RW; ++r) {
            const int y = q.y0 + RW * warp + r;
            ok[r] = col_valid && y >= 2 && y <= q.v.ny - 3;
            ylat[r] = (p.frob_samples && xmask && ok[r] && (y % p.sy == 0)) ? y / p.sy : -1;
        }
    }
    __device__ __forceinline__ void prefetch(int) {}
    __device__ __forceinline__ void track1(float v) {
        const float a = fabsf(v);
        g_max = fmaxf(g_max, a);
        g_minb = min(g_minb, __float_as_uint(a) - 1u);       // zero wraps to 0xffffffff and never wins
    }
    __device__ __forceinline__ void track(const V4& v) {
        track1(v.lo.x); track1(v.lo.y); track1(v.hi.x); track1(v.hi.y);
    }
    __device__ __forceinline__ void plane(int) {                // consecutive planes: counters only
        if (++zmod == p.sz) {
            zmod = 0;
            ++zq;
        }
        zrow = zmod == 0 ? (long long)zq * p.ly_n : -1;
    }
    __device__ __forceinline__ void fold(float& mm, const V4& a) {
        mm = max3abs(mm, a.lo.x, a.lo.y);
        mm = max3abs(mm, a.hi.x, a.hi.y);
    }
    __device__ __forceinline__ void flush_record(int z_end) {
        float mx = fmaxf(fmaxf(fmaxf(k.s[0] * m[0], k.s[1] * m[1]), fmaxf(k.s[2] * m[2], k.s[3] * m[3])),
                         fmaxf(k.s[4] * m[4], k.s[5] * m[5]));
        // NaN must not get lost in fmaxf: it has to reach the record so that the fallback takes over
        const bool nan = (m[0] != m[0]) || (m[1] != m[1]) || (m[2] != m[2]) || (m[3] != m[3]) || (m[4] != m[4]) || (m[5] != m[5]);
        record_max(mx, nan, p.hstats, p.wl, q.x0, q.y0 + RW * warp, rec_z0, z_end);
#pragma unroll
        for (int c = 0; c < 6; ++c) m[c] = 0.0f;
        rec_z0 = z_end;
        rec_planes = 0;
    }
    template <int C>
    __device__ __forceinline__ void put(int r, const V4& a) {      // component C of row r: fold into the maxima at once
        if (ok[r]) fold(m[C], a);
    }
    __device__ __forceinline__ void step(const StepCtx& cx) {
        plane(cx.o);
        if (zrow >= 0 && (ylat[0] >= 0 || ylat[1] >= 0))          // lattice plane and row: a few threads
            lattice_samples<MODE>(cx.ring, cx.sl[0], cx.sl[1], cx.sl[2], cx.sl[3], cx.sl[4], cx.off0, k, ylat[0], ylat[1],
                                  xmask, zrow, xlat, p.lx_n, p.frob_samples);
        if (++rec_planes == ZR || cx.o + 1 == ze) flush_record(cx.o + 1);
    }
    __device__ void finish() {
        for (int o = 16; o > 0; o >>= 1) {
            g_max = fmaxf(g_max, __shfl_xor_sync(0xffffffffu, g_max, o));
            g_minb = min(g_minb, __shfl_xor_sync(0xffffffffu, g_minb, o));
        }
        if (lane == 0) {
            if (g_minb != 0xffffffffu)
                atomicMax((unsigned long long*)&p.hstats[NB200_HS_MIN_NZ_COMPL], (unsigned long long)(0x7f800000u - (g_minb + 1u)));
            float gm = g_max;
            if (!(gm <= 3.0e38f)) gm = INFINITY;
            atomicMax((unsigned long long*)&p.hstats[NB200_HS_MAX_G_BITS], (unsigned long long)__float_as_uint(gm));
        }
    }
};

// NaN blurred values: fmaxf drops NaN, so g_max alone would not notice them.  The approximate numerators of the
// voxels around a NaN are NaN, and flush_record raises FALLBACK when a maximum is NaN; a NaN that only touches
// invalid lanes sits in the border shell, whose exact kernel propagates it (store_range).

template <int MODE, int D>
__global__ void __launch_bounds__(NT, NB200_STATS_CTAS)
stats_fast_kernel(const __grid_constant__ CUtensorMap map, nb200_vol v, Consts k, int zi0, int zi1, int zchunk,
                  StatsFastParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Ring<D>& rg = *reinterpret_cast<Ring<D>*>(smem_raw);
    const int ntx = (v.nx + TX - 1) / TX, nty = (v.ny - 4 + TYO - 1) / TYO;
    long long b = blockIdx.x;
    const int bx = (int)(b % ntx); b /= ntx;
    const int by = (int)(b % nty); b /= nty;
    Geo q;
    q.v = v;
    q.x0 = bx * TX;
    q.y0 = 2 + by * TYO;
    const int zs = zi0 + (int)b * zchunk;
    const int ze = min(zs + zchunk, zi1);
    if (zs >= ze) return;
    ring_init(rg, &map, q, zs, ze);
    StatsFastEpi<MODE> epi(p, k, q, zs, ze);
    consume<D, 0>(rg, &map, q, zs, ze, epi);
}

__global__ void wl_reset_kernel(unsigned* wl) {
    if (threadIdx.x < WL_HDR) wl[threadIdx.x] = 0u;
}

// exact re-evaluation of the listed sub-chunks that can hold the true maximum
template <int MODE>
__global__ void __launch_bounds__(256)
stats_fixup_kernel(const float* __restrict__ g, nb200_vol v, Consts k, const unsigned* __restrict__ wl,
                   long long* __restrict__ hstats, int zi0) {
    __shared__ float red[8];
    const unsigned n_all = wl[0];
    const float mt = __uint_as_float((unsigned)hstats[NB200_HS_APPROX_MAX_BITS]);
    const float gmax = __uint_as_float((unsigned)hstats[NB200_HS_MAX_G_BITS]);
    const float delta = 40.0f * U24 * gmax * k.smax;
    // the list only holds sub-chunks within 2^-10 of the running maximum: the argument needs 2 delta <= 2^-10 max
    const bool good = n_all <= WL_CAP && (2.0f * delta <= mt * 0.0009765625f) && hstats[NB200_HS_FALLBACK] == 0;
    if (!good) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax((unsigned long long*)&hstats[NB200_HS_FALLBACK], 1ull);
        return;
    }
    const float thr = mt - 2.0f * delta;
    const long long plane = (long long)v.ny * v.nx;
    const int xhi = 4 * ((v.nx - 2) / 4);
    for (unsigned e = blockIdx.x; e < n_all; e += gridDim.x) {
        const unsigned* w = wl + WL_HDR + 8ull * e;
        if (__uint_as_float(w[4]) < thr) continue;
        // the record of planes [z0, z1) also holds the mixed Z derivatives of plane z0-1 (lagged emission of the march)
        const int x0 = (int)w[0], y0 = (int)w[1], z0 = max((int)w[2] - 1, zi0), z1 = (int)w[3];
        float mx = 0.0f;
        const int total = (z1 - z0) * RW * TX;
        for (int t = threadIdx.x; t < total; t += blockDim.x) {
            const int x = x0 + (t % TX), r = (t / TX) % RW, z = z0 + t / (TX * RW);
            const int y = y0 + r;
            if (x < 4 || x >= xhi || y < 2 || y > v.ny - 3) continue;
            GlobalLoad L{g + (long long)(z - v.zg_off) * plane + (long long)y * v.nx + x, plane, v.nx};
            float h[6];
            hess_interior<MODE>(L, k, h);
            mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(h[0]), fabsf(h[1])), fmaxf(fabsf(h[2]), fabsf(h[3]))),
                                 fmaxf(fabsf(h[4]), fabsf(h[5]))));
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int j = 1; j < 8; ++j) mx = fmaxf(mx, red[j]);
            atomicMax((unsigned long long*)&hstats[NB200_HS_MAX_ABS_BITS], (unsigned long long)__float_as_uint(mx));
        }
        __syncthreads();
    }
}

// ==================================================================================================================
// pass B: mask + eigenvalues + vesselness + max/AND
// ==================================================================================================================
__device__ __noinline__ float eig_vesselness(float a00, float a01, float a02, float a11, float a12, float a22,
                                             float alpha_sq, float beta_sq, float gamma_sq) {
    float l1, l2, l3;
    nb::eig3_sym<2>(a00, a01, a02, a11, a12, a22, l1, l2, l3);
    if (l3 > 0.0f || l2 > 0.0f) return 0.0f;           // filtering.py:759-761 zeroes these responses
    return nb::vesselness3(l1, l2, l3, alpha_sq, beta_sq, gamma_sq);
}

constexpr int CQ_CAP = 512;           // candidate ring per warp (u8 entries: row << 7 | column): < 32 left over + <= 256 per step
constexpr int RQ_CAP = 64;            // survivor stack per warp (< 32 left over + <= 32 pushed)
struct WarpQueues {
    float rq[7][RQ_CAP];              // six second derivatives + voxel index (uint32 bits)
    unsigned char cq[CQ_CAP];         // ring indices are taken modulo CQ_CAP
};

struct FrangiFastParams {
    float* acc;
    float alpha_sq, beta_sq;
    const double* spd;
    unsigned long long* diag;
    int debug;               // profiling experiments only (NB200_K3_DEBUG): 1 = no eigen-solves, 2 = no exact evaluation at all
};

template <int MODE, int K>
struct FrangiFastEpi {
    static constexpr bool kLagZ = false;     // needs the six components of a voxel together
    const FrangiFastParams& p;
    const Consts& k;
    const Geo& q;
    WarpQueues& wq;
    float gamma_sq, fs_min, t_lo, t_hi, zc;
    bool ok[RW];
    V4 cur[RW];
    long long row_idx[RW];           // linear index of the thread's first voxel of each row at buffer plane 0
    long long plane;
    int lane, warp;
    int cq_head = 0, cq_n = 0, cq_old = 0;      // warp-uniform
    int rq_n = 0;                               // warp-uniform
    int prev_sl[5];
    unsigned long long d_cand = 0, d_kill = 0, d_surv = 0;
    __device__ FrangiFastEpi(const FrangiFastParams& p_, const Consts& k_, const Geo& q_, WarpQueues* wqs)
        : p(p_), k(k_), q(q_), wq(wqs[threadIdx.x >> 5]) {
        lane = threadIdx.x & 31;
        warp = threadIdx.x >> 5;
        gamma_sq = (float)p.spd[NB200_SP_GAMMA_SQ];
        fs_min = (float)p.spd[NB200_SP_FROBSQ_MIN];
        // thresholds in the units the approximate tests run in (numerator units when all scales are equal);
        // rounded away from fs_min so that the conversion never weakens them
        const double s2 = k.iso ? (double)k.s[0] * (double)k.s[0] : 1.0;
        const double lo = p.spd[NB200_SP_FS_LO] / s2, hi = p.spd[NB200_SP_FS_HI] / s2;
        t_lo = __double2float_rd(lo * (lo > 0.0 ? 0.9999995 : 1.0000005));
        t_hi = __double2float_ru(hi * (hi > 0.0 ? 1.0000005 : 0.9999995));
        zc = __double2float_ru(p.spd[NB200_SP_ZT_C] / (k.iso ? (double)k.s[0] : 1.0) * 1.0000005);
        const int x = q.x0 + 4 * lane;
        const int xhi = 4 * ((q.v.nx - 2) / 4);
        plane = (long long)q.v.ny * q.v.nx;
        for (int r = 0; r < RW; ++r) {
            const int y = q.y0 + RW * warp + r;
            ok[r] = x >= 4 && x < xhi && y >= 2 && y <= q.v.ny - 3;
            row_idx[r] = (long long)y * q.v.nx + x;
        }
    }
    __device__ __forceinline__ void track(const V4&) {}
    __device__ __forceinline__ void prefetch(int o) {
        const long long zoff = (long long)(o - q.v.zg_off) * plane;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            float4 a = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
            if (ok[r]) a = *reinterpret_cast<const float4*>(p.acc + zoff + row_idx[r]);
            cur[r].lo = make_float2(a.x, a.y);
            cur[r].hi = make_float2(a.z, a.w);
        }
    }
    // eigenvalues + vesselness for the top `cnt` (<= 32) survivors of this warp
    __device__ __forceinline__ void solve(int cnt) {
        __syncwarp();
        const int e = rq_n - cnt + lane;
        if (lane < cnt && p.debug == 0) {
            const unsigned at = __float_as_uint(wq.rq[6][e]);
            const float vv = eig_vesselness(wq.rq[0][e], wq.rq[1][e], wq.rq[2][e], wq.rq[3][e], wq.rq[4][e], wq.rq[5][e],
                                            p.alpha_sq, p.beta_sq, gamma_sq);
            // acc >= 0 here (only live voxels become candidates) and vv >= 0: non-negative floats order like their bit
            // patterns, so the running maximum is one fire-and-forget integer RED — no load in front of the solve
            if (vv > 0.0f) atomicMax(reinterpret_cast<int*>(p.acc) + at, __float_as_int(vv));
        }
        rq_n -= cnt;
        __syncwarp();
    }
    // exact evaluation of the `m` (<= 32) oldest candidates
    __device__ __forceinline__ void batch(const StepCtx& cx, int m) {
        __syncwarp();
        bool keep = false;
        float h[6];
        unsigned at = 0;
        if (lane < m) {
            const unsigned e = wq.cq[(cq_head + lane) & (CQ_CAP - 1)];
            const int r = (int)(e >> 7), col = (int)(e & 127u);
            const bool old = K > 0 && lane < cq_old;          // candidate of the previous plane
            SmemLoad L;
#pragma unroll
            for (int j = 0; j < 5; ++j) L.pl[j] = cx.ring + (old ? prev_sl[j] : cx.sl[j]) * SLOT;
            L.off = (RW * warp + 2 + r) * PITCH + 4 + col;
            hess_interior<MODE>(L, k, h);
            const float fs = nb::frob_sq3(h[0], h[1], h[2], h[3], h[4], h[5]);
            const int zb = cx.o - (old ? 1 : 0) - q.v.zg_off;
            at = (unsigned)((long long)zb * plane + (long long)(q.y0 + RW * warp + r) * q.v.nx + (q.x0 + col));
            if (!(fs >= fs_min)) {
                p.acc[at] = -1.0f;                             // an uncertain voxel that fails the mask (NaN fails too)
                ++d_kill;
            } else {
                // the exact provably-zero tests of sparse.cu / frangi.cu (voxel_code + pd_reject_full)
                const float mm = fmaxf(fmaxf(h[0] + h[3], h[0] + h[5]), h[3] + h[5]);
                const bool in_range = fs > 1e-20f && fs < 1e20f;
                const bool zero = mm > 0.0f && mm * mm > 1.001e-10f * fs && in_range;
                float tau2 = INFINITY, tau3 = INFINITY;
                if (in_range) {
                    const float f2 = 1.001f * fs;
                    tau2 = 1e-5f * f2;
                    tau3 = 1e-4f * (f2 * sqrtf(f2));
                }
                keep = !zero && !nb::pd_reject_full(h[0], h[1], h[2], h[3], h[4], h[5], tau2, tau3);
            }
        }
        cq_head = (cq_head + m) & (CQ_CAP - 1);
        cq_n -= m;
        cq_old = cq_old > m ? cq_old - m : 0;
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if (bits != 0u) {
            if (keep) {
                const int o = rq_n + __popc(bits & ((1u << lane) - 1u));
#pragma unroll
                for (int j = 0; j < 6; ++j) wq.rq[j][o] = h[j];
                wq.rq[6][o] = __uint_as_float(at);
                ++d_surv;
            }
            rq_n += __popc(bits);
            if (rq_n >= 32) solve(32);
        }
    }
    Num n;
    template <int C>
    __device__ __forceinline__ void put(int r, const V4& a) {
        if (C == 0) n.zz[r] = a;
        else if (C == 1) n.zy[r] = a;
        else if (C == 2) n.zx[r] = a;
        else if (C == 3) n.yy[r] = a;
        else if (C == 4) n.yx[r] = a;
        else n.xx[r] = a;
    }
    __device__ __forceinline__ void step(const StepCtx& cx) {
        if (K > 0) cq_old = cq_n;                       // what is still queued belongs to the previous plane
        unsigned cand = 0;
        const bool any_alive_r[RW] = {
            cur[0].lo.x >= 0.0f || cur[0].lo.y >= 0.0f || cur[0].hi.x >= 0.0f || cur[0].hi.y >= 0.0f,
            cur[1].lo.x >= 0.0f || cur[1].lo.y >= 0.0f || cur[1].hi.x >= 0.0f || cur[1].hi.y >= 0.0f};
        if (__any_sync(0xffffffffu, any_alive_r[0] || any_alive_r[1])) {
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                if (!k.iso) {                           // bring the numerators to Hessian units
                    const float2 s0 = make_float2(k.s[0], k.s[0]), s1 = make_float2(k.s[1], k.s[1]), s2 = make_float2(k.s[2], k.s[2]);
                    const float2 s3 = make_float2(k.s[3], k.s[3]), s4 = make_float2(k.s[4], k.s[4]), s5 = make_float2(k.s[5], k.s[5]);
                    n.zz[r].lo = __fmul2_rn(n.zz[r].lo, s0); n.zz[r].hi = __fmul2_rn(n.zz[r].hi, s0);
                    n.zy[r].lo = __fmul2_rn(n.zy[r].lo, s1); n.zy[r].hi = __fmul2_rn(n.zy[r].hi, s1);
                    n.zx[r].lo = __fmul2_rn(n.zx[r].lo, s2); n.zx[r].hi = __fmul2_rn(n.zx[r].hi, s2);
                    n.yy[r].lo = __fmul2_rn(n.yy[r].lo, s3); n.yy[r].hi = __fmul2_rn(n.yy[r].hi, s3);
                    n.yx[r].lo = __fmul2_rn(n.yx[r].lo, s4); n.yx[r].hi = __fmul2_rn(n.yx[r].hi, s4);
                    n.xx[r].lo = __fmul2_rn(n.xx[r].lo, s5); n.xx[r].hi = __fmul2_rn(n.xx[r].hi, s5);
                }
                float fsv[4];
                bool tv[4];
#pragma unroll
                for (int hl = 0; hl < 2; ++hl) {
                    const float2 zz = hl ? n.zz[r].hi : n.zz[r].lo, zy = hl ? n.zy[r].hi : n.zy[r].lo;
                    const float2 zx = hl ? n.zx[r].hi : n.zx[r].lo, yy = hl ? n.yy[r].hi : n.yy[r].lo;
                    const float2 yx = hl ? n.yx[r].hi : n.yx[r].lo, xx = hl ? n.xx[r].hi : n.xx[r].lo;
                    // approximate frob_sq: (zz^2 + yy^2 + xx^2) + 2 (zy^2 + zx^2 + yx^2)
                    float2 d = __fmul2_rn(zz, zz);
                    d = __ffma2_rn(yy, yy, d);
                    d = __ffma2_rn(xx, xx, d);
                    float2 f = __fmul2_rn(zy, zy);
                    f = __ffma2_rn(zx, zx, f);
                    f = __ffma2_rn(yx, yx, f);
                    const float2 fs = __ffma2_rn(make_float2(2.0f, 2.0f), f, d);
                    // largest sum of two diagonal entries, minus the margin
                    const float2 ab = __fadd2_rn(zz, yy), ac = __fadd2_rn(zz, xx), bc = __fadd2_rn(yy, xx);
                    const float2 mm = make_float2(max3(ab.x, ac.x, bc.x), max3(ab.y, ac.y, bc.y));
                    const float2 t = __fadd2_rn(mm, make_float2(-zc, -zc));
                    const float2 tt = __fmul2_rn(t, t);
                    const float2 kf = __fmul2_rn(fs, make_float2(2.0e-10f, 2.0e-10f));
                    fsv[2 * hl] = fs.x; fsv[2 * hl + 1] = fs.y;
                    // tv: provably zero (given fs >= t_hi)
                    tv[2 * hl] = t.x > 0.0f && tt.x > kf.x;
                    tv[2 * hl + 1] = t.y > 0.0f && tt.y > kf.y;
                }
                float a[4] = {cur[r].lo.x, cur[r].lo.y, cur[r].hi.x, cur[r].hi.y};
                bool kill_any = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool alive = a[j] >= 0.0f;
                    const bool lo = fsv[j] < t_lo, hi = fsv[j] >= t_hi;
                    const bool kill = alive && lo;
                    if (kill) { a[j] = -1.0f; kill_any = true; }
                    if (alive && !lo && !(hi && tv[j])) cand |= 1u << (4 * r + j);
                }
                if (kill_any) {                         // dead voxels stay dead (AND of masks); ok[r] holds: alive implies it
                    const long long at = (long long)(cx.o - q.v.zg_off) * plane + row_idx[r];
                    *reinterpret_cast<float4*>(p.acc + at) = make_float4(a[0], a[1], a[2], a[3]);
                }
            }
        }
        // ---- push the candidates: one ballot per voxel slot, every lane stays active (the per-lane loop over set bits
        //      of the first version ran with 14 of 32 threads and cost 12 % of the kernel's instructions) ----
        if (p.debug == 2) cand = 0u;
        if (__ballot_sync(0xffffffffu, cand != 0u) != 0u) {
            const unsigned lt = (1u << lane) - 1u;
            int tail = cq_head + cq_n;
#pragma unroll
            for (int j = 0; j < 4 * RW; ++j) {
                const bool mine = (cand >> j) & 1u;
                const unsigned b = __ballot_sync(0xffffffffu, mine);
                if (mine) wq.cq[(tail + __popc(b & lt)) & (CQ_CAP - 1)] = (unsigned char)(((j >> 2) << 7) | (4 * lane + (j & 3)));
                tail += __popc(b);
            }
            const int added = tail - cq_head - cq_n;
            cq_n += added;
            d_cand += (unsigned long long)(lane == 0 ? added : 0);
        }
        // ---- drain (ONE call site: the exact evaluation is the bulk of the kernel's code): full warps first, then
        //      whatever still belongs to a plane that is about to leave the ring ----
        while (cq_n >= 32 || (K == 0 ? cq_n > 0 : cq_old > 0)) batch(cx, cq_n < 32 ? cq_n : 32);
#pragma unroll
        for (int j = 0; j < 5; ++j) prev_sl[j] = cx.sl[j];
        last = cx;
    }
    StepCtx last;
    __device__ void finish() {
        cq_old = 0;
        while (cq_n > 0) batch(last, cq_n < 32 ? cq_n : 32);     // K > 0: candidates of the last plane
        if (rq_n > 0) solve(rq_n);
        if (p.diag != nullptr) {
            for (int o = 16; o > 0; o >>= 1) {
                d_cand += __shfl_xor_sync(0xffffffffu, d_cand, o);
                d_kill += __shfl_xor_sync(0xffffffffu, d_kill, o);
                d_surv += __shfl_xor_sync(0xffffffffu, d_surv, o);
            }
            if (lane == 0) {
                atomicAdd(&p.diag[0], d_cand);
                atomicAdd(&p.diag[1], d_kill);
                atomicAdd(&p.diag[2], d_surv);
            }
        }
    }
};

template <int D>
struct FrangiSmem {
    Ring<D> ring;
    WarpQueues wq[NW];
};

template <int MODE, int D, int K, int L, int MINB>
__global__ void __launch_bounds__(NT, MINB)
frangi_fast_kernel(const __grid_constant__ CUtensorMap map, nb200_vol v, Consts k, int zi0, int zi1, int zchunk,
                   FrangiFastParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (p.spd[NB200_SP_SKIP] != 0.0 || p.spd[NB200_SP_UNSAFE] != 0.0) return;   // skipped sigma / exact fallback runs instead
    FrangiSmem<D>& sm = *reinterpret_cast<FrangiSmem<D>*>(smem_raw);
    const int ntx = (v.nx + TX - 1) / TX, nty = (v.ny - 4 + TYO - 1) / TYO;
    long long b = blockIdx.x;
    const int bx = (int)(b % ntx); b /= ntx;
    const int by = (int)(b % nty); b /= nty;
    Geo q;
    q.v = v;
    q.x0 = bx * TX;
    q.y0 = 2 + by * TYO;
    const int zs = zi0 + (int)b * zchunk;
    const int ze = min(zs + zchunk, zi1);
    if (zs >= ze) return;
    ring_init(sm.ring, &map, q, zs, ze);
    FrangiFastEpi<MODE, K> epi(p, k, q, sm.wq);
    consume<D, K>(sm.ring, &map, q, zs, ze, epi);
}

Consts consts_from(const float* spacing) {
    Consts k;
    for (int a = 0; a < 3; ++a) {
        k.d2[a] = spacing[2 * a + 1];
        k.r2[a] = 1.0f / spacing[2 * a + 1];
    }
    const int ax[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    k.smax = 0.0f;
    for (int c = 0; c < 6; ++c) {
        k.s[c] = (float)(1.0 / ((double)k.d2[ax[c][0]] * (double)k.d2[ax[c][1]]));
        k.smax = fmaxf(k.smax, k.s[c]);
    }
    k.iso = (k.d2[0] == k.d2[1] && k.d2[1] == k.d2[2]) ? 1 : 0;
    return k;
}

// Interior of the compute window in tiles of TX x TYO columns; Z chunks sized for several waves of 2 CTAs/SM, no
// shorter than 32 planes (same policy as nb::plan_march, other tile height)
nb::MarchPlan plan_fast(const nb200_vol& v) {
    nb::MarchPlan m;
    const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off;
    m.zi0 = max(g0, 2);
    m.zi1 = min(g1, v.nz_glob - 2);
    m.zchunk = 0;
    m.n_ctas = 0;
    if (m.zi1 <= m.zi0 || v.ny < 5 || v.nx < 12) return m;
    const long long tiles = (long long)((v.nx + TX - 1) / TX) * ((v.ny - 4 + TYO - 1) / TYO);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// TMA descriptor of the (nz_buf, ny, nx) float32 volume with a box of one staged plane tile (PITCH x GR floats)
bool make_map_fast(const float* g, const nb200_vol& v, CUtensorMap* map) {
    memset(map, 0, sizeof(*map));
    static EncodeTiledFn enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)v.nx, (cuuint64_t)v.ny, (cuuint64_t)v.nz_buf};
    const cuuint64_t strides[2] = {(cuuint64_t)v.nx * 4ull, (cuuint64_t)v.nx * (cuuint64_t)v.ny * 4ull};
    const cuuint32_t box[3] = {(cuuint32_t)PITCH, (cuuint32_t)GR, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int fast_supported(const float* g, const nb200_vol& v, int div_mode, const char* who) {
    NB_REQUIRE(div_mode == hm::DIV_FAST || div_mode == hm::DIV_POW2, NB200_ERR_UNSUPPORTED,
               "%s: needs a verified division mode (FAST or POW2)", who);
    NB_REQUIRE(v.nx % 4 == 0 && (reinterpret_cast<unsigned long long>(g) & 15ull) == 0, NB200_ERR_UNSUPPORTED,
               "%s: needs nx %% 4 == 0 and a 16-byte aligned volume", who);
    NB_REQUIRE((long long)v.nz_buf * v.ny * v.nx < (1LL << 32), NB200_ERR_UNSUPPORTED, "%s: more than 2^32 voxels", who);
    return NB200_OK;
}

// NB200_FAST_CTAS=1 pads the dynamic shared memory of the march kernels beyond half an SM so that one CTA is resident
// per SM (experiment: leave room for the FP64-bound blur of the next sigma on a second stream)
size_t padded_smem(size_t bytes) {
    static const int one = getenv("NB200_FAST_CTAS") ? atoi(getenv("NB200_FAST_CTAS")) : 2;
    return (one == 1 && bytes < 118 * 1024) ? (size_t)118 * 1024 : bytes;
}

template <class Kern>
int set_smem(Kern kernel, size_t bytes, bool& done) {
    if (done) return NB200_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        nb::set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
        return NB200_ERR_CUDA;
    }
    done = true;
    return NB200_OK;
}

constexpr int D_STATS = NB200_STATS_D;

template <int MODE>
int launch_stats(const CUtensorMap& map, const float* g, const nb200_vol& v, const Consts& k, const nb::MarchPlan& m,
                 const StatsFastParams& p, cudaStream_t st) {
    static bool done = false;
    auto kernel = stats_fast_kernel<MODE, D_STATS>;
    const size_t smem = padded_smem(sizeof(Ring<D_STATS>));
    int rc = set_smem(kernel, smem, done);
    if (rc) return rc;
    kernel<<<(unsigned)m.n_ctas, NT, smem, st>>>(map, v, k, m.zi0, m.zi1, m.zchunk, p);
    rc = nb::check_launch("hessian_stats_fast");
    if (rc) return rc;
    stats_fixup_kernel<MODE><<<2 * nb::sm_count(), 256, 0, st>>>(g, v, k, p.wl, p.hstats, m.zi0);
    return nb::check_launch("hessian_stats_fast(fixup)");
}

// ring depth / lag / slack / CTAs per SM of the K3 march: tunable through NB200_FAST_CFG (0..3) for experiments
template <int MODE, int D, int K, int L, int MINB>
int launch_frangi_cfg(const CUtensorMap& map, const nb200_vol& v, const Consts& k, const nb::MarchPlan& m,
                      const FrangiFastParams& p, cudaStream_t st) {
    static bool done = false;
    auto kernel = frangi_fast_kernel<MODE, D, K, L, MINB>;
    const size_t smem = padded_smem(sizeof(FrangiSmem<D>));
    int rc = set_smem(kernel, smem, done);
    if (rc) return rc;
    kernel<<<(unsigned)m.n_ctas, NT, smem, st>>>(map, v, k, m.zi0, m.zi1, m.zchunk, p);
    return nb::check_launch("frangi_fast");
}
template <int MODE>
int launch_frangi(const CUtensorMap& map, const nb200_vol& v, const Consts& k, const nb::MarchPlan& m,
                  const FrangiFastParams& p, cudaStream_t st) {
    static const int cfg = getenv("NB200_FAST_CFG") ? atoi(getenv("NB200_FAST_CFG")) : 0;
    switch (cfg) {
        case 1: return launch_frangi_cfg<MODE, 8, 0, 0, 2>(map, v, k, m, p, st);
        default: return launch_frangi_cfg<MODE, 8, 1, 0, 2>(map, v, k, m, p, st);
    }
}

}  // namespace hf
}  // namespace

extern "C" {

size_t nb200_hessian_fast_workspace_bytes(void) { return sizeof(unsigned) * (hf::WL_HDR + 8ull * hf::WL_CAP); }

int nb200_hessian_stats_fast(const float* gauss, const nb200_vol* vol, const float* spacing, int div_mode,
                             int sz, int sy, int sx, float* frob_samples, long long* hstats, void* workspace,
                             void* stream) {
    NB_REQUIRE(gauss && vol && spacing && hstats && workspace && sz > 0 && sy > 0 && sx > 0, NB200_ERR_ARG,
               "nb200_hessian_stats_fast: bad argument");
    const nb200_vol v = *vol;
    int rc = nb::check_vol(v, "nb200_hessian_stats_fast");
    if (rc) return rc;
    rc = hf::fast_supported(gauss, v, div_mode, "nb200_hessian_stats_fast");
    if (rc) return rc;
    if (v.zc0 == v.zc1) return NB200_OK;
    cudaStream_t st = nb::as_stream(stream);
    hf::StatsFastParams p;
    p.sz = sz; p.sy = sy; p.sx = sx;
    const int g0 = v.zc0 + v.zg_off;
    p.g_first = ((g0 + sz - 1) / sz) * sz;
    p.ly_n = (v.ny + sy - 1) / sy;
    p.lx_n = (v.nx + sx - 1) / sx;
    p.frob_samples = frob_samples;
    p.hstats = hstats;
    p.wl = reinterpret_cast<unsigned*>(workspace);
    hf::wl_reset_kernel<<<1, 32, 0, st>>>(p.wl);
    rc = nb::check_launch("hessian_stats_fast(reset)");
    if (rc) return rc;
    const nb::MarchPlan m = hf::plan_fast(v);
    if (m.n_ctas > 0) {
        CUtensorMap map;
        NB_REQUIRE(hf::make_map_fast(gauss, v, &map), NB200_ERR_UNSUPPORTED,
                   "nb200_hessian_stats_fast: no TMA descriptor for this volume");
        const hf::Consts k = hf::consts_from(spacing);
        rc = div_mode == hm::DIV_POW2 ? hf::launch_stats<2>(map, gauss, v, k, m, p, st)
                                      : hf::launch_stats<1>(map, gauss, v, k, m, p, st);
        if (rc) return rc;
    }
    // border shell: exact per-voxel kernel (IEEE division, one-sided differences)
    return nb::launch_shell_stats(gauss, v, spacing, sz, sy, sx, frob_samples, hstats, nullptr, nullptr, -1, st);
}

int nb200_frangi_fast(const float* gauss, float* acc, const nb200_vol* vol, const float* spacing, int div_mode,
                      float alpha_sq, float beta_sq, const double* sp, unsigned long long* diag, void* stream) {
    NB_REQUIRE(gauss && acc && vol && spacing && sp, NB200_ERR_ARG, "nb200_frangi_fast: null argument");
    const nb200_vol v = *vol;
    int rc = nb::check_vol(v, "nb200_frangi_fast");
    if (rc) return rc;
    rc = hf::fast_supported(gauss, v, div_mode, "nb200_frangi_fast");
    if (rc) return rc;
    NB_REQUIRE((reinterpret_cast<unsigned long long>(acc) & 15ull) == 0, NB200_ERR_UNSUPPORTED,
               "nb200_frangi_fast: acc must be 16-byte aligned");
    if (v.zc0 == v.zc1) return NB200_OK;
    cudaStream_t st = nb::as_stream(stream);
    const nb::MarchPlan m = hf::plan_fast(v);
    if (m.n_ctas > 0) {
        CUtensorMap map;
        NB_REQUIRE(hf::make_map_fast(gauss, v, &map), NB200_ERR_UNSUPPORTED,
                   "nb200_frangi_fast: no TMA descriptor for this volume");
        const hf::Consts k = hf::consts_from(spacing);
        hf::FrangiFastParams p;
        p.acc = acc;
        p.alpha_sq = alpha_sq;
        p.beta_sq = beta_sq;
        p.spd = sp;
        p.diag = diag;
        {
            static const int dbg = getenv("NB200_K3_DEBUG") ? atoi(getenv("NB200_K3_DEBUG")) : 0;
            p.debug = dbg;
        }
        rc = div_mode == hm::DIV_POW2 ? hf::launch_frangi<2>(map, v, k, m, p, st) : hf::launch_frangi<1>(map, v, k, m, p, st);
        if (rc) return rc;
    }
    // border shell after the interior, exact; it returns at once for a skipped sigma.  With sp[UNSAFE] the fallback
    // (nb200_frangi_sparse_gated) covers the shell itself, so the shell kernel must not run twice: it is idempotent
    // (max / AND of the same values), so running it in both cases is harmless.
    return nb::launch_shell_frangi(gauss, acc, v, spacing, alpha_sq, beta_sq, sp, st);
}

}  // extern "C"
