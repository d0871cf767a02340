// Z-marching finite-difference Hessian of the frame INTERIOR (shared by K2 = hessian_stats and
// K3 = frangi_accumulate).  numpy.gradient(numpy.gradient(g)) semantics, SURVEY.md A.2 /
// nellie/segmentation/filtering.py:446-551.
//
// Split of the frame.  A voxel is "interior" when every second derivative uses central differences
// only: 2 <= z <= nz-3, 2 <= y <= ny-3 and its aligned group of four x values lies in [4, xhi),
// xhi = 4*floor((nx-2)/4).  Interior voxels (98.5 % of a 1024^3 frame) run through the marching kernel
// below, which has no border logic at all; the remaining shell is evaluated by the generic per-voxel
// kernels in frangi.cu (hessian.cuh carries the one-sided rules).  Results are identical either way.
//
// March.  A CTA (256 threads) owns 128 x 14 output columns and walks along Z.  Blurred planes arrive by
// TMA (cp.async.bulk.tensor.3d, one elected thread, mbarrier completion) into a 4-slot shared-memory
// ring: box 136 x 18 floats = the tile plus a halo of 4 columns / 2 rows (out-of-frame parts are
// zero-filled and only ever feed voxels of the border shell).  Every FIRST derivative (one float32
// subtraction + one correctly rounded division) is evaluated once and shared:
//     gz(z+1), gy(z)  -> shared memory (16 rows x 136 columns; rows/columns +-1 around the outputs)
// while each thread keeps the first derivatives of its own two adjacent rows in registers across planes
// (gz(z-1), gz(z), gz(z+1), gy(z)), so d2/dz2 needs no shared-memory read at all and the Y-direction
// second derivatives read one neighbour row each.  X-direction neighbours come from warp shuffles (the
// two halo columns of a tile from shared memory).  A thread produces 2 rows x 4 X-consecutive voxels per
// plane; the arithmetic runs on Blackwell's packed FFMA2/FMUL2.
//
// Division by the grid spacing: numpy divides by fl32(2h).  DIV_POW2 multiplies by the exact reciprocal
// when the divisor is a power of two; DIV_FAST uses q0 = n*r, q = fma(fma(-q0,d,n), r, q0) with
// r = RN(1/d) — enabled per divisor only after an exhaustive on-device comparison against IEEE division
// (nb200_divisor_mode); DIV_IEEE is `/`.
#pragma once
#include <cuda.h>

#include "devmath.cuh"

namespace hm {

constexpr int TX = 128;            // output columns per CTA
constexpr int TYO = 14;            // output rows per CTA
constexpr int DR = 16;             // first-derivative rows: d <-> y = y0 - 1 + d
constexpr int GR = 18;             // staged rows of g:      r <-> y = y0 - 2 + r   (derivative row d <-> g row d+1)
constexpr int PITCH = TX + 8;      // staged columns: c <-> x = x0 - 4 + c
constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int NG = 4;              // ring depth of g planes
constexpr int G_SLOT = 2464;       // floats per ring slot (GR*PITCH = 2448 rounded up to a multiple of 128 bytes)
constexpr int D_PLANE = DR * PITCH;
constexpr unsigned G_BOX_BYTES = GR * PITCH * 4;

enum DivMode { DIV_IEEE = 0, DIV_FAST = 1, DIV_POW2 = 2 };

struct Divs {
    float d2[3], r2[3];   // fl32(2h) and RN(1/fl32(2h)) for Z, Y, X
};

template <int MODE>
__device__ __forceinline__ float divc(float n, float d, float r) {
    if (MODE == DIV_POW2) return n * r;
    if (MODE == DIV_FAST) {
        const float q0 = n * r;
        const float e = fmaf(-q0, d, n);
        return fmaf(e, r, q0);
    }
    return n / d;
}

// ---- packed float32x2 arithmetic (FFMA2 / FMUL2: two IEEE round-to-nearest results per instruction) ----
struct DivK {
    float2 rr;   // (r, r),  r = RN(1/d)
    float2 nd;   // (-d, -d)
};
__device__ __forceinline__ DivK make_divk(float d, float r) {
    DivK k;
    k.rr = make_float2(r, r);
    k.nd = make_float2(-d, -d);
    return k;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {       // a - b, rounded exactly like FADD
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {       // a + b, rounded exactly like FADD
    return __ffma2_rn(a, make_float2(1.0f, 1.0f), b);
}
template <int MODE>
__device__ __forceinline__ float2 div2(float2 n, const DivK& k) {
    if (MODE == DIV_POW2) return __fmul2_rn(n, k.rr);
    if (MODE == DIV_FAST) {
        const float2 q0 = __fmul2_rn(n, k.rr);
        const float2 e = __ffma2_rn(q0, k.nd, n);
        return __ffma2_rn(e, k.rr, q0);
    }
    return make_float2(n.x / -k.nd.x, n.y / -k.nd.x);
}
template <int MODE>
__device__ __forceinline__ float4 div4(const float4& n, const DivK& k) {
    const float2 lo = div2<MODE>(make_float2(n.x, n.y), k), hi = div2<MODE>(make_float2(n.z, n.w), k);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// (a - b) / d for 4 vertically aligned elements
template <int MODE>
__device__ __forceinline__ float4 diff_div4(const float4& a, const float4& b, const DivK& k) {
    const float2 lo = div2<MODE>(sub2(make_float2(a.x, a.y), make_float2(b.x, b.y)), k);
    const float2 hi = div2<MODE>(sub2(make_float2(a.z, a.w), make_float2(b.z, b.w)), k);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

struct Smem {
    float g[NG][G_SLOT];          // TMA destinations, 128-byte aligned
    float gz[3][D_PLANE];
    float gy[D_PLANE];
    float gxh[DR][2];             // d/dx of g(z) at the two halo columns x0-1 and x0+TX
    unsigned long long bar[NG];   // "plane landed" mbarriers
};

// ---- TMA / mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

struct Geo {
    nb200_vol v;
    int x0, y0;
    bool tma;                 // planes arrive by TMA (nx % 4 == 0, 16-byte aligned base); else plain loads
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// six second derivatives of 4 X-consecutive voxels; component order of the reference's matrix:
// zz = d0d0 ("hxx"), zy = d1d0 ("hxy"), zx = d2d0 ("hxz"), yy = d1d1, yx = d2d1 ("hyz"), xx = d2d2 ("hzz")
struct Hess4 {
    float4 zz, zy, zx, yy, yx, xx;
};

// central difference along X of 4 consecutive elements held by this lane; neighbours by shuffle, the tile's
// two halo columns (lane 0 / lane 31) from shared memory.  Must be called by the whole warp.
template <int MODE>
__device__ __forceinline__ float4 ddx4(const float4& m, const float* halo_l, const float* halo_r, int lane,
                                       const DivK& kx) {
    float left = __shfl_up_sync(0xffffffffu, m.w, 1);
    float right = __shfl_down_sync(0xffffffffu, m.x, 1);
    if (lane == 0) left = *halo_l;
    if (lane == 31) right = *halo_r;
    return div4<MODE>(make_float4(m.y - left, m.z - m.x, m.w - m.y, right - m.z), kx);
}

// stage global plane `zg` into its ring slot (slot = relative plane index & 3)
__device__ __forceinline__ void load_plane(Smem& s, const CUtensorMap* map, const float* __restrict__ g, const Geo& q,
                                           int zg, int rel) {
    const int slot = rel & (NG - 1);
    const int zb = zg - q.v.zg_off;
    if (q.tma) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s.bar[slot], G_BOX_BYTES);
            tma_load_3d(&s.g[slot][0], map, &s.bar[slot], q.x0 - 4, q.y0 - 2, zb);
        }
    } else {
        const float* base = g + (long long)zb * q.v.ny * q.v.nx;
        for (int i = threadIdx.x; i < GR * PITCH; i += NT) {
            const int r = i / PITCH, c = i - r * PITCH;
            const int y = q.y0 - 2 + r, x = q.x0 - 4 + c;
            float val = 0.0f;
            if (y >= 0 && y < q.v.ny && x >= 0 && x < q.v.nx) val = __ldg(base + (long long)y * q.v.nx + x);
            s.g[slot][i] = val;
        }
    }
}
__device__ __forceinline__ void wait_plane(Smem& s, const Geo& q, int rel) {
    if (q.tma) mbar_wait(&s.bar[rel & (NG - 1)], (unsigned)(rel >> 2) & 1u);
}

// per-thread constants of the march
struct Lane {
    int warp, lane;
    int od0;       // float offset of (derivative row d0 = 2*warp, column 4 + 4*lane) inside a derivative plane
    int og0;       // float offset of the same voxel group inside a g slot (g row d0 + 1)
};

// gz at plane z from g(z+1) [hi] and g(z-1) [lo]: this thread's two rows -> registers and the shared plane `dst`.
// lo0 / lo1 return the two rows of `lo` (the centre plane of the caller) for reuse.
template <int MODE>
__device__ __forceinline__ void produce_gz(const float* hi, const float* lo, float* dst, const Lane& t, const DivK& kz,
                                           float4& r0, float4& r1, float4& lo0, float4& lo1) {
    lo0 = ld4(lo + t.og0);
    lo1 = ld4(lo + t.og0 + PITCH);
    r0 = diff_div4<MODE>(ld4(hi + t.og0), lo0, kz);
    r1 = diff_div4<MODE>(ld4(hi + t.og0 + PITCH), lo1, kz);
    st4(dst + t.od0, r0);
    st4(dst + t.od0 + PITCH, r1);
}

// The two halo columns x0-1 (warp 0) and x0+TX (warp NW-1) of the first-derivative planes.  These two warps
// output one row less than the others, so the extra work is free.  lanes 0-15: gz, lanes 16-31: gy; then gx.
template <int MODE>
__device__ __forceinline__ void produce_halo(Smem& s, const float* g_hi, const float* g_lo, const float* g_c,
                                             float* gz_dst, bool want_gy_gx, const Lane& t, const Divs& dv) {
    if (t.warp != 0 && t.warp != NW - 1) return;
    const int c = t.warp == 0 ? 3 : TX + 4;
    const int row = t.lane & 15;
    const bool is_y = t.lane >= 16;
    if (!is_y || want_gy_gx) {
        const float* pa = is_y ? g_c + (row + 2) * PITCH + c : g_hi + (row + 1) * PITCH + c;
        const float* pb = is_y ? g_c + row * PITCH + c : g_lo + (row + 1) * PITCH + c;
        const float d = is_y ? dv.d2[1] : dv.d2[0], r = is_y ? dv.r2[1] : dv.r2[0];
        const float val = divc<MODE>(*pa - *pb, d, r);
        (is_y ? s.gy : gz_dst)[row * PITCH + c] = val;
    }
    if (want_gy_gx && !is_y) {
        const float* rowp = g_c + (row + 1) * PITCH + c;
        s.gxh[row][t.warp == 0 ? 0 : 1] = divc<MODE>(rowp[1] - rowp[-1], dv.d2[2], dv.r2[2]);
    }
}

// The march over global planes [zs, ze) (all interior: 2 <= zs, ze <= nz_glob - 2).  Epi provides:
//   void plane(int zg);                                 // once per output plane (uniform)
//   void prefetch(int i, bool inb, long long idx);      // issue the loads the epilogue needs for row i of this plane
//   void voxels4(int i, bool valid, long long idx, const Hess4&);   // called by ALL lanes of the warp
//   void cta_sync_point();                              // called by every thread right after the plane's last barrier
//   void center(const float4&, const float4&);          // the thread's own blurred values of this plane (range tracking)
// idx = linear index (buffer coordinates) of the first voxel of the lane's group of four.
template <int MODE, class Epi>
__device__ __forceinline__ void march(Smem& s, const CUtensorMap* map, const float* __restrict__ g, const Geo& q,
                                      const Divs& dv, int zs, int ze, Epi& epi) {
    Lane t;
    t.warp = threadIdx.x >> 5;
    t.lane = threadIdx.x & 31;
    const int d0 = 2 * t.warp;
    t.od0 = d0 * PITCH + 4 + 4 * t.lane;
    t.og0 = t.od0 + PITCH;
    const DivK kz = make_divk(dv.d2[0], dv.r2[0]), ky = make_divk(dv.d2[1], dv.r2[1]), kx = make_divk(dv.d2[2], dv.r2[2]);
    const int x = q.x0 + 4 * t.lane;
    const int xhi = 4 * ((q.v.nx - 2) / 4);
    const bool col_inb = x + 3 < q.v.nx;
    const bool col_valid = x >= 4 && x < xhi;
    bool row_out[2], row_inb[2];          // warp-uniform
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int d = d0 + i, y = q.y0 - 1 + d;
        row_inb[i] = d >= 1 && d <= TYO && y < q.v.ny;
        row_out[i] = row_inb[i] && y >= 2 && y <= q.v.ny - 3;
    }
    const long long plane = (long long)q.v.ny * q.v.nx;
    long long idx0 = (long long)(zs - q.v.zg_off) * plane + (long long)(q.y0 - 1 + d0) * q.v.nx + x;

    // ---- prologue: planes zs-2 .. zs+1 (relative index 0..3), then gz(zs-1) and gz(zs) ----
    if (q.tma && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NG; ++k) mbar_init(&s.bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NG; ++k) load_plane(s, map, g, q, zs - 2 + k, k);
#pragma unroll
    for (int k = 0; k < NG; ++k) wait_plane(s, q, k);
    if (!q.tma) __syncthreads();
    float4 A0, A1, B0, B1, C0, C1, c0, c1;   // gz(z-1), gz(z), gz(z+1) and g(z) of this thread's two rows
    produce_gz<MODE>(s.g[2], s.g[0], s.gz[0], t, kz, A0, A1, c0, c1);
    produce_halo<MODE>(s, s.g[2], s.g[0], s.g[1], s.gz[0], false, t, dv);
    produce_gz<MODE>(s.g[3], s.g[1], s.gz[1], t, kz, B0, B1, c0, c1);
    produce_halo<MODE>(s, s.g[3], s.g[1], s.g[2], s.gz[1], false, t, dv);
    __syncthreads();                       // plane zs-2 is dead: its slot takes plane zs+2
    load_plane(s, map, g, q, zs + 2, NG);

    int rel = 2;             // relative index of the output plane tz; g(tz) sits in ring slot rel & 3
    int zc = 1;              // gz(tz) sits in gz slot zc, gz(tz+1) goes to slot zc+1 (mod 3)
    for (int tz = zs; tz < ze; ++tz, ++rel) {
        // plane tz+3 replaces plane tz-1 (dead since the barrier that ended the previous iteration)
        if (tz + 3 <= ze + 1) load_plane(s, map, g, q, tz + 3, rel + 3);
        wait_plane(s, q, rel + 2);
        if (!q.tma) __syncthreads();
        const float* g_c = s.g[rel & 3];
        const float* g_hi = s.g[(rel + 2) & 3];
        const int zn = zc == 2 ? 0 : zc + 1;
        // ---- phase 1: first derivatives gz(tz+1), gy(tz) (+ halo columns, gx halo) ----
        epi.plane(tz);
#pragma unroll
        for (int i = 0; i < 2; ++i) epi.prefetch(i, row_inb[i] && col_inb, idx0 + (long long)i * q.v.nx);
        produce_gz<MODE>(g_hi, g_c, s.gz[zn], t, kz, C0, C1, c0, c1);
        epi.center(c0, c1);                // the blurred values g(tz) of this thread's two rows
        float4 Y0, Y1;
        {
            const float4 gm = ld4(g_c + t.og0 - PITCH), gp = ld4(g_c + t.og0 + 2 * PITCH);
            Y0 = diff_div4<MODE>(c1, gm, ky);
            Y1 = diff_div4<MODE>(gp, c0, ky);
            st4(s.gy + t.od0, Y0);
            st4(s.gy + t.od0 + PITCH, Y1);
        }
        produce_halo<MODE>(s, g_hi, g_c, g_c, s.gz[zn], true, t, dv);
        __syncthreads();
        // ---- phase 2: second derivatives of this thread's rows ----
        const float* GZ = s.gz[zc];
        if (row_out[0]) {
            Hess4 h;
            h.zz = diff_div4<MODE>(C0, A0, kz);
            h.zy = diff_div4<MODE>(B1, ld4(GZ + t.od0 - PITCH), ky);
            h.yy = diff_div4<MODE>(Y1, ld4(s.gy + t.od0 - PITCH), ky);
            const float* zr = GZ + d0 * PITCH;
            const float* yr = s.gy + d0 * PITCH;
            const float* gr = g_c + (d0 + 1) * PITCH;
            h.zx = ddx4<MODE>(B0, zr + 3, zr + TX + 4, t.lane, kx);
            h.yx = ddx4<MODE>(Y0, yr + 3, yr + TX + 4, t.lane, kx);
            const float4 gx = ddx4<MODE>(c0, gr + 3, gr + TX + 4, t.lane, kx);
            h.xx = ddx4<MODE>(gx, &s.gxh[d0][0], &s.gxh[d0][1], t.lane, kx);
            epi.voxels4(0, col_valid, idx0, h);
        }
        if (row_out[1]) {
            Hess4 h;
            h.zz = diff_div4<MODE>(C1, A1, kz);
            h.zy = diff_div4<MODE>(ld4(GZ + t.od0 + 2 * PITCH), B0, ky);
            h.yy = diff_div4<MODE>(ld4(s.gy + t.od0 + 2 * PITCH), Y0, ky);
            const float* zr = GZ + (d0 + 1) * PITCH;
            const float* yr = s.gy + (d0 + 1) * PITCH;
            const float* gr = g_c + (d0 + 2) * PITCH;
            h.zx = ddx4<MODE>(B1, zr + 3, zr + TX + 4, t.lane, kx);
            h.yx = ddx4<MODE>(Y1, yr + 3, yr + TX + 4, t.lane, kx);
            const float4 gx = ddx4<MODE>(c1, gr + 3, gr + TX + 4, t.lane, kx);
            h.xx = ddx4<MODE>(gx, &s.gxh[d0 + 1][0], &s.gxh[d0 + 1][1], t.lane, kx);
            epi.voxels4(1, col_valid, idx0 + q.v.nx, h);
        }
        A0 = B0; A1 = B1; B0 = C0; B1 = C1;
        zc = zn;
        idx0 += plane;
        __syncthreads();                   // gy, the gz slot of gz(tz-1) and g(tz-1)'s ring slot may be overwritten
        epi.cta_sync_point();              // CTA-uniform hook (K3: cooperative eigen-solves)
    }
}

}  // namespace hm
