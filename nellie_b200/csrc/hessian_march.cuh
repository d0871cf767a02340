// Z-marching finite-difference Hessian (shared by K2 = hessian_stats and K3 = frangi_accumulate).
//
// A CTA owns a TX x TY column of the frame and marches along Z.  Shared memory holds a ring of
// blurred planes g (cp.async, prefetched one plane ahead) and the FIRST-derivative planes
//     gz(z-1), gz(z), gz(z+1)   (halo 1 in Y and X)      gy(z) (halo 1)      gx(z) (halo 1 in X)
// so every first derivative (one float32 subtraction + one correctly rounded division, numpy.gradient
// semantics) is evaluated ONCE and shared by the up to four second derivatives that
// use it.  Each thread produces 4 X-consecutive voxels per row with 128-bit shared-memory accesses;
// all per-thread shared-memory offsets are loop invariants, the ring slots rotate as base pointers.
//
// Frame borders (one-sided differences, divisor h instead of 2h): Z is uniform per plane, Y is uniform
// per row, X touches at most two voxels of a row, which are patched after the branch-free interior
// formula.  Values staged from outside the frame are clamped duplicates that no result ever uses.
//
// Division by the grid spacing: numpy divides by fl32(h) / fl32(2h).  DIV_POW2 multiplies by the exact
// reciprocal when the divisor is a power of two; DIV_FAST uses q0 = n*r, q = fma(fma(-q0,d,n), r, q0)
// with r = RN(1/d) — enabled per divisor only after an exhaustive on-device comparison against IEEE
// division over all numerators in the safe exponent range (nb200_divisor_mode); DIV_IEEE is `/`.
#pragma once
#include "devmath.cuh"

namespace hm {

constexpr int TX = 128;           // outputs along X per CTA
constexpr int TY = 16;            // outputs along Y per CTA
constexpr int NT = 256;           // threads
constexpr int NW = NT / 32;       // warps
constexpr int PITCH = TX + 8;     // floats per shared row; column 4 <-> x0 (16-byte aligned)
constexpr int GROWS = TY + 4;     // g rows:  row r <-> y = y0 - 2 + r
constexpr int DROWS = TY + 2;     // gz/gy rows: row r <-> y = y0 - 1 + r
constexpr int NG = 4;             // g ring depth (z, z+1, z+2 and the prefetch of z+3)

enum DivMode { DIV_IEEE = 0, DIV_FAST = 1, DIV_POW2 = 2 };

struct AxisDiv {
    float d1, r1;   // fl32(h),  RN(1/fl32(h))     (one-sided edges)
    float d2, r2;   // fl32(2h), RN(1/fl32(2h))    (interior)
};
struct Divs {
    AxisDiv a[3];   // Z, Y, X
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

// ---- packed float32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest results per
// issued instruction; per-element results are identical to the scalar instructions) -------------------------
struct DivK {            // one divisor, broadcast into register pairs
    float2 rr;           // (r, r),  r = RN(1/d)
    float2 nd;           // (-d, -d)
    float d, r;
};
__device__ __forceinline__ DivK make_divk(float d, float r) {
    DivK k;
    k.rr = make_float2(r, r);
    k.nd = make_float2(-d, -d);
    k.d = d;
    k.r = r;
    return k;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {       // a - b, exactly as FADD would round it
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
}
template <int MODE>
__device__ __forceinline__ float2 div2(float2 n, const DivK& k) {
    if (MODE == DIV_POW2) return __fmul2_rn(n, k.rr);
    if (MODE == DIV_FAST) {
        const float2 q0 = __fmul2_rn(n, k.rr);
        const float2 e = __ffma2_rn(q0, k.nd, n);
        return __ffma2_rn(e, k.rr, q0);
    }
    return make_float2(n.x / k.d, n.y / k.d);
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
    float g[NG][GROWS][PITCH];
    float gz[3][DROWS][PITCH];
    float gy[DROWS][PITCH];
    float gx[TY][PITCH];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Geo {
    nb200_vol v;
    int x0, y0;
    long long plane;
    bool vec_ok;      // 16-byte cp.async allowed (full tile inside the frame in X, nx % 4 == 0, aligned base)
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// per-thread loop invariants
struct Lane {
    int warp, lane;
    int col;          // 4 + 4*lane : first column of this thread's vector group
    int x;            // global x of that column
    int kfirst;       // 0 if the group holds x == 0, else -1
    int klast;        // index (0..3) of x == nx-1 inside the group, else -1
};

// per-CTA division constants in registers
struct DivSet {
    DivK z2, y2, x2;      // interior divisors fl32(2h)
    AxisDiv az, ay, ax;   // both divisors per axis (borders use d1)
};

// central difference along X of 4 consecutive elements of a shared row (numpy.gradient, axis X).
// EDGE: the group may contain x == 0 or x == nx-1, patched with the one-sided rule.
template <int MODE, bool EDGE>
__device__ __forceinline__ float4 ddx4(const float* row, const Lane& t, const DivSet& ds) {
    const float4 m = ld4(row + t.col);
    const float left = row[t.col - 1], right = row[t.col + 4];
    // numerators land in fresh register pairs, the division runs packed
    float4 o = div4<MODE>(make_float4(m.y - left, m.z - m.x, m.w - m.y, right - m.z), ds.x2);
    if (EDGE) {
        if ((t.kfirst & t.klast) != -1) {          // a frame border lies inside this group (rare)
            const float vals[6] = {left, m.x, m.y, m.z, m.w, right};
            float res[4] = {o.x, o.y, o.z, o.w};
            if (t.kfirst == 0) res[0] = divc<MODE>(m.y - m.x, ds.ax.d1, ds.ax.r1);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (t.klast == k && !(k == 0 && t.kfirst == 0))
                    res[k] = divc<MODE>(vals[k + 1] - vals[k], ds.ax.d1, ds.ax.r1);
            o = make_float4(res[0], res[1], res[2], res[3]);
        }
    }
    return o;
}

// issue the asynchronous load of global plane `zg` into ring slot zg & 3
__device__ __forceinline__ void load_g_plane(Smem& s, const float* __restrict__ g, const Geo& q, const Lane& t,
                                             int zg) {
    if (zg < 0 || zg > q.v.nz_glob - 1) return;    // outside the frame: never referenced
    const float* base = g + (long long)(zg - q.v.zg_off) * q.plane;
    float(*dst)[PITCH] = s.g[zg & 3];
#pragma unroll
    for (int i = 0; i < (GROWS + NW - 1) / NW; ++i) {
        const int r = t.warp + i * NW;
        if (r < GROWS) {
            const int y = min(max(q.y0 - 2 + r, 0), q.v.ny - 1);
            const float* row = base + (long long)y * q.v.nx;
            if (q.vec_ok) {
                cp_async16(&dst[r][t.col], row + t.x);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) cp_async4(&dst[r][t.col + k], row + min(t.x + k, q.v.nx - 1));
            }
            if (t.lane < 4) {   // halo columns x0-2, x0-1, x0+TX, x0+TX+1
                const int c = t.lane < 2 ? 2 + t.lane : TX + 2 + t.lane;
                cp_async4(&dst[r][c], row + min(max(q.x0 - 4 + c, 0), q.v.nx - 1));
            }
        }
    }
    cp_async_commit();
}

// The two halo columns (x0-1 and x0+TX) of a first-derivative plane: one warp, lane = row, so the
// vector passes above stay free of divergent lane tests.
template <int MODE, class F>
__device__ __forceinline__ void halo_columns(const Lane& t, int nrows, F elem) {
    if (t.warp == NW - 1 && t.lane < nrows) {
        elem(t.lane, 3);
        elem(t.lane, TX + 4);
    }
}

// gz(zz) = d/dz of the blurred volume at plane zz, rows y0-1..y0+TY, cols x0-1..x0+TX
template <int MODE>
__device__ __forceinline__ void produce_gz(Smem& s, const Geo& q, const Lane& t, const DivSet& ds, int zz, int slot) {
    const int n = q.v.nz_glob;
    if (zz < 0 || zz > n - 1) return;
    int hi = zz + 1, lo = zz - 1;
    DivK k = ds.z2;
    if (zz == 0 || zz == n - 1) {
        if (zz == 0) lo = 0;
        if (zz == n - 1) hi = n - 1;
        k = make_divk(ds.az.d1, ds.az.r1);
    }
    const float* A = &s.g[hi & 3][1][0];
    const float* B = &s.g[lo & 3][1][0];
    float* O = &s.gz[slot][0][0];
#pragma unroll
    for (int i = 0; i < (DROWS + NW - 1) / NW; ++i) {
        const int rr = t.warp + i * NW;
        if (rr < DROWS) {
            const int o = rr * PITCH + t.col;
            st4(O + o, diff_div4<MODE>(ld4(A + o), ld4(B + o), k));
        }
    }
    halo_columns<MODE>(t, DROWS, [&](int rr, int c) {
        const int o = rr * PITCH + c;
        O[o] = divc<MODE>(A[o] - B[o], k.d, k.r);
    });
}

// gy(z): rows y0-1 .. y0+TY (row-uniform border handling)
template <int MODE>
__device__ __forceinline__ void produce_gy(Smem& s, const Geo& q, const Lane& t, const DivSet& ds, int z) {
    const float* G = &s.g[z & 3][0][0];
    auto row_rule = [&](int rr, int& rhi, int& rlo, DivK& k) -> bool {
        const int y = q.y0 - 1 + rr;
        if (y < 0 || y > q.v.ny - 1) return false;       // outside the frame: never referenced
        rhi = rr + 2;
        rlo = rr;
        k = ds.y2;
        if (y == 0 || y == q.v.ny - 1) {
            if (y == 0) rlo = rr + 1;
            if (y == q.v.ny - 1) rhi = rr + 1;
            k = make_divk(ds.ay.d1, ds.ay.r1);
        }
        return true;
    };
#pragma unroll
    for (int i = 0; i < (DROWS + NW - 1) / NW; ++i) {
        const int rr = t.warp + i * NW;
        if (rr < DROWS) {
            int rhi, rlo;
            DivK k;
            if (row_rule(rr, rhi, rlo, k))
                st4(&s.gy[rr][t.col], diff_div4<MODE>(ld4(G + rhi * PITCH + t.col), ld4(G + rlo * PITCH + t.col), k));
        }
    }
    halo_columns<MODE>(t, DROWS, [&](int rr, int c) {
        int rhi, rlo;
        DivK k;
        if (row_rule(rr, rhi, rlo, k)) s.gy[rr][c] = divc<MODE>(G[rhi * PITCH + c] - G[rlo * PITCH + c], k.d, k.r);
    });
}

// gx(z): rows y0 .. y0+TY-1, cols x0-1 .. x0+TX
template <int MODE, bool EDGE>
__device__ __forceinline__ void produce_gx(Smem& s, const Geo& q, const Lane& t, const DivSet& ds, int z) {
    const float* G = &s.g[z & 3][2][0];
#pragma unroll
    for (int i = 0; i < TY / NW; ++i) {
        const int rr = t.warp + i * NW;
        st4(&s.gx[rr][t.col], ddx4<MODE, EDGE>(G + rr * PITCH, t, ds));
    }
    halo_columns<MODE>(t, TY, [&](int rr, int h) {
        const float* row = G + rr * PITCH;
        const int x = q.x0 - 4 + h;
        if (!EDGE || (x >= 0 && x <= q.v.nx - 1)) {
            float hi = row[h + 1], lo = row[h - 1], d = ds.ax.d2, r = ds.ax.r2;
            if (EDGE) {
                if (x == 0) { lo = row[h]; d = ds.ax.d1; r = ds.ax.r1; }
                if (x == q.v.nx - 1) { hi = row[h]; d = ds.ax.d1; r = ds.ax.r1; }
            }
            s.gx[rr][h] = divc<MODE>(hi - lo, d, r);
        }
    });
}

// six second derivatives of 4 X-consecutive voxels; component order of the reference's matrix:
// zz = d0d0 ("hxx"), zy = d1d0 ("hxy"), zx = d2d0 ("hxz"), yy = d1d1, yx = d2d1 ("hyz"), xx = d2d2 ("hzz")
struct Hess4 {
    float4 zz, zy, zx, yy, yx, xx;
};

// The march.  Epi provides:
//   void plane(int zg);                                            // once per output plane (uniform)
//   void preload(int row, int zb, int y, int x, int nvalid);       // issue loads the epilogue needs for that plane
//   bool skip4(int row, int zb, int y, int x, int nvalid);         // true: this group needs no Hessian
//   void voxels4(int row, bool active, int zb, int y, int x, int nvalid, const Hess4&);   // called by ALL lanes
// EDGE = the tile touches the frame border in X or Y or is partial (one-sided rules, bounds checks).
template <int MODE, bool EDGE, class Epi>
__device__ __forceinline__ void march(Smem& s, const float* __restrict__ g, const Geo& q, const Divs& dv,
                                      int zs, int ze, Epi& epi) {
    // zs, ze: GLOBAL plane range [zs, ze) this CTA computes
    Lane t;
    t.warp = threadIdx.x >> 5;
    t.lane = threadIdx.x & 31;
    t.col = 4 + 4 * t.lane;
    t.x = q.x0 + 4 * t.lane;
    t.kfirst = t.x == 0 ? 0 : -1;
    const int kl = q.v.nx - 1 - t.x;
    t.klast = (kl >= 0 && kl <= 3) ? kl : -1;
    DivSet ds;
    ds.az = dv.a[0]; ds.ay = dv.a[1]; ds.ax = dv.a[2];
    ds.z2 = make_divk(dv.a[0].d2, dv.a[0].r2);
    ds.y2 = make_divk(dv.a[1].d2, dv.a[1].r2);
    ds.x2 = make_divk(dv.a[2].d2, dv.a[2].r2);
    const int n = q.v.nz_glob;
    // loop-invariant row data of the TY/NW output rows of this thread
    int y_[TY / NW], nvalid_[TY / NW], off_[TY / NW], up_[TY / NW], dn_[TY / NW];
    bool yedge_[TY / NW];
#pragma unroll
    for (int i = 0; i < TY / NW; ++i) {
        const int ty = t.warp + i * NW;
        y_[i] = q.y0 + ty;
        nvalid_[i] = EDGE ? ((y_[i] < q.v.ny) ? max(0, min(4, q.v.nx - t.x)) : 0) : 4;
        off_[i] = (ty + 1) * PITCH + t.col;
        up_[i] = -PITCH;
        dn_[i] = PITCH;
        yedge_[i] = false;
        if (EDGE) {
            if (y_[i] == 0) { up_[i] = 0; yedge_[i] = true; }
            if (y_[i] == q.v.ny - 1) { dn_[i] = 0; yedge_[i] = true; }
        }
    }
    load_g_plane(s, g, q, t, zs - 2);
    load_g_plane(s, g, q, t, zs - 1);
    load_g_plane(s, g, q, t, zs);
#pragma unroll
    for (int i = 0; i < TY / NW; ++i)          // epilogue inputs of the first output plane
        if (!EDGE || nvalid_[i] > 0) epi.preload(i, zs - q.v.zg_off, y_[i], t.x, nvalid_[i]);
    // gz ring: the slot receiving gz(tz+1) advances by one per iteration (no integer division)
    int slot_new = 0;
    for (int tz = zs - 2; tz < ze; ++tz) {
        cp_async_wait_all();                // planes tz, tz+1, tz+2 were requested at least one iteration ago
        __syncthreads();
        if (tz + 3 <= ze + 1) load_g_plane(s, g, q, t, tz + 3);   // into the slot of plane tz-1 (free now)
        produce_gz<MODE>(s, q, t, ds, tz + 1, slot_new);
        if (tz >= zs) {
            produce_gy<MODE>(s, q, t, ds, tz);
            produce_gx<MODE, EDGE>(s, q, t, ds, tz);
        }
        __syncthreads();
        if (tz >= zs) {
            // gz(tz+1) lives in slot_new, gz(tz) one slot back, gz(tz-1) two slots back (mod 3)
            const int s_p = slot_new, s_c = slot_new == 0 ? 2 : slot_new - 1, s_m = slot_new == 2 ? 0 : slot_new + 1;
            int shi = s_p, slo = s_m;
            DivK kz = ds.z2;
            if (tz == 0 || tz == n - 1) {
                if (tz == 0) slo = s_c;
                if (tz == n - 1) shi = s_c;
                kz = make_divk(ds.az.d1, ds.az.r1);
            }
            const float* GZH = &s.gz[shi][0][0];
            const float* GZL = &s.gz[slo][0][0];
            const float* GZC = &s.gz[s_c][0][0];
            const float* GY = &s.gy[0][0];
            const int zb = tz - q.v.zg_off;
            epi.plane(tz);
            bool active_[TY / NW];
#pragma unroll
            for (int i = 0; i < TY / NW; ++i)      // all accumulator loads first: their latency overlaps the Hessians
                active_[i] = (!EDGE || nvalid_[i] > 0) && !epi.skip4(i, zb, y_[i], t.x, nvalid_[i]);
            if (tz + 1 < ze) {                     // next plane's epilogue inputs: a whole iteration to arrive
#pragma unroll
                for (int i = 0; i < TY / NW; ++i)
                    if (!EDGE || nvalid_[i] > 0) epi.preload(i, zb + 1, y_[i], t.x, nvalid_[i]);
            }
#pragma unroll
            for (int i = 0; i < TY / NW; ++i) {
                const bool active = active_[i];
                Hess4 h;
                if (active) {
                    const int o = off_[i];
                    DivK ky = ds.y2;
                    if (EDGE && yedge_[i]) ky = make_divk(ds.ay.d1, ds.ay.r1);
                    h.zz = diff_div4<MODE>(ld4(GZH + o), ld4(GZL + o), kz);
                    h.zy = diff_div4<MODE>(ld4(GZC + o + dn_[i]), ld4(GZC + o + up_[i]), ky);
                    h.yy = diff_div4<MODE>(ld4(GY + o + dn_[i]), ld4(GY + o + up_[i]), ky);
                    h.zx = ddx4<MODE, EDGE>(GZC + o - t.col, t, ds);
                    h.yx = ddx4<MODE, EDGE>(GY + o - t.col, t, ds);
                    h.xx = ddx4<MODE, EDGE>(&s.gx[t.warp + i * NW][0], t, ds);
                }
                epi.voxels4(i, active, zb, y_[i], t.x, nvalid_[i], h);   // every lane: epilogues may use warp votes
            }
        }
        slot_new = slot_new == 2 ? 0 : slot_new + 1;
    }
}

}  // namespace hm
