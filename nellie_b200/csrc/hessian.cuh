// Finite-difference Hessian of numpy.gradient(numpy.gradient(g)) — nellie/segmentation/filtering.py:446-551.
//
// first derivative along an axis of extent n with spacing h (float32 array, Python
// float spacing):   interior (g[c+1]-g[c-1]) / fl32(2h);   c==0: (g[1]-g[0]) / fl32(h);
// c==n-1: (g[n-1]-g[n-2]) / fl32(h).  A second derivative is the same rule applied to the first
// derivative (including its one-sided edge values).  The reference's names x,y,z are axes 0,1,2
// = Z,Y,X:  hxx=d0d0, hxy=d1d0, hxz=d2d0, hyy=d1d1, hyz=d2d1, hzz=d2d2  (filtering.py:518-536).
#pragma once
#include "devmath.cuh"

namespace nb {

struct Spacing3 {
    float h1[3];  // fl32(h)   per axis (Z,Y,X)
    float h2[3];  // fl32(2h)
    float r1[3];  // RN(1 / fl32(h)), RN(1 / fl32(2h)): only read by the verified constant-divisor modes
    float r2[3];
};

// one-sided / central difference selector along an axis: returns the two sample offsets and divisor
struct FdTap {
    int hi, lo;   // coordinate offsets (+1/-1, +1/0, 0/-1)
    float div, rcp;
};
NB_HD FdTap fd_tap(int c, int n, float h1, float h2, float r1 = 0.0f, float r2 = 0.0f) {
    FdTap t;
    if (c == 0) { t.hi = 1; t.lo = 0; t.div = h1; t.rcp = r1; }
    else if (c == n - 1) { t.hi = 0; t.lo = -1; t.div = h1; t.rcp = r1; }
    else { t.hi = 1; t.lo = -1; t.div = h2; t.rcp = r2; }
    return t;
}

// (hi - lo) / d.  MODE 0: IEEE division.  MODE 1 (NB200_DIV_FAST): q0 = n*r, q = fma(fma(-q0, d, n), r, q0),
// bit-identical to IEEE for the divisors and numerator range nb200_divisor_mode verified (same sequence as
// hessian_march.cuh).  MODE 2 (NB200_DIV_POW2): exact reciprocal product.
template <int MODE>
NB_HD float fd_div_m(float hi, float lo, const FdTap& t) {
    const float n = hi - lo;
    if (MODE == 2) return n * t.rcp;
    if (MODE == 1) {
        const float q0 = n * t.rcp;
        const float e = fmaf(-q0, t.div, n);
        return fmaf(e, t.rcp, q0);
    }
    return n / t.div;
}

// Hessian at (z,y,x) in GLOBAL frame coordinates; `G(dz,dy,dx)` returns the blurred value at an
// offset from the voxel.  n[3] = global extents.
template <class Load, int MODE = 0>
NB_HD void hessian3(const Load& G, int z, int y, int x, const int n[3], const Spacing3& s,
                    float& hzz_, float& hzy_, float& hzx_, float& hyy_, float& hyx_, float& hxx_) {
    const FdTap tz = fd_tap(z, n[0], s.h1[0], s.h2[0], s.r1[0], s.r2[0]);
    const FdTap ty = fd_tap(y, n[1], s.h1[1], s.h2[1], s.r1[1], s.r2[1]);
    const FdTap tx = fd_tap(x, n[2], s.h1[2], s.h2[2], s.r1[2], s.r2[2]);
    // first derivative along Z evaluated at an offset position (dz,dy,dx)
    auto dZ = [&](int dz, int dy, int dx) -> float {
        const FdTap t = fd_tap(z + dz, n[0], s.h1[0], s.h2[0], s.r1[0], s.r2[0]);
        return fd_div_m<MODE>(G(dz + t.hi, dy, dx), G(dz + t.lo, dy, dx), t);
    };
    auto dY = [&](int dz, int dy, int dx) -> float {
        const FdTap t = fd_tap(y + dy, n[1], s.h1[1], s.h2[1], s.r1[1], s.r2[1]);
        return fd_div_m<MODE>(G(dz, dy + t.hi, dx), G(dz, dy + t.lo, dx), t);
    };
    auto dX = [&](int dz, int dy, int dx) -> float {
        const FdTap t = fd_tap(x + dx, n[2], s.h1[2], s.h2[2], s.r1[2], s.r2[2]);
        return fd_div_m<MODE>(G(dz, dy, dx + t.hi), G(dz, dy, dx + t.lo), t);
    };
    // reference naming: axis0 = "x" (Z), axis1 = "y" (Y), axis2 = "z" (X)
    hzz_ = fd_div_m<MODE>(dZ(tz.hi, 0, 0), dZ(tz.lo, 0, 0), tz);   // d0 d0   ("hxx")
    hzy_ = fd_div_m<MODE>(dZ(0, ty.hi, 0), dZ(0, ty.lo, 0), ty);   // d1 d0   ("hxy")
    hzx_ = fd_div_m<MODE>(dZ(0, 0, tx.hi), dZ(0, 0, tx.lo), tx);   // d2 d0   ("hxz")
    hyy_ = fd_div_m<MODE>(dY(0, ty.hi, 0), dY(0, ty.lo, 0), ty);   // d1 d1   ("hyy")
    hyx_ = fd_div_m<MODE>(dY(0, 0, tx.hi), dY(0, 0, tx.lo), tx);   // d2 d1   ("hyz")
    hxx_ = fd_div_m<MODE>(dX(0, 0, tx.hi), dX(0, 0, tx.lo), tx);   // d2 d2   ("hzz")
}

// 2-D: axes 0,1 = Y,X; reference names hxx=d0d0, hxy=d1d0, hyy=d1d1 (filtering.py:477-486)
template <class Load>
NB_HD void hessian2(const Load& G, int y, int x, int ny, int nx, const float h1[2], const float h2[2],
                    float& h00, float& h01, float& h11) {
    const FdTap ty = fd_tap(y, ny, h1[0], h2[0]);
    const FdTap tx = fd_tap(x, nx, h1[1], h2[1]);
    auto dY = [&](int dy, int dx) -> float {
        const FdTap t = fd_tap(y + dy, ny, h1[0], h2[0]);
        return fd_div(G(dy + t.hi, dx), G(dy + t.lo, dx), t.div);
    };
    auto dX = [&](int dy, int dx) -> float {
        const FdTap t = fd_tap(x + dx, nx, h1[1], h2[1]);
        return fd_div(G(dy, dx + t.hi), G(dy, dx + t.lo), t.div);
    };
    h00 = fd_div(dY(ty.hi, 0), dY(ty.lo, 0), ty.div);
    h01 = fd_div(dY(0, tx.hi), dY(0, tx.lo), tx.div);
    h11 = fd_div(dX(0, tx.hi), dX(0, tx.lo), tx.div);
}

}  // namespace nb
