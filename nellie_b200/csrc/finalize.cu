// F9 tail + F11 — `vesselness * masks` (filtering.py:926) and _mask_volume (filtering.py:952-967):
//   V = max(acc, 0);  m = binary_opening(V > thr)  (6-/4-neighbour cross, 1 iteration, border 0);
//   out = V * m.   thr = np.percentile(positive lattice sample, 1) comes from device memory.
// One pass: the (V > thr) bits of a brick plus a 2-voxel halo are staged in shared memory, eroded
// into a second shared array (halo 1) and dilated while writing.  Algorithmic traffic 8 B/voxel.
#include "common.cuh"
#include "devmath.cuh"

namespace {

constexpr int TX = 64, TY = 8, TZ = 8;
constexpr int NTHREADS = 256;
constexpr int MX = TX + 4, MY = TY + 4, MZ = TZ + 4;   // mask tile (halo 2)
constexpr int EX = TX + 2, EY = TY + 2, EZ = TZ + 2;   // eroded tile (halo 1)

__global__ void __launch_bounds__(NTHREADS)
opening_kernel(const float* __restrict__ acc, float* __restrict__ out, nb200_vol v, const double* __restrict__ thr) {
    __shared__ unsigned char m[MZ][MY][MX];
    __shared__ unsigned char er[EZ][EY][EX];
    const bool passthrough = thr[1] == 0.0;     // no positive sample: frame returned unchanged (:959-960)
    const float cut = (float)thr[0];
    const int nbx = (v.nx + TX - 1) / TX, nby = (v.ny + TY - 1) / TY;
    long long b = blockIdx.x;
    const int bx = (int)(b % nbx); b /= nbx;
    const int by = (int)(b % nby); b /= nby;
    const int zb0 = v.zc0 + (int)b * TZ, y0 = by * TY, x0 = bx * TX;
    const long long plane = (long long)v.ny * v.nx;
    if (!passthrough) {
        for (int i = threadIdx.x; i < MZ * MY * MX; i += NTHREADS) {
            const int lx = i % MX, ly = (i / MX) % MY, lz = i / (MX * MY);
            const int zb = zb0 - 2 + lz, y = y0 - 2 + ly, x = x0 - 2 + lx;
            const int zg = zb + v.zg_off;
            unsigned char bit = 0;
            if (zb >= 0 && zb < v.nz_buf && zg >= 0 && zg < v.nz_glob && y >= 0 && y < v.ny && x >= 0 && x < v.nx)
                bit = __ldg(acc + (long long)zb * plane + (long long)y * v.nx + x) > cut ? 1 : 0;
            m[lz][ly][lx] = bit;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < EZ * EY * EX; i += NTHREADS) {
            const int lx = i % EX, ly = (i / EX) % EY, lz = i / (EX * EY);
            const int cz = lz + 1, cy = ly + 1, cx = lx + 1;   // position in m
            er[lz][ly][lx] = m[cz][cy][cx] & m[cz - 1][cy][cx] & m[cz + 1][cy][cx] & m[cz][cy - 1][cx] &
                             m[cz][cy + 1][cx] & m[cz][cy][cx - 1] & m[cz][cy][cx + 1];
        }
        __syncthreads();
    }
    const int tx = threadIdx.x % TX, ty0 = threadIdx.x / TX;
    for (int lz = 0; lz < TZ; ++lz) {
        const int zb = zb0 + lz;
        if (zb >= v.zc1) break;
        for (int ly = ty0; ly < TY; ly += NTHREADS / TX) {
            const int y = y0 + ly, x = x0 + tx;
            if (y >= v.ny || x >= v.nx) continue;
            const long long idx = (long long)zb * plane + (long long)y * v.nx + x;
            float val = acc[idx];
            val = val > 0.0f ? val : 0.0f;
            if (!passthrough) {
                const int cz = lz + 1, cy = ly + 1, cx = tx + 1;
                const unsigned char keep = er[cz][cy][cx] | er[cz - 1][cy][cx] | er[cz + 1][cy][cx] |
                                           er[cz][cy - 1][cx] | er[cz][cy + 1][cx] | er[cz][cy][cx - 1] |
                                           er[cz][cy][cx + 1];
                if (!keep) val = 0.0f;
            }
            out[idx] = val;
        }
    }
}

// ---- Z-marching opening on bit planes (fast path: nx % 4 == 0, 16-byte aligned volumes) ----------------
// A CTA (8 warps) owns a 128 x 36 window of the XY plane — 120 x 32 outputs plus a halo of 4 columns /
// 2 rows — and walks along Z.  Each incoming plane is read once (one 128-bit load per lane and row) and
// reduced to (V > thr) bits with four warp ballots per row: word k of a row holds the bit of voxel
// 4*lane + k in bit `lane` (interleaved layout, so the x +- 1 neighbours of word k are words k +- 1, and a
// one-bit shift of word 0 / 3 at the lane boundary).  Erosion and dilation by the 6-neighbour cross are
// word-wide AND / OR over three-plane rings of those rows in shared memory; out-of-frame voxels count as
// 0 (scipy border_value=0).  The output plane re-reads its accumulator values (requested two barriers
// earlier, L2-resident) and writes V * m.  HBM traffic: 4 B read + 4 B written per voxel.
namespace om {
constexpr int NT = 256, NW = 8;
constexpr int ROWS = 36, TYO = 32, TXO = 120;
constexpr int RPW = TYO / NW;          // output rows per warp

__global__ void __launch_bounds__(NT)
opening_march_kernel(const float* __restrict__ acc, float* __restrict__ out, nb200_vol v,
                     const double* __restrict__ thr, int zchunk) {
    __shared__ __align__(16) unsigned m[3][ROWS][4];
    __shared__ __align__(16) unsigned e[3][ROWS][4];
    const bool passthrough = thr[1] == 0.0;     // no positive sample: frame returned unchanged (:959-960)
    const float cut = (float)thr[0];
    const int ntx = (v.nx + TXO - 1) / TXO, nty = (v.ny + TYO - 1) / TYO;
    long long b = blockIdx.x;
    const int bx = (int)(b % ntx); b /= ntx;
    const int by = (int)(b % nty); b /= nty;
    const int zs = v.zc0 + (int)b * zchunk, ze = min(zs + zchunk, v.zc1);      // buffer planes
    if (zs >= ze) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = bx * TXO - 4 + 4 * lane;                  // first voxel of this lane's group of four
    const int yw = by * TYO - 2;                            // y of window row 0
    const bool x_in = x >= 0 && x < v.nx;                   // nx % 4 == 0: a group is inside or outside as a whole
    const bool x_out = x_in && lane >= 1 && lane <= 30;
    const long long plane = (long long)v.ny * v.nx;
    for (int t = threadIdx.x; t < 3 * ROWS * 4; t += NT) (&e[0][0][0])[t] = 0u;
    if (passthrough) {
        for (int z = zs; z < ze; ++z)
            for (int i = 0; i < RPW; ++i) {
                const int y = yw + 2 + warp * RPW + i;
                if (y < v.ny && x_out) {
                    const long long at = (long long)z * plane + (long long)y * v.nx + x;
                    const float4 a = *reinterpret_cast<const float4*>(acc + at);
                    *reinterpret_cast<float4*>(out + at) =
                        make_float4(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f), fmaxf(a.z, 0.0f), fmaxf(a.w, 0.0f));
                }
            }
        return;
    }
    // p = incoming mask plane; erosion lags one plane, the output two.  The rows of plane p+1 are requested while
    // plane p is processed (register double buffer), so no barrier ever waits for global memory.
    constexpr int RPL = (ROWS + NW - 1) / NW;               // mask rows per warp
    float4 nxt[RPL];
    auto fetch_rows = [&](int pl, float4 (&t)[RPL]) {
        const int pg = pl + v.zg_off;
        const bool z_in = pg >= 0 && pg < v.nz_glob && pl >= 0 && pl < v.nz_buf;
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int r = warp + NW * i, y = yw + r;
            // out-of-frame voxels are 0 bits whatever the threshold: -inf never exceeds it
            t[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (r < ROWS && z_in && x_in && y >= 0 && y < v.ny)
                t[i] = __ldg(reinterpret_cast<const float4*>(acc + (long long)pl * plane + (long long)y * v.nx + x));
        }
    };
    fetch_rows(zs - 2, nxt);
    for (int p = zs - 2; p <= ze + 1; ++p) {
        const int zo = p - 2;                               // output plane of this step
        const bool do_out = zo >= zs;
        float4 cur[RPW], now[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) now[i] = nxt[i];
        if (p + 1 <= ze + 1) fetch_rows(p + 1, nxt);
        if (do_out) {
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int y = yw + 2 + warp * RPW + i;
                cur[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (y < v.ny && x_out)
                    cur[i] = __ldg(reinterpret_cast<const float4*>(acc + (long long)zo * plane + (long long)y * v.nx + x));
            }
        }
        // ---- phase 1: bits of plane p ----
        {
            unsigned (*mp)[4] = m[(p + 3) % 3];
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const int r = warp + NW * i;
                if (r < ROWS) {                              // warp-uniform
                    const float4 a = now[i];
                    const unsigned w0 = __ballot_sync(0xffffffffu, a.x > cut);
                    const unsigned w1 = __ballot_sync(0xffffffffu, a.y > cut);
                    const unsigned w2 = __ballot_sync(0xffffffffu, a.z > cut);
                    const unsigned w3 = __ballot_sync(0xffffffffu, a.w > cut);
                    if (lane == 0) *reinterpret_cast<uint4*>(mp[r]) = make_uint4(w0, w1, w2, w3);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: erosion of plane p-1 (window rows 1..34) ----
        if (p - 1 >= zs - 1 && threadIdx.x < 34 * 4) {
            const int r = 1 + (threadIdx.x >> 2), k = threadIdx.x & 3;
            const unsigned (*c)[4] = m[(p + 2) % 3];
            const unsigned (*lo)[4] = m[(p + 1) % 3];
            const unsigned (*hi)[4] = m[(p + 3) % 3];
            const unsigned xm = k > 0 ? c[r][k - 1] : (c[r][3] << 1);
            const unsigned xp = k < 3 ? c[r][k + 1] : (c[r][0] >> 1);
            e[(p + 2) % 3][r][k] = c[r][k] & c[r - 1][k] & c[r + 1][k] & lo[r][k] & hi[r][k] & xm & xp;
        }
        __syncthreads();
        // ---- phase 3: dilation + output of plane p-2 ----
        if (do_out) {
            const unsigned (*c)[4] = e[(p + 1) % 3];
            const unsigned (*lo)[4] = e[(p + 0) % 3];
            const unsigned (*hi)[4] = e[(p + 2) % 3];
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int r = 2 + warp * RPW + i, y = yw + r;
                const uint4 cc = *reinterpret_cast<const uint4*>(c[r]);
                const uint4 up = *reinterpret_cast<const uint4*>(c[r - 1]);
                const uint4 dn = *reinterpret_cast<const uint4*>(c[r + 1]);
                const uint4 l = *reinterpret_cast<const uint4*>(lo[r]);
                const uint4 h = *reinterpret_cast<const uint4*>(hi[r]);
                const unsigned d0 = cc.x | up.x | dn.x | l.x | h.x | (cc.w << 1) | cc.y;
                const unsigned d1 = cc.y | up.y | dn.y | l.y | h.y | cc.x | cc.z;
                const unsigned d2 = cc.z | up.z | dn.z | l.z | h.z | cc.y | cc.w;
                const unsigned d3 = cc.w | up.w | dn.w | l.w | h.w | cc.z | (cc.x >> 1);
                if (y < v.ny && x_out) {
                    const float4 a = cur[i];
                    float4 o;
                    o.x = (d0 >> lane & 1u) ? fmaxf(a.x, 0.0f) : 0.0f;
                    o.y = (d1 >> lane & 1u) ? fmaxf(a.y, 0.0f) : 0.0f;
                    o.z = (d2 >> lane & 1u) ? fmaxf(a.z, 0.0f) : 0.0f;
                    o.w = (d3 >> lane & 1u) ? fmaxf(a.w, 0.0f) : 0.0f;
                    *reinterpret_cast<float4*>(out + (long long)zo * plane + (long long)y * v.nx + x) = o;
                }
            }
        }
    }
}
}  // namespace om

__global__ void __launch_bounds__(256)
opening_2d_kernel(const float* __restrict__ vin, float* __restrict__ out, int ny, int nx,
                  const double* __restrict__ thr) {
    const bool passthrough = thr[1] == 0.0;
    const float cut = (float)thr[0];
    const long long total = (long long)ny * nx;
    auto M = [&](int y, int x) -> bool {
        return y >= 0 && y < ny && x >= 0 && x < nx && __ldg(vin + (long long)y * nx + x) > cut;
    };
    auto E = [&](int y, int x) -> bool {
        if (y < 0 || y >= ny || x < 0 || x >= nx) return false;
        return M(y, x) && M(y - 1, x) && M(y + 1, x) && M(y, x - 1) && M(y, x + 1);
    };
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (long long)y * nx);
        float val = vin[i];
        val = val > 0.0f ? val : 0.0f;
        if (!passthrough && !(E(y, x) || E(y - 1, x) || E(y + 1, x) || E(y, x - 1) || E(y, x + 1))) val = 0.0f;
        out[i] = val;
    }
}

}  // namespace

// ---- F12: Filter._remove_edges (filtering.py:969-1000, _bbox :227-250), off by default ---------------------------
// Per Z slice: rows rmin..rmax of the bounding box of the positive response, then a band of min(margin, height) rows
// zeroed at the top and at the bottom of the box.  One CTA per slice: pass 1 finds the first / last row holding a
// positive value (warps take rows round robin, 128-bit loads when the row pitch allows), pass 2 clears the bands.
// An empty slice has the reference's degenerate box (0, 0): row 0 is "cleared", which changes nothing.
namespace {
__global__ void __launch_bounds__(256)
remove_edges_kernel(float* __restrict__ v, int nz, int ny, int nx, int margin) {
    __shared__ int s_min, s_max;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int z = blockIdx.x; z < nz; z += gridDim.x) {
        float* sl = v + (long long)z * ny * nx;
        if (threadIdx.x == 0) { s_min = ny; s_max = -1; }
        __syncthreads();
        for (int y = warp; y < ny; y += nwarp) {
            const float* row = sl + (long long)y * nx;
            bool any = false;
            for (int x = lane; x < nx; x += 32) any |= row[x] > 0.0f;
            if (__ballot_sync(0xffffffffu, any) != 0u && lane == 0) {
                atomicMin(&s_min, y);
                atomicMax(&s_max, y);
            }
        }
        __syncthreads();
        int rmin = s_min, rmax = s_max;
        if (rmax < 0) { rmin = 0; rmax = 0; }                   // _bbox of an empty slice: (0, 0, 0, 0)
        const int m = min(margin, rmax - rmin + 1);
        // rows [rmin, rmin + m) and (rmax - m, rmax]
        for (int k = warp; k < 2 * m; k += nwarp) {
            const int y = k < m ? rmin + k : rmax - (k - m);
            float* row = sl + (long long)y * nx;
            for (int x = lane; x < nx; x += 32) row[x] = 0.0f;
        }
        __syncthreads();
    }
}
}  // namespace

extern "C" {

int nb200_finalize_opening(const float* acc, float* out, const nb200_vol* vol, const double* thr, void* stream) {
    NB_REQUIRE(acc && out && vol && thr, NB200_ERR_ARG, "nb200_finalize_opening: null argument");
    NB_REQUIRE(acc != out, NB200_ERR_ARG, "nb200_finalize_opening: in-place not supported");
    const nb200_vol v = *vol;
    NB_REQUIRE(v.zc0 >= 0 && v.zc1 <= v.nz_buf && v.zc0 <= v.zc1, NB200_ERR_ARG, "nb200_finalize_opening: bad Z window");
    if (v.zc0 == v.zc1) return NB200_OK;
    {
        const int g0 = v.zc0 + v.zg_off, g1 = v.zc1 + v.zg_off;
        const int need_lo = g0 - 2 < 0 ? 0 : g0 - 2, need_hi = g1 + 1 >= v.nz_glob ? v.nz_glob - 1 : g1 + 1;
        NB_REQUIRE(need_lo - v.zg_off >= 0 && need_hi - v.zg_off < v.nz_buf, NB200_ERR_ARG,
                   "nb200_finalize_opening: Z halo of 2 planes missing");
    }
    if (v.nx % 4 == 0 && (((uintptr_t)acc | (uintptr_t)out) & 15) == 0) {
        const long long tiles = (long long)((v.nx + om::TXO - 1) / om::TXO) * ((v.ny + om::TYO - 1) / om::TYO);
        const int nzc = v.zc1 - v.zc0;
        int zchunk = 128;                       // 4 extra planes per chunk: two full waves of CTAs before going shorter
        while (zchunk > 16 && tiles * ((nzc + zchunk - 1) / zchunk) < 2LL * 8 * nb::sm_count()) zchunk /= 2;
        const long long grid = tiles * ((nzc + zchunk - 1) / zchunk);
        om::opening_march_kernel<<<(unsigned)grid, om::NT, 0, nb::as_stream(stream)>>>(acc, out, v, thr, zchunk);
        return nb::check_launch("finalize_opening(march)");
    }
    const long long nbx = (v.nx + TX - 1) / TX, nby = (v.ny + TY - 1) / TY, nbz = (v.zc1 - v.zc0 + TZ - 1) / TZ;
    opening_kernel<<<(unsigned)(nbx * nby * nbz), NTHREADS, 0, nb::as_stream(stream)>>>(acc, out, v, thr);
    return nb::check_launch("finalize_opening");
}

int nb200_finalize_opening_2d(const float* vin, float* out, int ny, int nx, const double* thr, void* stream) {
    NB_REQUIRE(vin && out && thr && ny > 0 && nx > 0 && vin != out, NB200_ERR_ARG, "nb200_finalize_opening_2d: bad argument");
    opening_2d_kernel<<<nb::grid_for((long long)ny * nx, 256, 4), 256, 0, nb::as_stream(stream)>>>(vin, out, ny, nx, thr);
    return nb::check_launch("finalize_opening_2d");
}

int nb200_remove_edges(float* v, int nz, int ny, int nx, int margin, void* stream) {
    NB_REQUIRE(v && nz >= 1 && ny >= 1 && nx >= 1 && margin >= 1, NB200_ERR_ARG, "nb200_remove_edges: bad argument");
    const int grid = nz < 8 * nb::sm_count() ? nz : 8 * nb::sm_count();
    remove_edges_kernel<<<grid, 256, 0, nb::as_stream(stream)>>>(v, nz, ny, nx, margin);
    return nb::check_launch("remove_edges");
}

}  // extern "C"
