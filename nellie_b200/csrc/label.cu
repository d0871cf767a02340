// L4/L5 — Label._get_labels on device (nellie/segmentation/labelling.py:467-509, :546-556):
//   mask = (frangi * (raw > intensity_thr)) > frangi_thr          (strict, float32)
//   3-D: mask = binary_fill_holes(mask)       = fg | background components not 6-connected to the border
//   labels = label(mask, ones(3,3[,3]))       26-/8-connectivity
//   drop components with fewer than min_area voxels
//   mask = uniform_filter(float32(mask), 3, mode="reflect") > 0.5   = (set voxels in the reflected 3^d window) >= 14 (5 in 2-D)
//   labels = label(mask) : int32, ids 1..n in raster order of each component's first voxel (scipy.ndimage.label)
//
// Masks are packed (one 32-bit membership word per 32-voxel strip of a row) from the threshold to the last labelling.
// Connected components: a CTA resolves a 32 x 8 x 8 tile (32 x 64 in 2-D) with a union-find in shared memory whose
// unions are found on whole words (one union per overlap of two x-runs) and carried out 32 at a time; a second kernel
// makes the unions that cross tile faces on global indices; tile roots walk to their root, every other voxel follows in
// one hop.  Unions always hang the larger root under the smaller one (atomicMin), so a component's root is its first
// voxel in raster order and scipy's numbering is reproduced by ranking the roots.  Background components for fill-holes
// use a virtual "outside" root (-2) that every border voxel links to.  Component sizes for the size filter are summed
// from the tiles' local counts while the tile roots are resolved.
//
// Integer / bit work: the algorithmic traffic is the frangi read + the int32 labels write = 8 B/voxel; the parent array
// (written and re-read by each of the three labellings) is honest extra traffic, see DESIGN.md.
#include "common.cuh"

namespace {

constexpr int OUTSIDE = -2;
constexpr int NOT_IN_SET = -1;
constexpr int THREADS = 256;

struct Dims {
    int nz, ny, nx;
    long long plane, total;
};

__device__ __forceinline__ int ld_parent(const int* parent, long long i) {
    return *reinterpret_cast<const volatile int*>(parent + i);
}

__device__ __forceinline__ int find_root(int* parent, int x) {
    if (x < 0) return x;   // OUTSIDE handed back by a lost atomicMin race
    int p = ld_parent(parent, x);
    while (p != x && p >= 0) {
        const int gp = ld_parent(parent, p);
        if (gp != p) parent[x] = gp;  // path halving; ancestors stay ancestors, so stale writes are harmless
        x = p;
        p = gp;
    }
    return p;  // own index for a root, OUTSIDE for a border-connected background tree
}

// read-only walk (no path halving): used by the flatten pass, where a stale halving store from
// another thread could overwrite a voxel's final root with an intermediate ancestor
__device__ __forceinline__ int find_root_ro(const int* parent, int x) {
    if (x < 0) return x;
    int p = ld_parent(parent, x);
    while (p != x && p >= 0) {
        x = p;
        p = ld_parent(parent, x);
    }
    return p;
}

__device__ __forceinline__ void unite(int* parent, int a, int b) {
    while (true) {
        a = find_root(parent, a);
        b = find_root(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // a > b, b may be OUTSIDE
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

// ---- membership words ------------------------------------------------------------------------------
// Every mask of this file is packed: one 32-bit word per 32-voxel strip of a row, word [row][xw], bit k = voxel
// x = 32 xw + k, bits beyond nx are 0 (row = z * ny + y, wpr = ceil(nx / 32) words per row).  All indices fit int32
// (the entry points refuse frames of 2^31 voxels or more).
__device__ __forceinline__ unsigned valid_bits(const Dims& d, int xw) {
    const int nvalid = d.nx - xw * 32;
    return nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
}

// members of the set `want` (1: the mask, 0: its complement inside the frame) in word (row, xw)
__device__ __forceinline__ unsigned set_word(const unsigned* __restrict__ bits, const Dims& d, int wpr, int row, int xw,
                                             unsigned char want) {
    const unsigned w = __ldg(bits + (long long)row * wpr + xw);
    return want ? w : (~w & valid_bits(d, xw));
}

// mask = (frangi * (raw > intensity_thr)) > frangi_thr, one warp per strip
__global__ void __launch_bounds__(THREADS)
threshold_bits_kernel(const float* __restrict__ frangi, const float* __restrict__ raw, int use_intensity,
                      float intensity_thresh, const double* __restrict__ thr, Dims d, unsigned* __restrict__ bits) {
    const bool none = thr[3] != 0.0;            // no samples: mask = zeros (labelling.py:475-476)
    const float cut = (float)thr[0];
    const int wpr = (d.nx + 31) / 32;
    const int nwin = d.nz * d.ny * wpr;
    const int lane = threadIdx.x & 31;
    constexpr int U = 8;                                    // strips per warp and iteration: eight loads in flight, one division
    const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
    for (int w0 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) * U; w0 < nwin; w0 += nwarps * U) {
        int row = w0 / wpr, xw = w0 - row * wpr;
        float f[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int x = xw * 32 + lane;
            f[k] = -INFINITY;                               // outside the frame: never above the cut
            if (w0 + k < nwin && x < d.nx) {
                const long long i = (long long)row * d.nx + x;
                f[k] = __ldg(frangi + i);
                if (use_intensity) f[k] = f[k] * ((__ldg(raw + i) > intensity_thresh) ? 1.0f : 0.0f);   // labelling.py:550-552
                if (f[k] != f[k]) f[k] = -INFINITY;         // NaN > cut is false
            }
            if (++xw == wpr) { xw = 0; ++row; }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const unsigned b = __ballot_sync(0xffffffffu, !none && f[k] > cut);
            if (lane == 0 && w0 + k < nwin) bits[w0 + k] = b;
        }
    }
}

// bytes -> words (entry points that take a uint8 mask); TEST: 0 -> byte != 0, 1 -> byte in 1..3 (a skeleton voxel that
// is not a junction, networking.py:758-797)
template <int TEST>
__global__ void __launch_bounds__(THREADS)
pack_bits_kernel(const unsigned char* __restrict__ mask, Dims d, unsigned* __restrict__ bits) {
    const int wpr = (d.nx + 31) / 32;
    const int nwin = d.nz * d.ny * wpr;
    const int lane = threadIdx.x & 31;
    for (int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5); w < nwin;
         w += (int)(((long long)gridDim.x * blockDim.x) >> 5)) {
        const int row = w / wpr;
        const int x = (w - row * wpr) * 32 + lane;
        bool in = false;
        if (x < d.nx) {
            const unsigned char c = mask[(long long)row * d.nx + x];
            in = TEST == 0 ? c != 0 : (c > 0 && c != 4);
        }
        const unsigned b = __ballot_sync(0xffffffffu, in);
        if (lane == 0) bits[w] = b;
    }
}

// ---- tiled CCL ------------------------------------------------------------------------------------
// A CTA resolves one tile of 32 x TY x TZ voxels in shared memory (TY * TZ = 64 rows = 64 membership words; the
// union-find runs on local indices with shared-memory atomics) and writes every voxel's parent as the GLOBAL index of
// its tile-local root (raster order inside a tile agrees with the global raster order, so a local root is the first
// voxel of its piece).  The unions are found on whole words, one thread per (row, backward neighbour row):
//   26-conn, ONE union per overlap of two x-runs: my run [a, b] touches every run of a backward row that meets
//   [a-1, b+1]; such a run either covers a-1 or a (linked by the first voxel of my run: m2) or starts at s in
//   [a+1, b+1] (linked by my voxel s-1, which sees the start diagonally: m1)
//   6-conn: the first voxel of every overlap of two runs
// ccl_border_kernel then makes the unions that cross a tile face with the same rules on global indices,
// ccl_flatten_roots_kernel lets every tile root walk to its root, and ccl_consume_kernel reads the result off (every other
// member is one hop from its root by then).
// History (512^3 frame of config #5, per labelling): voxel-wise init + merge + flatten with global atomics 2.8 ms;
// this tile scheme with per-lane bit tests on byte masks 2.9 ms (the bit fiddling, 26 K warp instructions per tile);
// word-wise as below ~1.3 ms (tile 0.55-0.62, border 0.23-0.4, roots 0.02-0.08, consume 0.35-0.48).
__device__ __forceinline__ int find_s(int* par, int x) {
    if (x < 0) return x;
    int p = *reinterpret_cast<const volatile int*>(par + x);
    while (p != x && p >= 0) {
        const int gp = *reinterpret_cast<const volatile int*>(par + p);
        if (gp != p) par[x] = gp;      // path halving (ancestors stay ancestors: a stale store is harmless)
        x = p;
        p = gp;
    }
    return p;
}

__device__ __forceinline__ void unite_s(int* par, int a, int b) {
    while (true) {
        a = find_s(par, a);
        b = find_s(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // a > b, b may be OUTSIDE
        const int old = atomicMin(par + a, b);
        if (old == a) return;
        a = old;
    }
}

struct TileGrid {
    int tiles_x, tiles_y;
};

template <int TY, int TZ, bool FULL_CONN, bool BORDER_OUTSIDE>
__global__ void __launch_bounds__(THREADS)
ccl_tile_kernel(const unsigned* __restrict__ set_bits, unsigned char want, Dims d, TileGrid tg, int* __restrict__ parent,
                unsigned* __restrict__ root_bits, int* __restrict__ area) {
    constexpr int ROWS = TY * TZ;
    constexpr int NW = THREADS / 32;
    static_assert(ROWS == 64 && THREADS == 256, "one thread per (row, backward neighbour row)");
    __shared__ unsigned bits[ROWS];
    __shared__ int par[ROWS * 32];
    __shared__ int cnt[ROWS * 32];
    __shared__ unsigned queue[NW][64];
    __shared__ unsigned short heads[ROWS * 16];
    __shared__ int n_heads;
    const int tx = blockIdx.x, ty = blockIdx.y, tz = blockIdx.z;
    const int x0 = tx * 32, y0 = ty * TY, z0 = tz * TZ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int plane = d.ny * d.nx;
    // ---- membership words and x-runs ----
    if (threadIdx.x < ROWS) {
        const int r = threadIdx.x;
        const int z = z0 + r / TY, y = y0 + r % TY;
        bits[r] = (y < d.ny && z < d.nz) ? set_word(set_bits, d, tg.tiles_x, z * d.ny + y, tx, want) : 0u;
    }
    // whole tiles outside / inside the set are the common case (background around the tubes, tube-free blocks)
    const bool row_full = threadIdx.x >= ROWS || bits[threadIdx.x] == 0xffffffffu;     // own word, written above
    const bool row_empty = threadIdx.x >= ROWS || bits[threadIdx.x] == 0u;
    const int all_full = __syncthreads_and(row_full), all_empty = __syncthreads_and(row_empty);
    const int x = x0 + lane;
    if (all_empty || all_full) {
        // full: one component, its root the tile's first voxel (every row is valid, else its word would be 0)
        const int origin = z0 * plane + y0 * d.nx + x0;
        const bool outside = BORDER_OUTSIDE && all_full && (z0 == 0 || z0 + TZ >= d.nz || y0 == 0 || y0 + TY >= d.ny ||
                                                             x0 == 0 || x0 + 32 >= d.nx);
        const int val = all_empty ? NOT_IN_SET : (outside ? OUTSIDE : origin);
#pragma unroll
        for (int k = 0; k < ROWS / NW; ++k) {
            const int r = warp + k * NW;
            const int z = z0 + r / TY, y = y0 + r % TY;
            if (y >= d.ny || z >= d.nz) continue;
            if (lane == 0) root_bits[(long long)(z * d.ny + y) * tg.tiles_x + tx] = (all_full && !outside && r == 0) ? 1u : 0u;
            if (x < d.nx) parent[z * plane + y * d.nx + x] = val;
        }
        if (all_full && !outside && area != nullptr && threadIdx.x == 0) area[origin] = ROWS * 32;
        return;
    }
#pragma unroll
    for (int k = 0; k < ROWS / NW; ++k) {
        const int r = warp + k * NW;
        const unsigned w = bits[r];
        if (w == 0u) continue;                             // nobody reads par / cnt of a voxel outside the set
        int val = NOT_IN_SET;
        if ((w >> lane) & 1u) {
            const unsigned below_zero = ~w & lt;
            val = r * 32 + (below_zero ? 32 - __clz(below_zero) : 0);
        }
        par[r * 32 + lane] = val;
        cnt[r * 32 + lane] = 0;
    }
    __syncthreads();
    // ---- unions: thread = (row r, backward neighbour row q) finds them as bit masks; the warp carries them out 32 at a
    // time from a shared queue (a thread walking its own bits alone left 2 of 32 lanes active) ----
    {
        unsigned* q_ = queue[warp];
        int n_q = 0;                                        // warp-uniform
        auto push = [&](bool has, int a, int b) {          // called by all lanes; b may be OUTSIDE
            const unsigned hb = __ballot_sync(0xffffffffu, has);
            if (hb == 0u) return;
            if (has) q_[n_q + __popc(hb & lt)] = (unsigned)a | ((unsigned)b << 16);
            n_q += __popc(hb);
            if (n_q >= 32) {
                __syncwarp();
                const unsigned v = q_[n_q - 32 + lane];
                n_q -= 32;
                unite_s(par, (int)(v & 0xffffu), (int)(short)(v >> 16));
                __syncwarp();
            }
        };
        const int r = threadIdx.x & (ROWS - 1), q = threadIdx.x / ROWS;      // q in 0..3, uniform per warp
        const unsigned w = bits[r];
        const int lz = r / TY, ly = r % TY;
        int dz, dy;
        bool active;
        if (FULL_CONN) {                                    // (0,-1), (-1,-1), (-1,0), (-1,+1)
            dz = q == 0 ? 0 : -1;
            dy = q == 0 ? -1 : q - 2;
            active = true;
        } else {                                            // (0,-1), (-1,0); q = 2: the frame border
            dz = q == 0 ? 0 : -1;
            dy = q == 0 ? -1 : 0;
            active = q < 2;
        }
        active = active && w != 0u && lz + dz >= 0 && ly + dy >= 0 && ly + dy < TY;   // else another tile: ccl_border_kernel
        const int rr = active ? r + dz * TY + dy : r;
        const unsigned nb = active ? bits[rr] : 0u;
        unsigned m1, m2;
        if (FULL_CONN) {
            m1 = w & ~nb & (nb >> 1);
            m2 = w & ~(w << 1) & (nb | (nb << 1));
        } else {
            m1 = w & nb & ~((w << 1) & (nb << 1));
            m2 = 0u;
        }
        if (!active) m1 = m2 = 0u;
        while (__any_sync(0xffffffffu, m1 != 0u)) {
            const bool has = m1 != 0u;
            const int b = has ? __ffs(m1) - 1 : 0;
            m1 &= m1 - 1u;
            push(has, r * 32 + b, FULL_CONN ? rr * 32 + b + 1 : rr * 32 + b);
        }
        if (FULL_CONN) {
            while (__any_sync(0xffffffffu, m2 != 0u)) {
                const bool has = m2 != 0u;
                const int b = has ? __ffs(m2) - 1 : 0;
                m2 &= m2 - 1u;
                push(has, r * 32 + b, ((nb >> b) & 1u) ? rr * 32 + b : rr * 32 + b - 1);
            }
        }
        if (BORDER_OUTSIDE) {
            // members on the frame border hang under OUTSIDE: one union per x-run of a border row, else the two ends
            unsigned m = 0u;
            if (q == 2 && w != 0u) {
                const int z = z0 + lz, y = y0 + ly;
                if (z == 0 || z == d.nz - 1 || y == 0 || y == d.ny - 1) m = w & ~(w << 1);
                else {
                    if (x0 == 0) m |= w & 1u;
                    const int last = d.nx - 1 - x0;
                    if (last >= 0 && last < 32) m |= w & (1u << last);
                }
            }
            while (__any_sync(0xffffffffu, m != 0u)) {
                const bool has = m != 0u;
                const int b = has ? __ffs(m) - 1 : 0;
                m &= m - 1u;
                push(has, r * 32 + b, OUTSIDE);
            }
        }
        __syncwarp();
        if (lane < n_q) {
            const unsigned v = q_[lane];
            unite_s(par, (int)(v & 0xffffu), (int)(short)(v >> 16));
        }
    }
    __syncthreads();
    // ---- tile-local roots.  Every inner node of the forest is the first voxel of an x-run (a "head": voxels start out
    // pointing at their head and only roots are ever linked), so the heads are gathered, each walks its chain once and
    // keeps the root, and every voxel is then two loads from its root.  (All 32 lanes of a row walking the same chain
    // was a third of this kernel's instructions.)  Sizes for the size filter are counted run by run. ----
    if (threadIdx.x == 0) n_heads = 0;
    __syncthreads();
    if (threadIdx.x < ROWS) {
        const int r = threadIdx.x;
        unsigned hm = bits[r] & ~(bits[r] << 1);
        if (hm) {
            int pos = atomicAdd(&n_heads, __popc(hm));
            while (hm) {
                heads[pos++] = (unsigned short)(r * 32 + __ffs(hm) - 1);
                hm &= hm - 1u;
            }
        }
    }
    __syncthreads();
    for (int h = threadIdx.x; h < n_heads; h += THREADS) {
        const int idx = heads[h];
        int x_ = idx, p_ = *reinterpret_cast<const volatile int*>(par + idx);
        while (p_ != x_ && p_ >= 0) {                       // read-only walk; the only store is to the own entry
            x_ = p_;
            p_ = *reinterpret_cast<const volatile int*>(par + x_);
        }
        par[idx] = p_;                                      // own index for a root, OUTSIDE for a border-connected tree
    }
    __syncthreads();
    int root[ROWS / NW];
#pragma unroll
    for (int k = 0; k < ROWS / NW; ++k) {
        const int r = warp + k * NW;
        const unsigned w = bits[r];
        root[k] = NOT_IN_SET;
        if ((w >> lane) & 1u) {
            const int h = par[r * 32 + lane];               // a head above me (or, for a head, its root / OUTSIDE)
            root[k] = h >= 0 ? par[h] : h;
            if (area != nullptr && root[k] >= 0 && !(lane > 0 && ((w >> (lane - 1)) & 1u))) {
                const unsigned rest = ~(w >> lane);                  // first zero above me ends the run
                atomicAdd(&cnt[root[k]], rest ? __ffs(rest) - 1 : 32 - lane);
            }
        }
    }
    if (area != nullptr) __syncthreads();
#pragma unroll
    for (int k = 0; k < ROWS / NW; ++k) {
        const int r = warp + k * NW;
        const int z = z0 + r / TY, y = y0 + r % TY;
        if (y >= d.ny || z >= d.nz) continue;                                   // warp-uniform
        int out = root[k];                                                      // NOT_IN_SET, OUTSIDE or a local index
        const bool is_root = out == r * 32 + lane;
        if (out >= 0) {
            const int rr = out >> 5;
            out = (z0 + rr / TY) * plane + (y0 + rr % TY) * d.nx + x0 + (out & 31);
        }
        const unsigned rb = __ballot_sync(0xffffffffu, is_root);
        if (lane == 0) root_bits[(long long)(z * d.ny + y) * tg.tiles_x + tx] = rb;
        if (x < d.nx) {
            const int i = z * plane + y * d.nx + x;
            parent[i] = out;
            if (area != nullptr && is_root) area[i] = cnt[r * 32 + lane];
        }
    }
}

// unions across tile faces, one THREAD per membership word (row, xw).  Completeness: two adjacent voxels u (row r) and
// v (a backward row r', or the same row) in different tiles are either
//   * in rows of different tiles (r' lies across a Y or Z face): the run-overlap rules of the tile kernel on the global
//     rows, x-neighbour bits taken from the adjacent words, or
//   * in rows of one tile, in adjacent 32-wide strips.  Then at the strip boundary xb = 32 xw:
//       same row:            (r, xb) ~ (r, xb-1)                                       the stitch of an x-run
//       u = (r, xb-1), v = (r', xb):   needed only if (r', xb-1) is not set (else u ~ (r', xb-1) inside the tile and
//                                      (r', xb-1) ~ v by the stitch)
//       u = (r, xb),   v = (r', xb-1): needed only if neither (r', xb) nor (r, xb-1) is set (same argument)
template <int TY, int TZ, bool FULL_CONN>
__global__ void __launch_bounds__(THREADS)
ccl_border_kernel(const unsigned* __restrict__ set_bits, unsigned char want, Dims d, int* __restrict__ parent) {
    __shared__ int2 queue[THREADS / 32][64];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int2* q_ = queue[threadIdx.x >> 5];
    int n_q = 0;                                            // warp-uniform
    auto push = [&](bool has, int a, int b) {              // called by all lanes of the warp
        const unsigned hb = __ballot_sync(0xffffffffu, has);
        if (hb == 0u) return;
        if (has) q_[n_q + __popc(hb & lt)] = make_int2(a, b);
        n_q += __popc(hb);
        if (n_q >= 32) {
            __syncwarp();
            const int2 pr = q_[n_q - 32 + lane];
            n_q -= 32;
            unite(parent, pr.x, pr.y);
            __syncwarp();
        }
    };
    const int wpr = (d.nx + 31) / 32;
    const int nwords = d.nz * d.ny * wpr;
    const int plane = d.ny * d.nx;
    const int stride = (int)min((long long)gridDim.x * blockDim.x, 2147483647LL - nwords);
    for (int base = (int)(blockIdx.x * (long long)blockDim.x + threadIdx.x) - lane; base < nwords; base += stride) {
        const int wi = base + lane;
        bool live = wi < nwords;
        const int row = live ? wi / wpr : 0, xw = live ? wi - row * wpr : 0;
        const int z = row / d.ny, y = row - z * d.ny;
        const bool up_y = (y % TY) == 0 && y > 0, dn_y = (y % TY) == TY - 1 && y + 1 < d.ny, up_z = (z % TZ) == 0 && z > 0;
        live = live && (xw > 0 || up_y || up_z || (FULL_CONN && dn_y && z > 0));
        unsigned w = 0u, w_prev = 0u;
        if (live) {
            w = set_word(set_bits, d, wpr, row, xw, want);
            if (xw > 0) w_prev = set_word(set_bits, d, wpr, row, xw - 1, want) >> 31;              // (row, xb - 1)
        }
        live = live && (w | w_prev) != 0u;
        if (!__any_sync(0xffffffffu, live)) continue;
        const int i0 = row * d.nx + xw * 32;                        // voxel (row, xb)
        push(live && (w & 1u) && w_prev, i0, i0 - 1);
#pragma unroll
        for (int dz = -1; dz <= 0; ++dz) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                if (dz == 0 && dy >= 0) continue;
                if (!FULL_CONN && dz != 0 && dy != 0) continue;     // 6-conn: (0,-1) and (-1,0) only
                const bool rc = (dz < 0 && up_z) || (dy < 0 && up_y) || (dy > 0 && dn_y);   // the row lies in another tile
                const bool on = live && z + dz >= 0 && y + dy >= 0 && y + dy < d.ny && (rc || (FULL_CONN && xw > 0));
                unsigned nb = 0u, nb_prev = 0u, nb_next = 0u;
                if (on) {
                    const int nrow = row + dz * d.ny + dy;
                    nb = set_word(set_bits, d, wpr, nrow, xw, want);
                    if (xw > 0) nb_prev = set_word(set_bits, d, wpr, nrow, xw - 1, want) >> 31;
                    if (FULL_CONN && rc && xw + 1 < wpr) nb_next = set_word(set_bits, d, wpr, nrow, xw + 1, want) & 1u;
                }
                const int j0 = i0 + dz * plane + dy * d.nx;          // voxel (nrow, xb)
                if (FULL_CONN) {                                     // one tile: the two diagonals over the strip boundary
                    push(on && !rc && w_prev && (nb & 1u) && !nb_prev, i0 - 1, j0);
                    push(on && !rc && (w & 1u) && nb_prev && !(nb & 1u) && !w_prev, i0, j0 - 1);
                }
                unsigned m1 = 0u, m2 = 0u;
                if (on && rc) {
                    const unsigned cm = (nb << 1) | nb_prev, w_in = (w << 1) | w_prev;
                    if (FULL_CONN) {
                        m1 = w & ~nb & ((nb >> 1) | (nb_next << 31));
                        m2 = w & ~w_in & (nb | cm);
                    } else {
                        m1 = w & nb & ~(w_in & cm);
                    }
                }
                while (__any_sync(0xffffffffu, m1 != 0u)) {
                    const bool has = m1 != 0u;
                    const int b = has ? __ffs(m1) - 1 : 0;
                    m1 &= m1 - 1u;
                    push(has, i0 + b, FULL_CONN ? j0 + b + 1 : j0 + b);
                }
                if (FULL_CONN) {
                    while (__any_sync(0xffffffffu, m2 != 0u)) {
                        const bool has = m2 != 0u;
                        const int b = has ? __ffs(m2) - 1 : 0;
                        m2 &= m2 - 1u;
                        push(has, i0 + b, ((nb >> b) & 1u) ? j0 + b : j0 + b - 1);
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane < n_q) unite(parent, q_[lane].x, q_[lane].y);
}

// parent[i] = root of i for every voxel of the set, in two kernels.  root_bits (written by the tile kernel) marks the
// voxels that were tile-local roots: only those can have been re-linked to another tile.  ccl_flatten_roots_kernel walks
// all of them at once (one thread per word; inside the tile-structured kernel of the first version one walker per CTA
// held its 255 siblings at the barrier for the 10-20 dependent L2 hops of a big component: 0.9 ms per pass);
// ccl_flatten_voxels_kernel then gives every other voxel its tile root's root, one cached hop.
// Every store is a true root (no unions run concurrently), so concurrent walks stay valid.
__global__ void __launch_bounds__(THREADS)
ccl_flatten_roots_kernel(Dims d, const unsigned* __restrict__ root_bits, int* __restrict__ parent, int* __restrict__ area) {
    const int wpr = (d.nx + 31) / 32;
    const int nwords = d.nz * d.ny * wpr;
    for (int wi = (int)(blockIdx.x * (long long)blockDim.x + threadIdx.x); wi < nwords; wi += (int)((long long)gridDim.x * blockDim.x)) {
        unsigned rb = __ldg(root_bits + wi);
        if (rb == 0u) continue;
        const int row = wi / wpr;
        const int i0 = row * d.nx + (wi - row * wpr) * 32;
        while (rb) {
            const int i = i0 + __ffs(rb) - 1;
            rb &= rb - 1u;
            const int p = ld_parent(parent, i);
            if (p >= 0 && p != i) {
                const int f = find_root_ro(parent, p);
                parent[i] = f;
                // the size of a component is the sum over its tile-local pieces; area[i] of a non-root is final
                if (area != nullptr && f >= 0) atomicAdd(area + f, area[i]);
            }
        }
    }
}

// What each labelling is FOR is read off in the same pass that would otherwise only flatten the forest: after
// ccl_flatten_roots_kernel a member is at most one hop from its root (a flagged tile root holds it, everybody else
// reads it from the tile root it points at: every ancestor is a former tile root), so this kernel never stores a parent.
//   HOLES   (background labelling): words |= members whose tree is not tied to OUTSIDE        (binary_fill_holes)
//   KEEP    (first labelling):      keep words = members of components with >= min_area voxels (area[root] is complete)
//   LABELS  (last labelling):       labels = number of the root (root_assign_kernel has run), 0 outside the set
// (Checking a root with one more load made every voxel of a big component read the same word: 0.9 ms per pass.)
enum { CONSUME_HOLES = 1, CONSUME_KEEP = 2, CONSUME_LABELS = 3 };

template <int MODE>
__global__ void __launch_bounds__(THREADS)
ccl_consume_kernel(Dims d, const unsigned* set_bits, unsigned char want, const unsigned* root_bits,
                   const int* __restrict__ parent, const int* area, long long min_area, unsigned* out_bits, int* labels) {
    const int wpr = (d.nx + 31) / 32;
    const int nwin = d.nz * d.ny * wpr;
    const int lane = threadIdx.x & 31;
    constexpr int U = 8;                                    // strips per warp and iteration: loads in flight, one division
    const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
    for (int w0 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5) * U; w0 < nwin; w0 += nwarps * U) {
        int row = w0 / wpr, xw = w0 - row * wpr;
        int idx[U], root[U];
        bool hop[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            root[k] = NOT_IN_SET;
            hop[k] = false;
            idx[k] = row * d.nx + xw * 32 + lane;
            if (w0 + k < nwin) {
                const unsigned m = set_word(set_bits, d, wpr, row, xw, want);
                if ((m >> lane) & 1u) {
                    root[k] = __ldg(parent + idx[k]);
                    hop[k] = root[k] >= 0 && !((root_bits[w0 + k] >> lane) & 1u);     // may alias out_bits (KEEP): own word, read first
                }
            }
            if (++xw == wpr) { xw = 0; ++row; }
        }
#pragma unroll
        for (int k = 0; k < U; ++k)
            if (hop[k]) root[k] = __ldg(parent + root[k]);
        if (MODE == CONSUME_HOLES) {
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const unsigned b = __ballot_sync(0xffffffffu, root[k] >= 0);
                if (lane == 0 && b) out_bits[w0 + k] |= b;
            }
        } else if (MODE == CONSUME_KEEP) {
            int a[U];
#pragma unroll
            for (int k = 0; k < U; ++k) a[k] = root[k] >= 0 ? area[root[k]] : 0;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const unsigned b = __ballot_sync(0xffffffffu, root[k] >= 0 && (long long)a[k] >= min_area);
                if (lane == 0 && w0 + k < nwin) out_bits[w0 + k] = b;
            }
        } else {
            int l[U];
#pragma unroll
            for (int k = 0; k < U; ++k) l[k] = (root[k] >= 0 && root[k] != idx[k]) ? labels[root[k]] : 0;
            row = w0 / wpr;
            xw = w0 - row * wpr;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                // a root keeps the number root_assign_kernel gave it
                if (w0 + k < nwin && xw * 32 + lane < d.nx && root[k] != idx[k]) labels[idx[k]] = l[k];
                if (++xw == wpr) { xw = 0; ++row; }
            }
        }
    }
}

// ---- 3^d majority with reflected borders, bit-sliced ------------------------------------------------
// uniform_filter(float32(mask), 3, mode="reflect") > 0.5  <=>  (set voxels in the 3^d window, indices clamped) >= 14
// (5 in 2-D).  One THREAD per membership word: per window row the three horizontally shifted words are added as bit
// planes (a 2-bit sum per voxel) and accumulated into a 5-plane counter; 32 voxels cost ~15 logic instructions per row
// instead of 3 byte loads each (the per-voxel form took 1.5 ms of a 512^3 frame, a warp-per-strip form with shuffles 2.9).
__global__ void __launch_bounds__(THREADS)
majority_bits_kernel(const unsigned* __restrict__ bits, Dims d, unsigned* __restrict__ out) {
    const int wpr = (d.nx + 31) / 32;
    const int nwords = d.nz * d.ny * wpr;
    const bool three_d = d.nz > 1;
    for (int wi = (int)(blockIdx.x * (long long)blockDim.x + threadIdx.x); wi < nwords; wi += (int)((long long)gridDim.x * blockDim.x)) {
        const int row = wi / wpr;
        const int xw = wi - row * wpr;
        const int z = row / d.ny, y = row - z * d.ny;
        const int nvalid = min(32, d.nx - xw * 32);
        unsigned t0 = 0u, t1 = 0u, t2 = 0u, t3 = 0u, t4 = 0u;
        for (int dz = three_d ? -1 : 0; dz <= (three_d ? 1 : 0); ++dz) {
            const int zz = min(max(z + dz, 0), d.nz - 1);      // reflect of a 1-voxel overhang = clamp
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = min(max(y + dy, 0), d.ny - 1);
                const unsigned* rw = bits + (long long)(zz * d.ny + yy) * wpr;
                unsigned c = __ldg(rw + xw);
                const unsigned left = xw > 0 ? (__ldg(rw + xw - 1) >> 31) : (c & 1u);
                unsigned right;
                if (nvalid < 32) {                              // the row ends inside this word: x + 1 clamps to nx - 1
                    right = (c >> (nvalid - 1)) & 1u;
                    if (right) c |= ~0u << nvalid;
                } else {
                    right = xw + 1 < wpr ? (__ldg(rw + xw + 1) & 1u) : (c >> 31);
                }
                const unsigned L = (c << 1) | left, R = (c >> 1) | (right << 31);
                const unsigned lc = L ^ c;
                const unsigned s0 = lc ^ R, s1 = (L & c) | (R & lc);
                // counter += s0 + 2 s1
                unsigned cy = t0 & s0;
                t0 ^= s0;
                const unsigned a = t1 ^ s1;
                const unsigned cy2 = (t1 & s1) | (cy & a);
                t1 = a ^ cy;
                const unsigned cy3 = t2 & cy2;
                t2 ^= cy2;
                const unsigned cy4 = t3 & cy3;
                t3 ^= cy3;
                t4 ^= cy4;
            }
        }
        unsigned res = three_d ? (t4 | (t3 & t2 & t1)) : (t4 | t3 | (t2 & (t1 | t0)));
        if (nvalid < 32) res &= (1u << nvalid) - 1u;
        out[wi] = res;
    }
}

// ---- raster-order numbering of the roots -------------------------------------------------------------
// A root of the final labelling is a tile-local root that no union re-linked, so only the voxels flagged in root_bits
// are tested.  Words are in raster order; one warp counts / numbers a chunk of 32 words, a single CTA scans the chunks.
constexpr int RANK_WORDS = 32;

__device__ __forceinline__ int count_roots(unsigned rb, int i0, const int* __restrict__ parent) {
    int c = 0;
    while (rb) {
        const int b = __ffs(rb) - 1;
        rb &= rb - 1u;
        c += parent[i0 + b] == i0 + b;
    }
    return c;
}

__global__ void __launch_bounds__(THREADS)
root_count_kernel(Dims d, const unsigned* __restrict__ root_bits, const int* __restrict__ parent, int nwords,
                  int* __restrict__ block_counts) {
    const int wpr = (d.nx + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int chunk = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int wi = chunk * RANK_WORDS + lane;
    if (chunk * RANK_WORDS >= nwords) return;               // warp-uniform
    int c = 0;
    if (wi < nwords) {
        const int row = wi / wpr;
        c = count_roots(__ldg(root_bits + wi), row * d.nx + (wi - row * wpr) * 32, parent);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) block_counts[chunk] = c;
}

// single CTA: exclusive scan of block_counts in place; total -> *n_labels.  A thread owns 8 consecutive counts per round
// (the one-count-per-thread form needed 128 rounds of four barriers for the 1.3 * 10^5 chunks of a 512^3 frame: 115 us)
__global__ void __launch_bounds__(1024)
block_scan_kernel(int* __restrict__ block_counts, long long nblocks, long long* __restrict__ n_labels) {
    constexpr int V = 8;
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long base = 0; base < nblocks; base += 1024 * V) {
        const long long i0 = base + (long long)threadIdx.x * V;
        int v[V];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            v[j] = i0 + j < nblocks ? block_counts[i0 + j] : 0;
            sum += v[j];
        }
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        if (w == 0) {
            int t = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;   // inclusive totals per warp
        }
        __syncthreads();
        const int carry = carry_s;
        int run = (w ? warp_tot[w - 1] : 0) + carry + incl - sum;      // everything before this thread's first count
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (i0 + j < nblocks) block_counts[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_labels = (long long)carry_s;
}

__global__ void __launch_bounds__(THREADS)
root_assign_kernel(Dims d, const unsigned* __restrict__ root_bits, const int* __restrict__ parent, int nwords,
                   const int* __restrict__ block_offsets, int* __restrict__ labels) {
    const int wpr = (d.nx + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int chunk = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int wi = chunk * RANK_WORDS + lane;
    if (chunk * RANK_WORDS >= nwords) return;               // warp-uniform
    unsigned rb = 0u;
    int i0 = 0;
    if (wi < nwords) {
        const int row = wi / wpr;
        rb = __ldg(root_bits + wi);
        i0 = row * d.nx + (wi - row * wpr) * 32;
    }
    const int c = count_roots(rb, i0, parent);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int next = block_offsets[chunk] + incl - c;             // roots before this word
    while (rb) {
        const int b = __ffs(rb) - 1;
        rb &= rb - 1u;
        if (parent[i0 + b] == i0 + b) labels[i0 + b] = ++next;
    }
}

Dims make_dims(int nz, int ny, int nx) {
    Dims d;
    d.nz = nz; d.ny = ny; d.nx = nx;
    d.plane = (long long)ny * nx;
    d.total = d.plane * nz;
    return d;
}

unsigned gs(long long n) { return nb::grid_for(n, THREADS, 8); }


// The labelling of the set `want` of the packed mask, up to the point where every tile root (flagged in root_bits) holds
// its root — the first voxel of its component in raster order, or OUTSIDE — and every other member points at a tile root;
// ccl_consume_kernel reads the result off.  area (optional): on return area[r] = voxels of the component with root r.
template <int TY, int TZ>
int run_ccl_tiled(const unsigned* set_bits, unsigned char want, const Dims& d, bool full_conn, bool border_outside,
                  int* parent, unsigned* root_bits, int* area, cudaStream_t st) {
    TileGrid tg;
    tg.tiles_x = (d.nx + 31) / 32;
    tg.tiles_y = (d.ny + TY - 1) / TY;
    const int tiles_z = (d.nz + TZ - 1) / TZ;
    NB_REQUIRE(tg.tiles_y <= 65535 && tiles_z <= 65535, NB200_ERR_UNSUPPORTED, "ccl: more than 65535 tiles along Y or Z");
    const dim3 g((unsigned)tg.tiles_x, (unsigned)tg.tiles_y, (unsigned)tiles_z);
    const long long words = (long long)d.nz * d.ny * tg.tiles_x;
    if (full_conn && !border_outside) ccl_tile_kernel<TY, TZ, true, false><<<g, THREADS, 0, st>>>(set_bits, want, d, tg, parent, root_bits, area);
    else if (!full_conn && border_outside) ccl_tile_kernel<TY, TZ, false, true><<<g, THREADS, 0, st>>>(set_bits, want, d, tg, parent, root_bits, area);
    else if (!full_conn) ccl_tile_kernel<TY, TZ, false, false><<<g, THREADS, 0, st>>>(set_bits, want, d, tg, parent, root_bits, area);
    else ccl_tile_kernel<TY, TZ, true, true><<<g, THREADS, 0, st>>>(set_bits, want, d, tg, parent, root_bits, area);
    if (full_conn) ccl_border_kernel<TY, TZ, true><<<gs(words), THREADS, 0, st>>>(set_bits, want, d, parent);
    else ccl_border_kernel<TY, TZ, false><<<gs(words), THREADS, 0, st>>>(set_bits, want, d, parent);
    ccl_flatten_roots_kernel<<<gs(words), THREADS, 0, st>>>(d, root_bits, parent, area);
    return nb::check_launch("ccl");
}

int run_ccl(const unsigned* set_bits, unsigned char want, const Dims& d, bool full_conn, bool border_outside,
            int* parent, unsigned* root_bits, int* area, cudaStream_t st) {
    if (d.nz > 1) return run_ccl_tiled<8, 8>(set_bits, want, d, full_conn, border_outside, parent, root_bits, area, st);
    return run_ccl_tiled<64, 1>(set_bits, want, d, full_conn, border_outside, parent, root_bits, area, st);
}

struct Workspace {
    int* parent;
    unsigned* bits_a;
    unsigned* bits_b;
    int* block_counts;
    long long words, nblocks;
};

long long al256(long long b) { return (b + 255) / 256 * 256; }

// parent int32[n] | mask words u32[words] | root / keep words u32[words] | block counts int32[nblocks], 256-byte aligned
Workspace carve(void* workspace, const Dims& d) {
    Workspace w;
    w.words = (long long)d.nz * d.ny * ((d.nx + 31) / 32);
    w.nblocks = (w.words + RANK_WORDS - 1) / RANK_WORDS;
    char* ws = static_cast<char*>(workspace);
    w.parent = reinterpret_cast<int*>(ws);
    w.bits_a = reinterpret_cast<unsigned*>(ws + al256(4 * d.total));
    w.bits_b = reinterpret_cast<unsigned*>(ws + al256(4 * d.total) + al256(4 * w.words));
    w.block_counts = reinterpret_cast<int*>(ws + al256(4 * d.total) + 2 * al256(4 * w.words));
    return w;
}

int number_components(const Dims& d, const Workspace& w, int* labels, long long* n_labels, cudaStream_t st) {
    // w.bits_b holds the tile-root words of the labelling that has just run
    const unsigned ctas = (unsigned)((w.nblocks + THREADS / 32 - 1) / (THREADS / 32));
    root_count_kernel<<<ctas, THREADS, 0, st>>>(d, w.bits_b, w.parent, (int)w.words, w.block_counts);
    block_scan_kernel<<<1, 1024, 0, st>>>(w.block_counts, w.nblocks, n_labels);
    root_assign_kernel<<<ctas, THREADS, 0, st>>>(d, w.bits_b, w.parent, (int)w.words, w.block_counts, labels);
    ccl_consume_kernel<CONSUME_LABELS><<<gs(d.total), THREADS, 0, st>>>(d, w.bits_a, 1, w.bits_b, w.parent, nullptr, 0, nullptr, labels);
    return nb::check_launch("number_components");
}

}  // namespace

extern "C" {

size_t nb200_label_workspace_bytes(int nz, int ny, int nx) {
    const Dims d = make_dims(nz, ny, nx);
    const long long words = (long long)nz * ny * ((nx + 31) / 32);
    const long long nblocks = (words + RANK_WORDS - 1) / RANK_WORDS;
    return (size_t)(al256(4 * d.total) + 2 * al256(4 * words) + al256(4 * nblocks));
}

int nb200_label_frame(const float* frangi, const float* raw, int use_intensity, float intensity_thresh,
                      const double* thr, int nz, int ny, int nx, long long min_area, int fill_holes,
                      int* labels, void* workspace, long long* n_labels, void* stream) {
    NB_REQUIRE(frangi && thr && labels && workspace && n_labels, NB200_ERR_ARG, "nb200_label_frame: null argument");
    NB_REQUIRE(!use_intensity || raw, NB200_ERR_ARG, "nb200_label_frame: raw frame required for intensity gating");
    NB_REQUIRE(nz >= 1 && ny >= 1 && nx >= 1, NB200_ERR_ARG, "nb200_label_frame: bad shape");
    const long long n = (long long)nz * ny * nx;
    NB_REQUIRE(n < 2147483000LL, NB200_ERR_UNSUPPORTED, "nb200_label_frame: frame exceeds int32 voxel indexing");
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    const Workspace w = carve(workspace, d);
    int rc;

    threshold_bits_kernel<<<gs(n), THREADS, 0, st>>>(frangi, raw, use_intensity, intensity_thresh, thr, d, w.bits_a);
    if (fill_holes && nz > 1) {   // labelling.py:485-486 (3-D only)
        rc = run_ccl(w.bits_a, 0, d, /*full_conn=*/false, /*border_outside=*/true, w.parent, w.bits_b, nullptr, st);
        if (rc) return rc;
        ccl_consume_kernel<CONSUME_HOLES><<<gs(n), THREADS, 0, st>>>(d, w.bits_a, 0, w.bits_b, w.parent, nullptr, 0, w.bits_a, nullptr);
    }
    // first labelling + size filter (labelling.py:489-501); the int32 output doubles as the size table
    rc = run_ccl(w.bits_a, 1, d, true, false, w.parent, w.bits_b, /*area=*/labels, st);
    if (rc) return rc;
    // keep words overwrite the root flags word by word (each word is read before it is written, by the same warp)
    ccl_consume_kernel<CONSUME_KEEP><<<gs(n), THREADS, 0, st>>>(d, w.bits_a, 1, w.bits_b, w.parent, labels, min_area, w.bits_b, nullptr);
    // smoothing (labelling.py:503-505) and second labelling (:507)
    majority_bits_kernel<<<gs(w.words), THREADS, 0, st>>>(w.bits_b, d, w.bits_a);
    rc = run_ccl(w.bits_a, 1, d, true, false, w.parent, w.bits_b, nullptr, st);
    if (rc) return rc;
    return number_components(d, w, labels, n_labels, st);
}

/* scipy.ndimage.label(mask, structure=ones) alone (tests, the Z-sharded Label, the Network stage):
 * mask: uint8 device array; connectivity_full: 1 = 26/8, 0 = 6/4. */
int nb200_ccl_label(const unsigned char* mask, int nz, int ny, int nx, int connectivity_full, int* labels,
                    void* workspace, long long* n_labels, void* stream) {
    NB_REQUIRE(mask && labels && workspace && n_labels, NB200_ERR_ARG, "nb200_ccl_label: null argument");
    NB_REQUIRE(nz >= 1 && ny >= 1 && nx >= 1, NB200_ERR_ARG, "nb200_ccl_label: bad shape");
    const long long n = (long long)nz * ny * nx;
    NB_REQUIRE(n < 2147483000LL, NB200_ERR_UNSUPPORTED, "nb200_ccl_label: frame exceeds int32 voxel indexing");
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    const Workspace w = carve(workspace, d);
    pack_bits_kernel<0><<<gs(n), THREADS, 0, st>>>(mask, d, w.bits_a);
    const int rc = run_ccl(w.bits_a, 1, d, connectivity_full != 0, false, w.parent, w.bits_b, nullptr, st);
    if (rc) return rc;
    return number_components(d, w, labels, n_labels, st);
}

}  // extern "C"

// ==================================================================================================================
// Network stage, the three array kernels the reference runs on its GPU backend:
//   nellie/segmentation/networking.py:669-680  _get_pixel_class_impl            3^d neighbour count of the skeleton
//   nellie/segmentation/networking.py:758-797  _get_branch_skel_labels          label() of the non-junction skeleton
//   nellie/segmentation/networking.py:261-296  _remove_connected_label_pixels   3^d min / max label filters
// (skeletonize, _add_missing_skeleton_labels and _relabel_objects stay on the host in the reference as well.)
// ==================================================================================================================
namespace {

// out = skel > 0 ? min(4, number of set voxels in the 3^d window, zero outside the frame) : 0      (uint8)
__global__ void __launch_bounds__(THREADS)
pixel_class_kernel(const int* __restrict__ skel, Dims d, unsigned char* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d.total;
         i += (long long)gridDim.x * blockDim.x) {
        unsigned char res = 0;
        if (__ldg(skel + i) > 0) {
            const int z = (int)(i / d.plane);
            const long long rem = i - (long long)z * d.plane;
            const int y = (int)(rem / d.nx), x = (int)(rem - (long long)y * d.nx);
            int cnt = 0;
            for (int dz = (d.nz > 1 ? -1 : 0); dz <= (d.nz > 1 ? 1 : 0); ++dz) {
                const int zz = z + dz;
                if (zz < 0 || zz >= d.nz) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= d.ny) continue;
                    const int* row = skel + (long long)zz * d.plane + (long long)yy * d.nx;
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = x + dx;
                        if (xx >= 0 && xx < d.nx) cnt += __ldg(row + xx) > 0 ? 1 : 0;
                    }
                }
            }
            res = (unsigned char)(cnt > 4 ? 4 : cnt);
        }
        out[i] = res;
    }
}

// out = label, except for labelled voxels off the frame boundary whose 3^d window holds two different positive labels
__global__ void __launch_bounds__(THREADS)
remove_connected_kernel(const int* __restrict__ labels, Dims d, int* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d.total;
         i += (long long)gridDim.x * blockDim.x) {
        const int me = __ldg(labels + i);
        int res = me;
        if (me > 0) {
            const int z = (int)(i / d.plane);
            const long long rem = i - (long long)z * d.plane;
            const int y = (int)(rem / d.nx), x = (int)(rem - (long long)y * d.nx);
            const bool boundary = y == 0 || y == d.ny - 1 || x == 0 || x == d.nx - 1 || (d.nz > 1 && (z == 0 || z == d.nz - 1));
            if (!boundary) {
                int lo = me, hi = me;
                for (int dz = (d.nz > 1 ? -1 : 0); dz <= (d.nz > 1 ? 1 : 0); ++dz)
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int* row = labels + i + (long long)dz * d.plane + (long long)dy * d.nx;
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int v = __ldg(row + dx);
                            if (v > 0) { lo = min(lo, v); hi = max(hi, v); }
                        }
                    }
                if (lo != hi) res = 0;
            }
        }
        out[i] = res;
    }
}

}  // namespace

extern "C" {

int nb200_pixel_class(const int* skel, int nz, int ny, int nx, unsigned char* out, void* stream) {
    NB_REQUIRE(skel && out && nz >= 1 && ny >= 1 && nx >= 1, NB200_ERR_ARG, "nb200_pixel_class: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    pixel_class_kernel<<<gs(d.total), THREADS, 0, nb::as_stream(stream)>>>(skel, d, out);
    return nb::check_launch("pixel_class");
}

int nb200_branch_labels(const unsigned char* pixel_class, int nz, int ny, int nx, int* labels, void* workspace,
                        long long* n_labels, void* stream) {
    NB_REQUIRE(pixel_class && labels && workspace && n_labels, NB200_ERR_ARG, "nb200_branch_labels: null argument");
    NB_REQUIRE(nz >= 1 && ny >= 1 && nx >= 1, NB200_ERR_ARG, "nb200_branch_labels: bad shape");
    const long long n = (long long)nz * ny * nx;
    NB_REQUIRE(n < 2147483000LL, NB200_ERR_UNSUPPORTED, "nb200_branch_labels: frame exceeds int32 voxel indexing");
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    const Workspace w = carve(workspace, d);
    pack_bits_kernel<1><<<gs(n), THREADS, 0, st>>>(pixel_class, d, w.bits_a);        // class 1..3: not a junction
    const int rc = run_ccl(w.bits_a, 1, d, true, false, w.parent, w.bits_b, nullptr, st);
    if (rc) return rc;
    return number_components(d, w, labels, n_labels, st);
}

int nb200_remove_connected_label_pixels(const int* labels, int nz, int ny, int nx, int* out, void* stream) {
    NB_REQUIRE(labels && out && labels != out && nz >= 1 && ny >= 1 && nx >= 1, NB200_ERR_ARG,
               "nb200_remove_connected_label_pixels: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    remove_connected_kernel<<<gs(d.total), THREADS, 0, nb::as_stream(stream)>>>(labels, d, out);
    return nb::check_launch("remove_connected_label_pixels");
}

}  // extern "C"
