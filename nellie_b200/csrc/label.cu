// L4/L5 — Label._get_labels on device (nellie/segmentation/labelling.py:467-509, :546-556):
//   mask = (frangi * (raw > intensity_thr)) > frangi_thr          (strict, float32)
//   3-D: mask = binary_fill_holes(mask)       = fg | background components not 6-connected to the border
//   labels = label(mask, ones(3,3[,3]))       26-/8-connectivity
//   drop components with fewer than min_area voxels
//   mask = uniform_filter(float32(mask), 3, mode="reflect") > 0.5   = (set voxels in the reflected 3^d window) >= 14 (5 in 2-D)
//   labels = label(mask) : int32, ids 1..n in raster order of each component's first voxel (scipy.ndimage.label)
//
// Connected components: union-find over voxel indices in global memory.  X-runs are pre-linked
// with warp ballots (every voxel starts pointing at the first voxel of its run inside a 32-wide
// window), so only run heads ever take part in atomics; unions always hang the larger root under
// the smaller one (atomicMin), so a component's root is its first voxel in raster order and
// scipy's numbering is reproduced by ranking the roots.  Background components for fill-holes use
// a virtual "outside" root (-2) that every border voxel links to.
//
// Integer / byte work: the bound is HBM traffic (frangi read + int32 labels write = 8 B/voxel
// algorithmic; the parent array and the byte masks are honest extra traffic).
#include "common.cuh"

#include <stdlib.h>

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

__device__ __forceinline__ void link_outside(int* parent, int a) {
    while (true) {
        a = find_root(parent, a);
        if (a < 0) return;
        const int old = atomicMin(parent + a, OUTSIDE);
        if (old == a) return;
        a = old;
    }
}

// ---- threshold ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
threshold_mask_kernel(const float* __restrict__ frangi, const float* __restrict__ raw, int use_intensity,
                      float intensity_thresh, const double* __restrict__ thr, long long n,
                      unsigned char* __restrict__ mask) {
    const bool none = thr[3] != 0.0;            // no samples: mask = zeros (labelling.py:475-476)
    const float cut = (float)thr[0];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float f = __ldg(frangi + i);
        if (use_intensity) f = f * ((__ldg(raw + i) > intensity_thresh) ? 1.0f : 0.0f);   // labelling.py:550-552
        mask[i] = (!none && f > cut) ? 1 : 0;
    }
}

// ---- CCL ---------------------------------------------------------------------------------------
// `want` selects the set: voxels with mask == want are in the set.
__global__ void __launch_bounds__(THREADS)
ccl_init_kernel(const unsigned char* __restrict__ mask, unsigned char want, Dims d, int* __restrict__ parent) {
    // one warp per 32-wide x window; windows per row = ceil(nx/32)
    const int wpr = (d.nx + 31) / 32;
    const long long nwin = (long long)d.nz * d.ny * wpr;
    const int lane = threadIdx.x & 31;
    for (long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; w < nwin;
         w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const long long row = w / wpr;
        const int x = (int)(w - row * wpr) * 32 + lane;
        const long long idx = row * d.nx + x;
        const bool in = x < d.nx && mask[idx] == want;
        const unsigned bits = __ballot_sync(0xffffffffu, in);
        if (x < d.nx) {
            int val = NOT_IN_SET;
            if (in) {
                const unsigned below_zero = ~bits & ((1u << lane) - 1u);
                const int start = below_zero ? 32 - __clz(below_zero) : 0;
                val = (int)(idx - lane + start);
            }
            parent[idx] = val;
        }
    }
}

// Unions are found in lock step (every lane tests the same neighbour relation of its own voxel) but carried out
// through a per-warp shared-memory queue of (a, b) pairs: 32 pairs are united at a time, one per lane, so the
// pointer-chasing find / atomicMin loops run with full warps instead of the one or two lanes per instruction
// that a voxel-by-voxel unite() leaves active (measured: 1.2-1.7 active threads per instruction, 9-16 ms per merge
// of a 512^3 frame).
template <bool FULL_CONN, bool BORDER_OUTSIDE>
__global__ void __launch_bounds__(THREADS)
ccl_merge_kernel(const unsigned char* __restrict__ mask, unsigned char want, Dims d, int* __restrict__ parent) {
    __shared__ int2 queue[THREADS / 32][64];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int2* q = queue[threadIdx.x >> 5];
    int n_q = 0;                                          // warp-uniform
    auto push = [&](bool has, int a, int b) {            // called by all lanes of the warp
        const unsigned bits = __ballot_sync(0xffffffffu, has);
        if (bits == 0u) return;
        if (has) q[n_q + __popc(bits & lt)] = make_int2(a, b);
        n_q += __popc(bits);
        if (n_q >= 32) {
            __syncwarp();
            const int2 pr = q[n_q - 32 + lane];
            n_q -= 32;
            unite(parent, pr.x, pr.y);
            __syncwarp();
        }
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); base < d.total; base += stride) {
        const long long i = base + lane;
        const bool valid = i < d.total && mask[i] == want;
        // a 32-voxel strip without a member of the set has nothing to link (every test below starts from `valid`):
        // the foreground of a frangi frame is a few percent of the voxels, so most strips end here
        if (__ballot_sync(0xffffffffu, valid) == 0u) continue;
        int z = 0, y = 0, x = 0;
        if (valid) {
            z = (int)(i / d.plane);
            const long long rem = i - (long long)z * d.plane;
            y = (int)(rem / d.nx);
            x = (int)(rem - (long long)y * d.nx);
        }
        const int me = (int)i;
        auto in = [&](int dz, int dy, int dx) -> bool {
            const int zz = z + dz, yy = y + dy, xx = x + dx;
            if (!valid || zz < 0 || yy < 0 || yy >= d.ny || xx < 0 || xx >= d.nx) return false;
            return mask[i + (long long)dz * d.plane + (long long)dy * d.nx + dx] == want;
        };
        auto off = [&](int dz, int dy, int dx) -> int {
            return (int)(i + (long long)dz * d.plane + (long long)dy * d.nx + dx);
        };
        const bool w_in = in(0, 0, -1);
        push(w_in && (x & 31) == 0, me, me - 1);               // stitch 32-wide windows of one run
        if (!FULL_CONN) {
            // 6-/4-connectivity: link to the row above / plane above once per overlap of two runs
            push(in(0, -1, 0) && !(w_in && in(0, -1, -1)), me, off(0, -1, 0));
            push(in(-1, 0, 0) && !(w_in && in(-1, 0, -1)), me, off(-1, 0, 0));
        } else {
            // 26-/8-connectivity, backward half, ONE union per overlap of two x-runs: my run [a,b] touches every
            // run of a backward row that meets [a-1, b+1].  Such a run either covers a-1 or a (linked by the first
            // voxel of my run) or starts at s in [a+1, b+1] (linked by my voxel s-1, which sees the start diagonally).
#pragma unroll
            for (int dz = -1; dz <= 0; ++dz) {
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    if (dz == 0 && dy >= 0) continue;          // same plane: only the row above
                    const bool c0 = in(dz, dy, 0);
                    const bool cp = in(dz, dy, 1), cm = in(dz, dy, -1);
                    push(valid && !c0 && cp, me, off(dz, dy, 1));
                    push(valid && !w_in && (c0 || cm), me, c0 ? off(dz, dy, 0) : off(dz, dy, -1));
                }
            }
        }
        if (BORDER_OUTSIDE) {
            const bool edge = z == 0 || z == d.nz - 1 || y == 0 || y == d.ny - 1 || x == 0 || x == d.nx - 1;
            push(valid && edge, me, OUTSIDE);
        }
    }
    __syncwarp();
    if (lane < n_q) unite(parent, q[lane].x, q[lane].y);
}

// One warp per 32-voxel strip of a row.  After ccl_init every voxel of an x-run inside the strip points at the run's
// first voxel, so only that voxel (parent outside the strip, or itself) has to chase pointers; the rest of the run takes
// the root from its lane by shuffle.  (One find per voxel cost 1.8 ms per pass of a 512^3 frame, three passes per frame.)
__global__ void __launch_bounds__(THREADS)
ccl_flatten_kernel(Dims d, int* __restrict__ parent) {
    const int wpr = (d.nx + 31) / 32;
    const long long nwin = (long long)d.nz * d.ny * wpr;
    const int lane = threadIdx.x & 31;
    for (long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; w < nwin;
         w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const long long row = w / wpr;
        const int x = (int)(w - row * wpr) * 32 + lane;
        const long long idx = row * d.nx + x;
        const long long first = idx - lane;                    // index of the strip's first voxel
        const int p = x < d.nx ? ld_parent(parent, idx) : NOT_IN_SET;
        // a voxel whose parent is another voxel of this strip (its run head, possibly re-linked by path halving to any
        // ancestor inside the strip) copies that lane's result; chains inside a strip only point backwards
        const bool local = p != NOT_IN_SET && p >= 0 && (long long)p >= first && (long long)p < idx;
        int r = p;
        if (p != NOT_IN_SET && !local) r = find_root_ro(parent, (int)idx);
        // resolve local links: at most 5 rounds (a chain of backward links inside 32 lanes halves... each round follows
        // one link, the lanes it lands on are final after as many rounds as the chain is long; chains are short (run
        // head, or one halving step), loop until no lane changes
        unsigned pending = __ballot_sync(0xffffffffu, local);
        int src = local ? (int)((long long)p - first) : lane;
        while (pending) {
            const int rr = __shfl_sync(0xffffffffu, r, src);
            const bool src_pending = (pending >> src) & 1u;
            if (local && ((pending >> lane) & 1u) && !src_pending) r = rr;
            const unsigned now = __ballot_sync(0xffffffffu, local && ((pending >> lane) & 1u) && src_pending);
            if (now == pending) break;                           // cannot happen (links point backwards); never spin
            pending = now;
        }
        if (p != NOT_IN_SET && x < d.nx) parent[idx] = r;
    }
}

// ---- tiled CCL (round 2) ---------------------------------------------------------------------------
// A CTA resolves one tile of 32 x TY x TZ voxels in shared memory (row = one 32-bit word of membership bits, the
// union-find runs on local indices with shared-memory atomics), and writes every voxel's parent as the GLOBAL index of
// its tile-local root (raster order inside a tile agrees with the global raster order, so the local root is the
// component's first voxel there).  ccl_border_kernel then makes only the unions that cross a tile face — the rule of
// ccl_merge_kernel restricted to pairs in different tiles — and ccl_flatten_tile_kernel resolves tile roots first and
// everything else in one or two cached hops.  Measured on the 512^3 frame of config #5: init + merge + flatten
// 0.36 + 1.53 + 0.9 ms per labelling before (the global atomics' latency, not the scan, was the cost).
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

__device__ __forceinline__ int find_s_ro(const int* par, int x) {
    int p = *reinterpret_cast<const volatile int*>(par + x);
    while (p != x && p >= 0) {
        x = p;
        p = *reinterpret_cast<const volatile int*>(par + x);
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
ccl_tile_kernel(const unsigned char* __restrict__ mask, unsigned char want, Dims d, TileGrid tg, int* __restrict__ parent,
                unsigned* __restrict__ root_bits, int* __restrict__ area) {
    constexpr int ROWS = TY * TZ;
    constexpr int NW = THREADS / 32;
    static_assert(ROWS * 32 <= 32768, "local indices are queued as 16-bit values");
    __shared__ unsigned bits[ROWS];
    __shared__ int par[ROWS * 32];
    __shared__ int cnt[ROWS * 32];
    __shared__ unsigned queue[NW][64];
    long long t = blockIdx.x;
    const int tx = (int)(t % tg.tiles_x);
    t /= tg.tiles_x;
    const int ty = (int)(t % tg.tiles_y), tz = (int)(t / tg.tiles_y);
    const int x0 = tx * 32, y0 = ty * TY, z0 = tz * TZ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int x = x0 + lane;
    // ---- membership words and x-runs ----
    unsigned char mv[ROWS / NW];
#pragma unroll
    for (int k = 0; k < ROWS / NW; ++k) {                  // all loads in flight before the first ballot
        const int r = warp + k * NW;
        const int z = z0 + r / TY, y = y0 + r % TY;
        mv[k] = (x < d.nx && y < d.ny && z < d.nz) ? mask[(long long)z * d.plane + (long long)y * d.nx + x] : (unsigned char)(want ^ 1);
    }
#pragma unroll
    for (int k = 0; k < ROWS / NW; ++k) {
        const int r = warp + k * NW;
        const bool in = mv[k] == want;
        const unsigned w = __ballot_sync(0xffffffffu, in);
        if (lane == 0) bits[r] = w;
        int val = NOT_IN_SET;
        if (in) {
            const unsigned below_zero = ~w & lt;
            val = r * 32 + (below_zero ? 32 - __clz(below_zero) : 0);
        }
        par[r * 32 + lane] = val;
        cnt[r * 32 + lane] = 0;
    }
    __syncthreads();
    // ---- unions, found in lock step and carried out 32 at a time (see ccl_merge_kernel) ----
    unsigned* q = queue[warp];
    int n_q = 0;                                           // warp-uniform
    auto push = [&](bool has, int a, int b) {             // called by all lanes; b may be OUTSIDE
        const unsigned hb = __ballot_sync(0xffffffffu, has);
        if (hb == 0u) return;
        if (has) q[n_q + __popc(hb & lt)] = (unsigned)a | ((unsigned)b << 16);
        n_q += __popc(hb);
        if (n_q >= 32) {
            __syncwarp();
            const unsigned v = q[n_q - 32 + lane];
            n_q -= 32;
            unite_s(par, (int)(v & 0xffffu), (int)(short)(v >> 16));
            __syncwarp();
        }
    };
#pragma unroll 1
    for (int r = warp; r < ROWS; r += NW) {
        const unsigned w = bits[r];
        if (w == 0u) continue;
        const int lz = r / TY, ly = r % TY;
        const bool me = (w >> lane) & 1u;
        const bool w_in = lane > 0 && ((w >> (lane - 1)) & 1u);
        const int idx = r * 32 + lane;
        if (!FULL_CONN) {
            if (ly > 0) {
                const unsigned nb = bits[r - 1];
                push(me && ((nb >> lane) & 1u) && !(w_in && ((nb >> (lane - 1)) & 1u)), idx, idx - 32);
            }
            if (lz > 0) {
                const unsigned nb = bits[r - TY];
                push(me && ((nb >> lane) & 1u) && !(w_in && ((nb >> (lane - 1)) & 1u)), idx, idx - TY * 32);
            }
        } else {
#pragma unroll
            for (int dz = -1; dz <= 0; ++dz) {
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    if (dz == 0 && dy >= 0) continue;
                    if (lz + dz < 0 || ly + dy < 0 || ly + dy >= TY) continue;       // another tile: ccl_face_kernel
                    const int rr = r + dz * TY + dy;
                    const unsigned nb = bits[rr];
                    if (nb == 0u) continue;
                    const bool c0 = (nb >> lane) & 1u;
                    const bool cp = lane < 31 && ((nb >> (lane + 1)) & 1u);
                    const bool cm = lane > 0 && ((nb >> (lane - 1)) & 1u);
                    push(me && !c0 && cp, idx, rr * 32 + lane + 1);
                    push(me && !w_in && (c0 || cm), idx, c0 ? rr * 32 + lane : rr * 32 + lane - 1);
                }
            }
        }
        if (BORDER_OUTSIDE) {
            const int z = z0 + lz, y = y0 + ly;
            const bool edge = z == 0 || z == d.nz - 1 || y == 0 || y == d.ny - 1 || x == 0 || x == d.nx - 1;
            push(me && edge, idx, OUTSIDE);
        }
    }
    __syncwarp();
    if (lane < n_q) {
        const unsigned v = q[lane];
        unite_s(par, (int)(v & 0xffffu), (int)(short)(v >> 16));
    }
    __syncthreads();
    // ---- tile-local roots (and, for the size filter, voxels per local component, counted run by run) ----
    int root[ROWS / NW];
#pragma unroll
    for (int k = 0; k < ROWS / NW; ++k) {
        const int r = warp + k * NW;
        const unsigned w = bits[r];
        root[k] = NOT_IN_SET;
        if ((w >> lane) & 1u) {
            root[k] = find_s_ro(par, r * 32 + lane);
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
            out = (int)((long long)(z0 + rr / TY) * d.plane + (long long)(y0 + rr % TY) * d.nx + x0 + (out & 31));
        }
        const unsigned rb = __ballot_sync(0xffffffffu, is_root);
        if (lane == 0) root_bits[((long long)z * d.ny + y) * tg.tiles_x + tx] = rb;
        if (x < d.nx) {
            const long long i = (long long)z * d.plane + (long long)y * d.nx + x;
            parent[i] = out;
            if (area != nullptr && is_root) area[i] = cnt[r * 32 + lane];
        }
    }
}

// unions across tile faces.  Completeness: two 26-adjacent voxels u (row r) and v (backward row r', or the same row)
// in different tiles are either
//   * in rows of different tiles (r' lies across a Y or Z face): the rule of ccl_merge_kernel for that row pair, all
//     lanes (ccl_face_kernel: one warp per row that has such a neighbour row, nothing else is touched), or
//   * in rows of one tile, in adjacent 32-wide strips (ccl_seam_x_kernel, one THREAD per strip boundary xb):
//       same row:            (r, xb) ~ (r, xb-1)                                       the stitch of an x-run
//       u = (r, xb-1), v = (r', xb):   needed only if (r', xb-1) is not set (else u ~ (r', xb-1) inside the tile and
//                                      (r', xb-1) ~ v by the stitch)
//       u = (r, xb),   v = (r', xb-1): needed only if neither (r', xb) nor (r, xb-1) is set (same argument)
template <int TY, int TZ, bool FULL_CONN>
__global__ void __launch_bounds__(THREADS)
ccl_face_kernel(const unsigned char* __restrict__ mask, unsigned char want, Dims d, int* __restrict__ parent) {
    __shared__ int2 queue[THREADS / 32][64];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int2* q = queue[threadIdx.x >> 5];
    int n_q = 0;                                          // warp-uniform
    auto push = [&](bool has, int a, int b) {            // called by all lanes of the warp
        const unsigned bits = __ballot_sync(0xffffffffu, has);
        if (bits == 0u) return;
        if (has) q[n_q + __popc(bits & lt)] = make_int2(a, b);
        n_q += __popc(bits);
        if (n_q >= 32) {
            __syncwarp();
            const int2 pr = q[n_q - 32 + lane];
            n_q -= 32;
            unite(parent, pr.x, pr.y);
            __syncwarp();
        }
    };
    const int nrows = d.nz * d.ny;                        // < 2^31 (one voxel per row at least)
    const int wpr = (d.nx + 31) / 32;
    for (int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5); row < nrows;
         row += (int)(((long long)gridDim.x * blockDim.x) >> 5)) {
        const int z = row / d.ny, y = row - z * d.ny;
        const bool up_y = (y % TY) == 0 && y > 0, dn_y = (y % TY) == TY - 1 && y + 1 < d.ny, up_z = (z % TZ) == 0 && z > 0;
        if (!up_y && !up_z && !(FULL_CONN && dn_y && z > 0)) continue;
        for (int xw = 0; xw < wpr; ++xw) {
            const int x = xw * 32 + lane;
            const long long i = (long long)row * d.nx + x;
            const bool valid = x < d.nx && mask[i] == want;
            if (__ballot_sync(0xffffffffu, valid) == 0u) continue;
            const int me = (int)i;
            auto in = [&](int dz, int dy, int dx) -> bool {
                const int zz = z + dz, yy = y + dy, xx = x + dx;
                if (!valid || zz < 0 || yy < 0 || yy >= d.ny || xx < 0 || xx >= d.nx) return false;
                return mask[i + (long long)dz * d.plane + (long long)dy * d.nx + dx] == want;
            };
            auto off = [&](int dz, int dy, int dx) -> int {
                return (int)(i + (long long)dz * d.plane + (long long)dy * d.nx + dx);
            };
            const bool w_in = in(0, 0, -1);
            if (!FULL_CONN) {
                if (up_y) push(in(0, -1, 0) && !(w_in && in(0, -1, -1)), me, off(0, -1, 0));
                if (up_z) push(in(-1, 0, 0) && !(w_in && in(-1, 0, -1)), me, off(-1, 0, 0));
            } else {
#pragma unroll
                for (int dz = -1; dz <= 0; ++dz) {
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        if (dz == 0 && dy >= 0) continue;
                        const bool rc = (dz < 0 && up_z) || (dy < 0 && up_y) || (dy > 0 && dn_y);   // row in another tile
                        if (!rc) continue;
                        const bool c0 = in(dz, dy, 0), cp = in(dz, dy, 1), cm = in(dz, dy, -1);
                        push(valid && !c0 && cp, me, off(dz, dy, 1));
                        push(valid && !w_in && (c0 || cm), me, c0 ? off(dz, dy, 0) : off(dz, dy, -1));
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane < n_q) unite(parent, q[lane].x, q[lane].y);
}

template <int TY, int TZ, bool FULL_CONN>
__global__ void __launch_bounds__(THREADS)
ccl_seam_x_kernel(const unsigned char* __restrict__ mask, unsigned char want, Dims d, int* __restrict__ parent) {
    const int nb = (d.nx - 1) / 32;                       // strip boundaries per row: xb = 32, 64, ... < nx
    if (nb <= 0) return;
    const long long total = (long long)d.nz * d.ny * nb;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(k / nb);
        const int xb = ((int)(k - (long long)row * nb) + 1) * 32;
        const long long i = (long long)row * d.nx + xb;          // voxel (row, xb); i - 1 = (row, xb - 1)
        const bool e = mask[i] == want, w = mask[i - 1] == want;
        if (!e && !w) continue;
        if (e && w) unite(parent, (int)i, (int)i - 1);
        if (!FULL_CONN) continue;
        const int z = row / d.ny, y = row - z * d.ny;
        const bool up_y = (y % TY) == 0, dn_y = (y % TY) == TY - 1, up_z = (z % TZ) == 0;
#pragma unroll
        for (int dz = -1; dz <= 0; ++dz) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                if (dz == 0 && dy >= 0) continue;
                if ((dz < 0 && up_z) || (dy < 0 && up_y) || (dy > 0 && dn_y)) continue;   // ccl_face_kernel's pair
                if (z + dz < 0 || y + dy < 0 || y + dy >= d.ny) continue;
                const long long j = i + (long long)dz * d.plane + (long long)dy * d.nx;   // (r', xb)
                const bool ne = mask[j] == want, nw = mask[j - 1] == want;
                if (w && ne && !nw) unite(parent, (int)i - 1, (int)j);
                if (e && nw && !ne && !w) unite(parent, (int)i, (int)j - 1);
            }
        }
    }
}

// parent[i] = root of i for every voxel of the set.  root_bits (one word per 32-voxel strip row, written by the tile
// kernel) marks the voxels that were tile-local roots: only those can have been re-linked to another tile, so they
// walk first; after the CTA barrier every other voxel reaches a root through its tile root in one or two cached hops.
// Every store is a true root (no unions run concurrently), so concurrent walks of other CTAs stay valid.
template <int TY, int TZ>
__global__ void __launch_bounds__(THREADS)
ccl_flatten_tile_kernel(Dims d, TileGrid tg, const unsigned* __restrict__ root_bits, int* __restrict__ parent,
                        int* __restrict__ area) {
    constexpr int ROWS = TY * TZ;
    long long t = blockIdx.x;
    const int tx = (int)(t % tg.tiles_x);
    t /= tg.tiles_x;
    const int ty = (int)(t % tg.tiles_y), tz = (int)(t / tg.tiles_y);
    const int x0 = tx * 32, y0 = ty * TY, z0 = tz * TZ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = x0 + lane;
#pragma unroll
    for (int k = 0; k < ROWS / (THREADS / 32); ++k) {
        const int r = warp + k * (THREADS / 32);
        const int z = z0 + r / TY, y = y0 + r % TY;
        if (y >= d.ny || z >= d.nz) continue;
        const unsigned rb = __ldg(root_bits + ((long long)z * d.ny + y) * tg.tiles_x + tx);
        if (!((rb >> lane) & 1u)) continue;
        const long long i = (long long)z * d.plane + (long long)y * d.nx + x;
        const int p = ld_parent(parent, i);
        if (p >= 0 && p != (int)i) {
            const int f = find_root_ro(parent, p);
            parent[i] = f;
            // the size of a component is the sum over its tile-local pieces; area[i] of a non-root is final (nobody adds to it)
            if (area != nullptr && f >= 0) atomicAdd(area + f, area[i]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ROWS / (THREADS / 32); ++k) {
        const int r = warp + k * (THREADS / 32);
        const int z = z0 + r / TY, y = y0 + r % TY;
        if (x >= d.nx || y >= d.ny || z >= d.nz) continue;
        const long long i = (long long)z * d.plane + (long long)y * d.nx + x;
        const int p = ld_parent(parent, i);
        if (p < 0 || p == (int)i) continue;
        const int q = ld_parent(parent, p);
        if (q != p) parent[i] = find_root_ro(parent, q);       // q == p: p is a root and parent[i] is final already
    }
}

// ---- fill holes: mask |= background voxels whose tree is not tied to OUTSIDE ----------------------
__global__ void __launch_bounds__(THREADS)
fill_holes_kernel(Dims d, const int* __restrict__ parent, unsigned char* __restrict__ mask) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d.total;
         i += (long long)gridDim.x * blockDim.x) {
        const int p = parent[i];            // flattened: root index, OUTSIDE, or NOT_IN_SET (foreground)
        if (p >= 0) mask[i] = 1;
    }
}

// ---- component sizes -----------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
area_count_kernel(Dims d, const int* __restrict__ parent, int* __restrict__ area) {
    // consecutive voxels with the same root (x-runs, after the flatten) are counted once: the first lane of each group
    // adds the group's length
    for (long long base = blockIdx.x * (long long)blockDim.x; base < d.total; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        const int lane = threadIdx.x & 31;
        const int r = i < d.total ? parent[i] : NOT_IN_SET;
        const int prev = __shfl_up_sync(0xffffffffu, r, 1);
        const bool head = lane == 0 || prev != r;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        if (head && r >= 0) {
            const unsigned later = lane == 31 ? 0u : (heads >> (lane + 1));
            const int len = later ? __ffs(later) : 32 - lane;
            atomicAdd(area + r, len);
        }
    }
}

// keep = component has at least min_area voxels, written as one membership word per 32-voxel strip of a row
// (bit k of word [row][xw] = voxel x = 32 xw + k; bits beyond nx are 0)
__global__ void __launch_bounds__(THREADS)
area_keep_bits_kernel(Dims d, const int* __restrict__ parent, const int* __restrict__ area, long long min_area,
                      unsigned* __restrict__ keep_bits) {
    const int wpr = (d.nx + 31) / 32;
    const int nwin = d.nz * d.ny * wpr;                     // <= voxels < 2^31
    const int lane = threadIdx.x & 31;
    for (int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5); w < nwin;
         w += (int)(((long long)gridDim.x * blockDim.x) >> 5)) {
        const int row = w / wpr;
        const int x = (w - row * wpr) * 32 + lane;
        bool keep = false;
        if (x < d.nx) {
            const int r = parent[(long long)row * d.nx + x];
            keep = r >= 0 && (long long)area[r] >= min_area;
        }
        const unsigned bits = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) keep_bits[w] = bits;
    }
}

// ---- 3^d majority with reflected borders, bit-sliced ------------------------------------------------
// uniform_filter(float32(mask), 3, mode="reflect") > 0.5  <=>  (set voxels in the 3^d window, indices clamped) >= 14
// (5 in 2-D).  One THREAD per membership word: per window row the three horizontally shifted words are added as bit
// planes (a 2-bit sum per voxel) and accumulated into a 5-plane counter; 32 voxels cost ~15 logic instructions per row
// instead of 3 byte loads each (the per-voxel form took 1.5 ms of a 512^3 frame, a warp-per-strip form with shuffles 2.9).
__global__ void __launch_bounds__(THREADS)
majority_bits_kernel(const unsigned* __restrict__ bits, Dims d, unsigned char* __restrict__ out) {
    const int wpr = (d.nx + 31) / 32;
    const long long nwords = (long long)d.nz * d.ny * wpr;
    const bool three_d = d.nz > 1;
    for (long long wi = blockIdx.x * (long long)blockDim.x + threadIdx.x; wi < nwords; wi += (long long)gridDim.x * blockDim.x) {
        const long long row = wi / wpr;
        const int xw = (int)(wi - row * wpr);
        const int z = (int)(row / d.ny), y = (int)(row - (long long)z * d.ny);
        const int nvalid = min(32, d.nx - xw * 32);
        unsigned t0 = 0u, t1 = 0u, t2 = 0u, t3 = 0u, t4 = 0u;
        for (int dz = three_d ? -1 : 0; dz <= (three_d ? 1 : 0); ++dz) {
            const int zz = min(max(z + dz, 0), d.nz - 1);      // reflect of a 1-voxel overhang = clamp
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = min(max(y + dy, 0), d.ny - 1);
                const unsigned* rw = bits + ((long long)zz * d.ny + yy) * wpr;
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
        unsigned char* o = out + row * d.nx + (long long)xw * 32;
        if (nvalid == 32 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0u) {
            uint4 lo, hi;
            lo.x = ((res >> 0) & 15u) * 0x00204081u & 0x01010101u;
            lo.y = ((res >> 4) & 15u) * 0x00204081u & 0x01010101u;
            lo.z = ((res >> 8) & 15u) * 0x00204081u & 0x01010101u;
            lo.w = ((res >> 12) & 15u) * 0x00204081u & 0x01010101u;
            hi.x = ((res >> 16) & 15u) * 0x00204081u & 0x01010101u;
            hi.y = ((res >> 20) & 15u) * 0x00204081u & 0x01010101u;
            hi.z = ((res >> 24) & 15u) * 0x00204081u & 0x01010101u;
            hi.w = ((res >> 28) & 15u) * 0x00204081u & 0x01010101u;
            reinterpret_cast<uint4*>(o)[0] = lo;
            reinterpret_cast<uint4*>(o)[1] = hi;
        } else {
            for (int k = 0; k < nvalid; ++k) o[k] = (res >> k) & 1u;
        }
    }
}

// ---- raster-order numbering of the roots -------------------------------------------------------------
constexpr int RANK_CHUNK = 2048;    // voxels per block in the counting / assigning kernels

__global__ void __launch_bounds__(THREADS)
root_count_kernel(Dims d, const int* __restrict__ parent, int* __restrict__ block_counts) {
    __shared__ int warp_sums[THREADS / 32];
    const long long base = (long long)blockIdx.x * RANK_CHUNK;
    int c = 0;
    for (int k = threadIdx.x; k < RANK_CHUNK; k += THREADS) {
        const long long i = base + k;
        if (i < d.total && parent[i] == (int)i) ++c;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < THREADS / 32; ++k) s += warp_sums[k];
        block_counts[blockIdx.x] = s;
    }
}

// single CTA: exclusive scan of block_counts in place; total -> *n_labels
__global__ void __launch_bounds__(1024)
block_scan_kernel(int* __restrict__ block_counts, long long nblocks, long long* __restrict__ n_labels) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long i = base + threadIdx.x;
        const int v = i < nblocks ? block_counts[i] : 0;
        int incl = v;
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
        const int before = (w ? warp_tot[w - 1] : 0) + carry;
        if (i < nblocks) block_counts[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_labels = (long long)carry_s;
}

__global__ void __launch_bounds__(THREADS)
root_assign_kernel(Dims d, const int* __restrict__ parent, const int* __restrict__ block_offsets,
                   int* __restrict__ labels) {
    // ranks inside a chunk follow raster order: the chunk is scanned in THREADS-wide strips
    __shared__ int warp_sums[THREADS / 32];
    __shared__ int running;
    if (threadIdx.x == 0) running = block_offsets[blockIdx.x];
    __syncthreads();
    const long long base = (long long)blockIdx.x * RANK_CHUNK;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k0 = 0; k0 < RANK_CHUNK; k0 += THREADS) {
        const long long i = base + k0 + threadIdx.x;
        const bool is_root = i < d.total && parent[i] == (int)i;
        const unsigned bits = __ballot_sync(0xffffffffu, is_root);
        const int before_in_warp = __popc(bits & ((1u << lane) - 1u));
        if (lane == 0) warp_sums[w] = __popc(bits);
        __syncthreads();
        int before = running;
        for (int k = 0; k < w; ++k) before += warp_sums[k];
        if (is_root) labels[i] = before + before_in_warp + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int k = 0; k < THREADS / 32; ++k) s += warp_sums[k];
            running += s;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(THREADS)
label_propagate_kernel(Dims d, const int* __restrict__ parent, int* __restrict__ labels) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d.total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = parent[i];
        if (r < 0) labels[i] = 0;
        else if (r != (int)i) labels[i] = labels[r];
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

long long max_ll(long long a, long long b) { return a > b ? a : b; }

int run_ccl_legacy(const unsigned char* mask, unsigned char want, const Dims& d, bool full_conn, bool border_outside,
                   int* parent, cudaStream_t st) {
    ccl_init_kernel<<<gs(d.total), THREADS, 0, st>>>(mask, want, d, parent);
    if (full_conn && !border_outside) ccl_merge_kernel<true, false><<<gs(d.total), THREADS, 0, st>>>(mask, want, d, parent);
    else if (!full_conn && border_outside) ccl_merge_kernel<false, true><<<gs(d.total), THREADS, 0, st>>>(mask, want, d, parent);
    else if (!full_conn) ccl_merge_kernel<false, false><<<gs(d.total), THREADS, 0, st>>>(mask, want, d, parent);
    else ccl_merge_kernel<true, true><<<gs(d.total), THREADS, 0, st>>>(mask, want, d, parent);
    ccl_flatten_kernel<<<gs(d.total), THREADS, 0, st>>>(d, parent);
    return nb::check_launch("ccl");
}

template <int TY, int TZ>
int run_ccl_tiled(const unsigned char* mask, unsigned char want, const Dims& d, bool full_conn, bool border_outside,
                  int* parent, unsigned* root_bits, int* area, cudaStream_t st) {
    static const int stage = [] { const char* e = getenv("NB200_CCL_STAGE"); return e ? atoi(e) : 9; }();   // timing aid
    TileGrid tg;
    tg.tiles_x = (d.nx + 31) / 32;
    tg.tiles_y = (d.ny + TY - 1) / TY;
    const long long tiles = (long long)tg.tiles_x * tg.tiles_y * ((d.nz + TZ - 1) / TZ);
    NB_REQUIRE(tiles < 2147483647LL, NB200_ERR_UNSUPPORTED, "ccl: too many tiles");
    const unsigned g = (unsigned)tiles;
    if (full_conn && !border_outside) ccl_tile_kernel<TY, TZ, true, false><<<g, THREADS, 0, st>>>(mask, want, d, tg, parent, root_bits, area);
    else if (!full_conn && border_outside) ccl_tile_kernel<TY, TZ, false, true><<<g, THREADS, 0, st>>>(mask, want, d, tg, parent, root_bits, area);
    else if (!full_conn) ccl_tile_kernel<TY, TZ, false, false><<<g, THREADS, 0, st>>>(mask, want, d, tg, parent, root_bits, area);
    else ccl_tile_kernel<TY, TZ, true, true><<<g, THREADS, 0, st>>>(mask, want, d, tg, parent, root_bits, area);
    if (stage < 2) return nb::check_launch("ccl(tiled)");
    const long long rows = (long long)d.nz * d.ny;
    const long long seams = rows * ((d.nx - 1) / 32);
    if (full_conn) {
        if (seams > 0) ccl_seam_x_kernel<TY, TZ, true><<<gs(seams), THREADS, 0, st>>>(mask, want, d, parent);
        if (stage >= 3) ccl_face_kernel<TY, TZ, true><<<gs(rows * 32), THREADS, 0, st>>>(mask, want, d, parent);
    } else {
        if (seams > 0) ccl_seam_x_kernel<TY, TZ, false><<<gs(seams), THREADS, 0, st>>>(mask, want, d, parent);
        if (stage >= 3) ccl_face_kernel<TY, TZ, false><<<gs(rows * 32), THREADS, 0, st>>>(mask, want, d, parent);
    }
    if (stage < 4) return nb::check_launch("ccl(tiled)");
    ccl_flatten_tile_kernel<TY, TZ><<<g, THREADS, 0, st>>>(d, tg, root_bits, parent, area);
    return nb::check_launch("ccl(tiled)");
}

// area (optional): on return area[r] = voxels of the component with root r (other entries are scratch)
int run_ccl(const unsigned char* mask, unsigned char want, const Dims& d, bool full_conn, bool border_outside,
            int* parent, unsigned* root_bits, int* area, cudaStream_t st) {
    static const int legacy = [] { const char* e = getenv("NB200_CCL_LEGACY"); return e ? atoi(e) : 0; }();
    if (legacy) {
        const int rc = run_ccl_legacy(mask, want, d, full_conn, border_outside, parent, st);
        if (rc || area == nullptr) return rc;
        cudaMemsetAsync(area, 0, sizeof(int) * d.total, st);
        area_count_kernel<<<gs(d.total), THREADS, 0, st>>>(d, parent, area);
        return nb::check_launch("area_count");
    }
    if (d.nz > 1) return run_ccl_tiled<8, 8>(mask, want, d, full_conn, border_outside, parent, root_bits, area, st);
    return run_ccl_tiled<64, 1>(mask, want, d, full_conn, border_outside, parent, root_bits, area, st);
}

}  // namespace

extern "C" {

size_t nb200_label_workspace_bytes(int nz, int ny, int nx) {
    const long long n = (long long)nz * ny * nx;
    const long long nblocks = (n + RANK_CHUNK - 1) / RANK_CHUNK;
    // parent int32[n] | mask_a u8[n] | keep words u32[rows * ceil(nx/32)] (at least n bytes) | block counts int32[nblocks]
    // (each 256-byte aligned)
    auto al = [](long long b) { return (b + 255) / 256 * 256; };
    const long long words = (long long)nz * ny * ((nx + 31) / 32);
    return (size_t)(al(4 * n) + al(n) + al(max_ll(n, 4 * words)) + al(4 * nblocks));
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
    auto al = [](long long b) { return (b + 255) / 256 * 256; };
    char* ws = static_cast<char*>(workspace);
    int* parent = reinterpret_cast<int*>(ws);
    unsigned char* mask_a = reinterpret_cast<unsigned char*>(ws + al(4 * n));
    const long long words = (long long)nz * ny * ((nx + 31) / 32);
    unsigned* keep_bits = reinterpret_cast<unsigned*>(mask_a + al(n));
    int* block_counts = reinterpret_cast<int*>(mask_a + al(n) + al(max_ll(n, 4 * words)));
    const long long nblocks = (n + RANK_CHUNK - 1) / RANK_CHUNK;
    int rc;

    threshold_mask_kernel<<<gs(n), THREADS, 0, st>>>(frangi, raw, use_intensity, intensity_thresh, thr, n, mask_a);
    if (fill_holes && nz > 1) {   // labelling.py:485-486 (3-D only)
        rc = run_ccl(mask_a, 0, d, /*full_conn=*/false, /*border_outside=*/true, parent, keep_bits, nullptr, st);
        if (rc) return rc;
        fill_holes_kernel<<<gs(n), THREADS, 0, st>>>(d, parent, mask_a);
    }
    // first labelling + size filter (labelling.py:489-501)
    rc = run_ccl(mask_a, 1, d, true, false, parent, keep_bits, /*area=*/labels, st);
    if (rc) return rc;
    area_keep_bits_kernel<<<gs(n), THREADS, 0, st>>>(d, parent, labels, min_area, keep_bits);
    // smoothing (labelling.py:503-505) and second labelling (:507)
    majority_bits_kernel<<<gs(words), THREADS, 0, st>>>(keep_bits, d, mask_a);
    rc = run_ccl(mask_a, 1, d, true, false, parent, keep_bits, nullptr, st);
    if (rc) return rc;
    root_count_kernel<<<(unsigned)nblocks, THREADS, 0, st>>>(d, parent, block_counts);
    block_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nblocks, n_labels);
    root_assign_kernel<<<(unsigned)nblocks, THREADS, 0, st>>>(d, parent, block_counts, labels);
    label_propagate_kernel<<<gs(n), THREADS, 0, st>>>(d, parent, labels);
    return nb::check_launch("label_frame");
}

/* scipy.ndimage.label(mask, structure=ones) alone (used by tests and by the Network stage later):
 * mask: uint8 device array; connectivity_full: 1 = 26/8, 0 = 6/4. */
int nb200_ccl_label(const unsigned char* mask, int nz, int ny, int nx, int connectivity_full, int* labels,
                    void* workspace, long long* n_labels, void* stream) {
    NB_REQUIRE(mask && labels && workspace && n_labels, NB200_ERR_ARG, "nb200_ccl_label: null argument");
    const long long n = (long long)nz * ny * nx;
    NB_REQUIRE(n < 2147483000LL, NB200_ERR_UNSUPPORTED, "nb200_ccl_label: frame exceeds int32 voxel indexing");
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    auto al = [](long long b) { return (b + 255) / 256 * 256; };
    char* ws = static_cast<char*>(workspace);
    int* parent = reinterpret_cast<int*>(ws);
    const long long words = (long long)nz * ny * ((nx + 31) / 32);
    unsigned* root_bits = reinterpret_cast<unsigned*>(ws + al(4 * n) + al(n));
    int* block_counts = reinterpret_cast<int*>(ws + al(4 * n) + al(n) + al(max_ll(n, 4 * words)));
    const long long nblocks = (n + RANK_CHUNK - 1) / RANK_CHUNK;
    int rc = run_ccl(mask, 1, d, connectivity_full != 0, false, parent, root_bits, nullptr, st);
    if (rc) return rc;
    root_count_kernel<<<(unsigned)nblocks, THREADS, 0, st>>>(d, parent, block_counts);
    block_scan_kernel<<<1, 1024, 0, st>>>(block_counts, nblocks, n_labels);
    root_assign_kernel<<<(unsigned)nblocks, THREADS, 0, st>>>(d, parent, block_counts, labels);
    label_propagate_kernel<<<gs(n), THREADS, 0, st>>>(d, parent, labels);
    return nb::check_launch("ccl_label");
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

__global__ void __launch_bounds__(THREADS)
non_junction_mask_kernel(const unsigned char* __restrict__ pixel_class, long long n, unsigned char* __restrict__ mask) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned char c = pixel_class[i];
        mask[i] = (c > 0 && c != 4) ? 1 : 0;
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
    const long long n = (long long)nz * ny * nx;
    NB_REQUIRE(n < 2147483000LL, NB200_ERR_UNSUPPORTED, "nb200_branch_labels: frame exceeds int32 voxel indexing");
    auto al = [](long long b) { return (b + 255) / 256 * 256; };
    unsigned char* mask = reinterpret_cast<unsigned char*>(static_cast<char*>(workspace) + al(4 * n));   // mask_a of the label workspace
    non_junction_mask_kernel<<<gs(n), THREADS, 0, nb::as_stream(stream)>>>(pixel_class, n, mask);
    int rc = nb::check_launch("branch_labels(mask)");
    if (rc) return rc;
    return nb200_ccl_label(mask, nz, ny, nx, 1, labels, workspace, n_labels, stream);
}

int nb200_remove_connected_label_pixels(const int* labels, int nz, int ny, int nx, int* out, void* stream) {
    NB_REQUIRE(labels && out && labels != out && nz >= 1 && ny >= 1 && nx >= 1, NB200_ERR_ARG,
               "nb200_remove_connected_label_pixels: bad argument");
    const Dims d = make_dims(nz, ny, nx);
    remove_connected_kernel<<<gs(d.total), THREADS, 0, nb::as_stream(stream)>>>(labels, d, out);
    return nb::check_launch("remove_connected_label_pixels");
}

}  // extern "C"
