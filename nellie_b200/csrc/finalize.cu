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
    const long long nbx = (v.nx + TX - 1) / TX, nby = (v.ny + TY - 1) / TY, nbz = (v.zc1 - v.zc0 + TZ - 1) / TZ;
    opening_kernel<<<(unsigned)(nbx * nby * nbz), NTHREADS, 0, nb::as_stream(stream)>>>(acc, out, v, thr);
    return nb::check_launch("finalize_opening");
}

int nb200_finalize_opening_2d(const float* vin, float* out, int ny, int nx, const double* thr, void* stream) {
    NB_REQUIRE(vin && out && thr && ny > 0 && nx > 0 && vin != out, NB200_ERR_ARG, "nb200_finalize_opening_2d: bad argument");
    opening_2d_kernel<<<nb::grid_for((long long)ny * nx, 256, 4), 256, 0, nb::as_stream(stream)>>>(vin, out, ny, nx, thr);
    return nb::check_launch("finalize_opening_2d");
}

}  // extern "C"
