// Network stage, the two array steps that the reference keeps on the host even in its GPU backend:
//
//   networking.py:315-392  _add_missing_skeleton_labels: every object without a skeleton voxel gets one at the position of
//                          its largest Frangi response (scipy.ndimage.maximum_position)
//   networking.py:485-577  _relabel_objects: every voxel of an object takes the branch label of the nearest skeleton voxel of
//                          the SAME object — per object a scipy.ndimage.distance_transform_edt(sampling=scaling,
//                          return_indices=True) on the object's bounding-box crop
//
// (pixel class, branch labelling and the label clean-up of the same stage are in label.cu.)
//
// _relabel_objects must reproduce not only the distances but scipy's choice among EQUIDISTANT seeds (they may carry different
// branch labels next to a junction).  scipy's feature transform (scipy/ndimage/src/ni_morphology.c: _ComputeFT / _VoronoiFT,
// the dimension-by-dimension Voronoi construction of Maurer et al., not part of the reference checkout) is therefore
// restated step for step, with its float64 arithmetic in its order: stage d = for every line along axis d, build the lower
// envelope of the candidate seeds of the line's voxels (partial feature transform of the previous stages), dropping a
// candidate when  c*vR - b*uR - a*wR - a*b*c > 0,  then walk the line and advance to the next candidate while it is
// STRICTLY closer.  tests/test_network_cpu.py checks this restatement against scipy itself on random crops full of ties.
//
// All objects are processed at once: the bounding-box crops are laid out one after the other in a "crop space" of
// V = sum of box volumes voxels (object of a crop voxel by binary search over the offsets), one thread per crop voxel for
// the initialisation and the final gather, one thread per crop line for each Voronoi stage (ping-pong between two feature
// buffers; the candidate stack of a line lives in the line's own voxels of a third buffer).  Barrier-free; the atomics
// (bounding boxes, per-object arg-max) go through nb_atomic_* so the file also compiles for the host (oracle/cuda_emu.h).
#ifdef NB200_HOST_EMU
#include NB200_HOST_EMU
#else
#include "common.cuh"
#define NB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define nb_atomic_min_i32(p, v) atomicMin((p), (v))
#define nb_atomic_max_i32(p, v) atomicMax((p), (v))
#define nb_atomic_max_u64(p, v) atomicMax((p), (v))
#endif
#include "devmath.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int CTAS_PER_SM = 8;

struct Dims {
    int nz, ny, nx;
    long long plane, total;
};

inline Dims make_dims(int nz, int ny, int nx) {
    Dims d;
    d.nz = nz; d.ny = ny; d.nx = nx;
    d.plane = (long long)ny * nx;
    d.total = d.plane * nz;
    return d;
}

inline unsigned grid_of(long long n) { return nb::grid_for(n, THREADS, CTAS_PER_SM); }

__device__ __forceinline__ unsigned ordered_of(float f) {
    const unsigned u = nb::f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- _add_missing_skeleton_labels -----------------------------------------------------------------------------------------
// key[lab] = max over the object's voxels of (ordered Frangi value << 32 | ~index): largest value, first voxel in raster
// order among equal values; in_skel[lab] = 1 when the skeleton already holds the label.
__global__ void __launch_bounds__(THREADS)
object_argmax_kernel(const int* __restrict__ labels, const int* __restrict__ skel, const float* __restrict__ frangi,
                     long long n, int max_label, unsigned long long* key, unsigned char* in_skel) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int s = skel[idx];
        if (s > 0 && s <= max_label) in_skel[s] = 1;
        const int lab = labels[idx];
        if (lab > 0 && lab <= max_label) {
            const unsigned long long k = ((unsigned long long)ordered_of(frangi[idx]) << 32) |
                                         (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
            nb_atomic_max_u64(&key[lab], k);
        }
    }
}

__global__ void __launch_bounds__(THREADS)
add_missing_kernel(const unsigned long long* __restrict__ key, const unsigned char* __restrict__ in_skel, int max_label,
                   int* skel) {
    for (int lab = 1 + blockIdx.x * blockDim.x + threadIdx.x; lab <= max_label; lab += gridDim.x * blockDim.x) {
        const unsigned long long k = key[lab];
        if (k != 0ull && !in_skel[lab]) skel[0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)] = lab;
    }
}

// skel_pre = (skel > 0) * labels   (networking.py:835)
__global__ void __launch_bounds__(THREADS)
skeleton_labels_kernel(const int* __restrict__ skel, const int* __restrict__ labels, long long n, int* __restrict__ out) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x)
        out[idx] = skel[idx] > 0 ? labels[idx] : 0;
}

// ---- _relabel_objects -------------------------------------------------------------------------------------------------------
// boxes: int32 (max_label + 1, 6) = min z, y, x (initialised to INT_MAX) and max z, y, x (initialised to -1);
// seeded[lab] = 1 when the object holds a branch-labelled voxel.
__global__ void __launch_bounds__(THREADS)
object_boxes_kernel(const int* __restrict__ labels, const int* __restrict__ branch, Dims d, int max_label, int* boxes,
                    unsigned char* seeded) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < d.total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int lab = labels[idx];
        if (lab <= 0 || lab > max_label) continue;
        const int z = (int)(idx / d.plane);
        const long long rem = idx - (long long)z * d.plane;
        const int y = (int)(rem / d.nx);
        const int x = (int)(rem - (long long)y * d.nx);
        int* b = boxes + 6ll * lab;
        if (z < b[0]) nb_atomic_min_i32(b + 0, z);
        if (y < b[1]) nb_atomic_min_i32(b + 1, y);
        if (x < b[2]) nb_atomic_min_i32(b + 2, x);
        if (z > b[3]) nb_atomic_max_i32(b + 3, z);
        if (y > b[4]) nb_atomic_max_i32(b + 4, y);
        if (x > b[5]) nb_atomic_max_i32(b + 5, x);
        if (branch[idx] > 0) seeded[lab] = 1;
    }
}

// crop table, one row per object that is relabelled (rows in ascending label order):
//   crops int64 (m, 8) = label, z0, y0, x0, ez, ey, ex, offset of the crop in crop space
struct Crop {
    int label, lo[3], ext[3];
    long long off;
};

__device__ __forceinline__ Crop crop_of(const long long* __restrict__ crops, long long k) {
    Crop c;
    const long long* r = crops + 8 * k;
    c.label = (int)r[0];
    for (int a = 0; a < 3; ++a) { c.lo[a] = (int)r[1 + a]; c.ext[a] = (int)r[4 + a]; }
    c.off = r[7];
    return c;
}

// index of the crop that holds position q of a prefix-summed space: largest k with start(k) <= q
__device__ __forceinline__ long long find_crop(const long long* __restrict__ starts, long long stride, long long m,
                                               long long q) {
    long long lo = 0, hi = m - 1;
    while (lo < hi) {
        const long long mid = (lo + hi + 1) >> 1;
        if (starts[mid * stride] <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// features of the seeds: their own crop coordinates; -1 elsewhere (ni_morphology.c _ComputeFT, d == 0 initialisation)
__global__ void __launch_bounds__(THREADS)
ft_init_kernel(const int* __restrict__ labels, const int* __restrict__ branch, Dims d, const long long* __restrict__ crops,
               long long m, long long V, int* __restrict__ ft) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < V; q += (long long)gridDim.x * blockDim.x) {
        const Crop c = crop_of(crops, find_crop(crops + 7, 8, m, q));
        const long long r = q - c.off;
        const long long yx = (long long)c.ext[1] * c.ext[2];
        const int cz = (int)(r / yx);
        const int cy = (int)((r - cz * yx) / c.ext[2]);
        const int cx = (int)(r - cz * yx - (long long)cy * c.ext[2]);
        const long long idx = (long long)(c.lo[0] + cz) * d.plane + (long long)(c.lo[1] + cy) * d.nx + (c.lo[2] + cx);
        const bool seed = labels[idx] == c.label && branch[idx] > 0;
        ft[q] = seed ? cz : -1;
        ft[V + q] = seed ? cy : -1;
        ft[2 * V + q] = seed ? cx : -1;
    }
}

// one Voronoi stage: lines along axis `axis` (0 = Z, 1 = Y, 2 = X) of every crop; line_starts int64 (m) = exclusive prefix
// sum of the number of such lines per crop.  in / out: feature buffers (3, V); stack: int32 (V).
__global__ void __launch_bounds__(THREADS)
voronoi_stage_kernel(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ stack,
                     const long long* __restrict__ crops, const long long* __restrict__ line_starts, long long m,
                     long long n_lines, long long V, int axis, double s0, double s1, double s2) {
    const double sampling[3] = {s0, s1, s2};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_lines;
         t += (long long)gridDim.x * blockDim.x) {
        const long long k = find_crop(line_starts, 1, m, t);
        const Crop c = crop_of(crops, k);
        const long long li = t - line_starts[k];
        // the two coordinates that are fixed along the line, fastest-varying last
        const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
        int coor[3];
        coor[axis] = 0;
        coor[a1] = (int)(li / c.ext[a2]);
        coor[a2] = (int)(li - (long long)coor[a1] * c.ext[a2]);
        const long long stride[3] = {(long long)c.ext[1] * c.ext[2], (long long)c.ext[2], 1};
        const long long base = c.off + coor[0] * stride[0] + coor[1] * stride[1] + coor[2] * stride[2];
        const long long st = stride[axis];
        const int len = c.ext[axis];
        const int* f0 = in;
        const int* f1 = in + V;
        const int* f2 = in + 2 * V;
#define NB_F(ii, jj) ((jj) == 0 ? f0[base + (ii) * st] : (jj) == 1 ? f1[base + (ii) * st] : f2[base + (ii) * st])
#define NB_G(l) stack[base + (l) * st]
        int l = -1;
        for (int ii = 0; ii < len; ++ii) {
            if (NB_F(ii, 0) < 0) continue;
            const double fd = (double)NB_F(ii, axis);
            double wR = 0.0;
            for (int jj = 0; jj < 3; ++jj)
                if (jj != axis) {
                    double tw = (double)(NB_F(ii, jj) - coor[jj]);
                    tw *= sampling[jj];
                    wR += tw * tw;
                }
            while (l >= 1) {
                const int idx1 = NB_G(l), idx2 = NB_G(l - 1);
                const double fv = (double)NB_F(idx1, axis);
                double a = fv - (double)NB_F(idx2, axis);
                double b = fd - fv;
                a *= sampling[axis];
                b *= sampling[axis];
                const double cc = a + b;
                double uR = 0.0, vR = 0.0;
                for (int jj = 0; jj < 3; ++jj)
                    if (jj != axis) {
                        const double co = (double)coor[jj];
                        double tu = (double)NB_F(idx2, jj) - co;
                        double tv = (double)NB_F(idx1, jj) - co;
                        tu *= sampling[jj];
                        tv *= sampling[jj];
                        uR += tu * tu;
                        vR += tv * tv;
                    }
                if (cc * vR - b * uR - a * wR - a * b * cc <= 0.0) break;
                --l;
            }
            ++l;
            NB_G(l) = ii;
        }
        const int maxl = l;
        if (maxl < 0) {
            for (int ii = 0; ii < len; ++ii) {
                out[base + ii * st] = NB_F(ii, 0);
                out[V + base + ii * st] = NB_F(ii, 1);
                out[2 * V + base + ii * st] = NB_F(ii, 2);
            }
            continue;
        }
        l = 0;
        for (int ii = 0; ii < len; ++ii) {
            double delta1 = 0.0;
            for (int jj = 0; jj < 3; ++jj) {
                double tt = jj == axis ? (double)(NB_F(NB_G(l), jj) - ii) : (double)(NB_F(NB_G(l), jj) - coor[jj]);
                tt *= sampling[jj];
                delta1 += tt * tt;
            }
            while (l < maxl) {
                double delta2 = 0.0;
                for (int jj = 0; jj < 3; ++jj) {
                    double tt = jj == axis ? (double)(NB_F(NB_G(l + 1), jj) - ii) : (double)(NB_F(NB_G(l + 1), jj) - coor[jj]);
                    tt *= sampling[jj];
                    delta2 += tt * tt;
                }
                if (delta1 <= delta2) break;
                delta1 = delta2;
                ++l;
            }
            const int g = NB_G(l);
            out[base + ii * st] = NB_F(g, 0);
            out[V + base + ii * st] = NB_F(g, 1);
            out[2 * V + base + ii * st] = NB_F(g, 2);
        }
#undef NB_F
#undef NB_G
    }
}

// every voxel of the crop's own object takes the branch label of its nearest seed (networking.py:565-575)
__global__ void __launch_bounds__(THREADS)
relabel_gather_kernel(const int* __restrict__ labels, const int* __restrict__ branch, Dims d,
                      const long long* __restrict__ crops, long long m, long long V, const int* __restrict__ ft,
                      unsigned int* __restrict__ out) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < V; q += (long long)gridDim.x * blockDim.x) {
        const Crop c = crop_of(crops, find_crop(crops + 7, 8, m, q));
        const long long r = q - c.off;
        const long long yx = (long long)c.ext[1] * c.ext[2];
        const int cz = (int)(r / yx);
        const int cy = (int)((r - cz * yx) / c.ext[2]);
        const int cx = (int)(r - cz * yx - (long long)cy * c.ext[2]);
        const long long idx = (long long)(c.lo[0] + cz) * d.plane + (long long)(c.lo[1] + cy) * d.nx + (c.lo[2] + cx);
        if (labels[idx] != c.label) continue;
        const int sz = ft[q], sy = ft[V + q], sx = ft[2 * V + q];
        if (sz < 0) continue;
        const long long sidx = (long long)(c.lo[0] + sz) * d.plane + (long long)(c.lo[1] + sy) * d.nx + (c.lo[2] + sx);
        out[idx] = (unsigned int)branch[sidx];
    }
}

inline bool dims_ok(int nz, int ny, int nx) {
    return nz >= 1 && ny >= 1 && nx >= 1 && (long long)nz * ny * nx < (1ll << 32);
}

}  // namespace

extern "C" {

int nb200_network_add_missing(const int* labels, const float* frangi, int* skel, int nz, int ny, int nx, int max_label,
                              unsigned long long* key, unsigned char* in_skel, void* stream) {
    NB_REQUIRE(labels && frangi && skel && key && in_skel && max_label >= 0, NB200_ERR_ARG,
               "nb200_network_add_missing: bad argument");
    NB_REQUIRE(dims_ok(nz, ny, nx), NB200_ERR_UNSUPPORTED, "nb200_network_add_missing: frames of 2^32 voxels or more");
    if (max_label == 0) return NB200_OK;
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    NB_LAUNCH(object_argmax_kernel, grid_of(d.total), THREADS, st, labels, (const int*)skel, frangi, d.total, max_label, key,
              in_skel);
    NB_LAUNCH(add_missing_kernel, grid_of(max_label), THREADS, st, (const unsigned long long*)key,
              (const unsigned char*)in_skel, max_label, skel);
    return nb::check_launch("add-missing kernels");
}

int nb200_network_skeleton_labels(const int* skel, const int* labels, long long n, int* out, void* stream) {
    NB_REQUIRE(skel && labels && out && n >= 0, NB200_ERR_ARG, "nb200_network_skeleton_labels: bad argument");
    if (n == 0) return NB200_OK;
    NB_LAUNCH(skeleton_labels_kernel, grid_of(n), THREADS, nb::as_stream(stream), skel, labels, n, out);
    return nb::check_launch("skeleton_labels_kernel");
}

int nb200_network_object_boxes(const int* labels, const int* branch, int nz, int ny, int nx, int max_label, int* boxes,
                               unsigned char* seeded, void* stream) {
    NB_REQUIRE(labels && branch && boxes && seeded && max_label >= 0, NB200_ERR_ARG,
               "nb200_network_object_boxes: bad argument");
    NB_REQUIRE(dims_ok(nz, ny, nx), NB200_ERR_UNSUPPORTED, "nb200_network_object_boxes: frames of 2^32 voxels or more");
    const Dims d = make_dims(nz, ny, nx);
    NB_LAUNCH(object_boxes_kernel, grid_of(d.total), THREADS, nb::as_stream(stream), labels, branch, d, max_label, boxes, seeded);
    return nb::check_launch("object_boxes_kernel");
}

int nb200_network_relabel(const int* labels, const int* branch, int nz, int ny, int nx, const long long* crops, long long m,
                          long long crop_voxels, const long long* line_starts, const long long* n_lines,
                          const double* sampling, int* ft_a, int* ft_b, int* stack, unsigned int* out, void* stream) {
    NB_REQUIRE(labels && branch && out && m >= 0 && crop_voxels >= 0 && sampling && n_lines, NB200_ERR_ARG,
               "nb200_network_relabel: bad argument");
    NB_REQUIRE(dims_ok(nz, ny, nx), NB200_ERR_UNSUPPORTED, "nb200_network_relabel: frames of 2^32 voxels or more");
    if (m == 0 || crop_voxels == 0) return NB200_OK;
    NB_REQUIRE(crops && line_starts && ft_a && ft_b && stack, NB200_ERR_ARG, "nb200_network_relabel: null workspace");
    const Dims d = make_dims(nz, ny, nx);
    cudaStream_t st = nb::as_stream(stream);
    const long long V = crop_voxels;
    NB_LAUNCH(ft_init_kernel, grid_of(V), THREADS, st, labels, branch, d, crops, m, V, ft_a);
    int* src = ft_a;
    int* dst = ft_b;
    for (int axis = 0; axis < 3; ++axis) {
        NB_LAUNCH(voronoi_stage_kernel, grid_of(n_lines[axis]), THREADS, st, (const int*)src, dst, stack, crops,
                  line_starts + axis * m, m, n_lines[axis], V, axis, sampling[0], sampling[1], sampling[2]);
        int* t = src; src = dst; dst = t;
    }
    NB_LAUNCH(relabel_gather_kernel, grid_of(V), THREADS, st, labels, branch, d, crops, m, V, (const int*)src, out);
    return nb::check_launch("relabel kernels");
}

}  // extern "C"
