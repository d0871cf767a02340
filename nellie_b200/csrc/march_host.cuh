// Host-side pieces shared by the translation units that launch Z-marching Hessian kernels
// (frangi.cu: exact march + border shell; hessian_fast.cu: approximate-classify march).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace nb {

struct MarchPlan {
    int zi0, zi1, zchunk;     // interior planes (GLOBAL coordinates) and chunk length
    long long n_ctas;
};
// Interior of the compute window; Z chunks sized for several waves over the SMs, no shorter than 32 planes.
// Tiles are hm::TX x hm::TYO (128 x 14) columns.
MarchPlan plan_march(const nb200_vol& v);

// TMA descriptor of a (nz_buf, ny, nx) float32 volume, box = one staged plane tile (136 x 18 floats).
// Returns false when the volume cannot be described (nx % 4 != 0, unaligned base, driver entry point missing).
bool make_plane_map(const float* g, const nb200_vol& v, CUtensorMap* map);

int check_vol(const nb200_vol& v, const char* who);

// Border shell (one-sided differences of numpy.gradient, IEEE division), exact per-voxel kernels of frangi.cu.
// stats: max|H|, max frob_sq, value range, sqrt(frob_sq) at the lattice points of the shell.
// gate: -1 always; 1 = only when sp[UNSAFE]; 2 = only when sp[AMBIG] and not sp[UNSAFE].
int launch_shell_stats(const float* g, const nb200_vol& v, const float* spacing, int sz, int sy, int sx,
                       float* frob_samples, long long* hstats, float* code, const double* sp, int gate,
                       cudaStream_t st);
// frangi: mask + eigenvalues + vesselness + max/AND for the shell voxels (returns at once when sp[SKIP]).
int launch_shell_frangi(const float* g, float* acc, const nb200_vol& v, const float* spacing, float alpha_sq,
                        float beta_sq, const double* sp, cudaStream_t st);

}  // namespace nb
