// Host build of nellie_b200/csrc/markers.cu through oracle/cuda_emu.h (TEST INFRASTRUCTURE ONLY): exports the same
// nb200_markers_* entry points, executing the same kernel bodies serially, plus a plain restatement of one axis of
// scipy.ndimage.correlate1d (symmetric taps, mode="reflect") standing in for nb200_gauss_axis / nb200_gauss_yx, which are
// tiled CUDA kernels checked against scipy on the GPU (tests/test_kernels_gpu.py).  Built by __graft_entry__.build() with
//   g++ -O2 -ffp-contract=off -shared -fPIC -DNB200_HOST_EMU='"<repo>/oracle/cuda_emu.h"' -x c++ oracle/markers_host.cpp
// Never loaded by nellie_b200.
#include "../nellie_b200/csrc/markers.cu"

namespace {
inline int reflect_index(int i, int n) {
    if (i >= 0 && i < n) return i;
    if (n == 1) return 0;
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}

int emu_axis(const float* src, float* dst, const nb200_vol* v, int axis, const double* w, int radius) {
    // scipy NI_Correlate1D, symmetric branch: tmp = x[l] * w[0]; for j = r..1: tmp += (x[l-j] + x[l+j]) * w[j]; float32 store
    const long long plane = (long long)v->ny * v->nx;
    const int n_axis = axis == 0 ? v->nz_buf : axis == 1 ? v->ny : v->nx;
    const long long stride = axis == 0 ? plane : axis == 1 ? v->nx : 1;
    for (int z = 0; z < v->nz_buf; ++z)
        for (int y = 0; y < v->ny; ++y)
            for (int x = 0; x < v->nx; ++x) {
                const long long idx = z * plane + (long long)y * v->nx + x;
                const int a = axis == 0 ? z : axis == 1 ? y : x;
                double acc = (double)src[idx] * w[0];
                for (int j = radius; j >= 1; --j) {
                    const double lo = src[idx + (long long)(reflect_index(a - j, n_axis) - a) * stride];
                    const double hi = src[idx + (long long)(reflect_index(a + j, n_axis) - a) * stride];
                    const double pair = lo + hi;
                    acc = acc + pair * w[j];
                }
                dst[idx] = (float)acc;
            }
    return 0;
}
}  // namespace

extern "C" {
int nb200_gauss_axis(const float* src, float* dst, const nb200_vol* vol, int axis, const double* weights, int radius,
                     void*) {
    if (!src || !dst || src == dst || !vol || vol->zc0 != 0 || vol->zc1 != vol->nz_buf || vol->zg_off != 0) return -1;
    return emu_axis(src, dst, vol, axis, weights, radius);
}
// Y then X with the float32 intermediate scipy stores between the axes; dst doubles as that intermediate's destination
int nb200_gauss_yx(const float* src, float* dst, const nb200_vol* vol, const double* wy, const double* wx, int radius,
                   void*) {
    if (!src || !dst || src == dst || !vol || radius < 1 || radius > 8) return -4;
    const long long n = (long long)vol->nz_buf * vol->ny * vol->nx;
    float* tmp = new float[n];
    emu_axis(src, tmp, vol, 1, wy, radius);
    emu_axis(tmp, dst, vol, 2, wx, radius);
    delete[] tmp;
    return 0;
}
const char* nb200_last_error(void) { return nb::g_err; }
}
