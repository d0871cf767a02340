// Host build of nellie_b200/csrc/devmath.cuh for CPU-side numerics tests (TEST INFRASTRUCTURE ONLY).
// Compiled by tests with:  g++ -O2 -ffp-contract=off -mfma -shared -fPIC
// It lets `pytest -m "not gpu"` check the exact per-voxel arithmetic the CUDA kernels run
// (exp, eigenvalues, vesselness) against numpy without a GPU.  Never loaded by nellie_b200.
#include "../nellie_b200/csrc/devmath.cuh"

extern "C" {
void hm_expf(const float* x, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = nb::np_expf(x[i]);
}
void hm_log10f(const float* x, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = nb::np_log10f(x[i]);
}
void hm_expf_nonpos(const float* x, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = nb::np_expf_nonpos(x[i]);
}
void hm_eig3(const float* h6, float* ev3, long n, int iters) {
    for (long i = 0; i < n; ++i) {
        const float* a = h6 + 6 * i;
        float e0, e1, e2;
        if (iters == 1) nb::eig3_sym<1>(a[0], a[1], a[2], a[3], a[4], a[5], e0, e1, e2);
        else if (iters == 2) nb::eig3_sym<2>(a[0], a[1], a[2], a[3], a[4], a[5], e0, e1, e2);
        else nb::eig3_sym<3>(a[0], a[1], a[2], a[3], a[4], a[5], e0, e1, e2);
        ev3[3 * i] = e0; ev3[3 * i + 1] = e1; ev3[3 * i + 2] = e2;
    }
}
void hm_eig2(const float* h3, float* ev2, long n) {
    for (long i = 0; i < n; ++i) nb::eig2_sym(h3[3 * i], h3[3 * i + 1], h3[3 * i + 2], ev2[2 * i], ev2[2 * i + 1]);
}
void hm_vesselness3(const float* ev3, float* v, long n, float alpha_sq, float beta_sq, float gamma_sq) {
    for (long i = 0; i < n; ++i) v[i] = nb::vesselness3(ev3[3 * i], ev3[3 * i + 1], ev3[3 * i + 2], alpha_sq, beta_sq, gamma_sq);
}
// The two "response is provably zero" tests with margins relative to the voxel's own Frobenius norm, restated from
// nellie_b200/csrc/frangi.cu (voxel_code: diagonal test, K2) and sparse.cu (full minor / determinant test, K3):
// out[i] bit 0 = K2 flags the voxel, bit 1 = K3 rejects it.  frob_sq is computed as the kernels do (frob_sq3).
void hm_zero_tests(const float* h6, unsigned char* out, long n) {
    for (long i = 0; i < n; ++i) {
        const float* a = h6 + 6 * i;           // zz, zy, zx, yy, yx, xx
        const float fs = nb::frob_sq3(a[0], a[1], a[2], a[3], a[4], a[5]);
        const float m = fmaxf(fmaxf(a[0] + a[3], a[0] + a[5]), a[3] + a[5]);
        const bool diag = m > 0.0f && m * m > 1.001e-10f * fs && fs > 1e-20f && fs < 1e20f;
        float tau2 = INFINITY, tau3 = INFINITY;
        if (fs > 1e-20f && fs < 1e20f) {
            const float f2 = 1.001f * fs;
            tau2 = 1e-5f * f2;
            tau3 = 1e-4f * (f2 * sqrtf(f2));
        }
        const bool full = nb::pd_reject_full(a[0], a[1], a[2], a[3], a[4], a[5], tau2, tau3);
        out[i] = (unsigned char)((diag ? 1 : 0) | (full ? 2 : 0));
    }
}
void hm_vesselness2(const float* ev2, float* v, long n, float beta_sq, float gamma_sq) {
    for (long i = 0; i < n; ++i) v[i] = nb::vesselness2(ev2[2 * i], ev2[2 * i + 1], beta_sq, gamma_sq);
}
}
