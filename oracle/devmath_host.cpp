// Host build of nellie_b200/csrc/devmath.cuh for CPU-side numerics tests (TEST INFRASTRUCTURE ONLY).
// Compiled by tests with:  g++ -O2 -ffp-contract=off -mfma -shared -fPIC
// It lets `pytest -m "not gpu"` check the exact per-voxel arithmetic the CUDA kernels run
// (exp, eigenvalues, vesselness) against numpy without a GPU.  Never loaded by nellie_b200.
#include "../nellie_b200/csrc/devmath.cuh"

extern "C" {
void hm_expf(const float* x, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = nb::np_expf(x[i]);
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
void hm_vesselness2(const float* ev2, float* v, long n, float beta_sq, float gamma_sq) {
    for (long i = 0; i < n; ++i) v[i] = nb::vesselness2(ev2[2 * i], ev2[2 * i + 1], beta_sq, gamma_sq);
}
}
