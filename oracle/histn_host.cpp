// Host build of nellie_b200/csrc/histn.cu through oracle/cuda_emu.h (TEST INFRASTRUCTURE ONLY), plus plain restatements of the
// two production entry points it builds on (nb200_hist_reset, nb200_hist_minmax: shared-memory / shuffle kernels of
// thresholds.cu, GPU-tested) and of nb200_strided_sample.  Built by __graft_entry__.build() with
//   g++ -O2 -ffp-contract=off -mfma -shared -fPIC -DNB200_HOST_EMU='"<repo>/oracle/cuda_emu.h"' -x c++ oracle/histn_host.cpp
// Never loaded by nellie_b200.
#include "../nellie_b200/csrc/histn.cu"

namespace {
inline unsigned ordered_of(float f) {
    const unsigned u = nb::f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
}  // namespace

extern "C" {
int nb200_hist_reset(long long* state, void*) {
    for (int i = 0; i < NB200_HIST_WORDS; ++i) state[i] = 0;
    state[NB200_HIST_MIN] = 0xffffffffLL;
    return 0;
}
int nb200_hist_minmax(const float* vals, long long n, int transform, const double*, long long* state, void*) {
    if (transform == NB200_TF_DIV) return -4;
    for (long long i = 0; i < n; ++i) {
        float v;
        if (!kept_value(vals[i], transform == NB200_TF_LOG10, v)) continue;
        const long long k = (long long)ordered_of(v);
        if (k < state[NB200_HIST_MIN]) state[NB200_HIST_MIN] = k;
        if (k > state[NB200_HIST_MAX]) state[NB200_HIST_MAX] = k;
        state[NB200_HIST_COUNT] += 1;
    }
    return 0;
}
int nb200_strided_sample(const float* src, long long n, long long offset, long long step, const float* gate, float gate_thresh,
                         float* out, void*) {
    long long k = 0;
    for (long long j = offset; j < n; j += step, ++k) {
        float v = src[j];
        if (gate && !(gate[j] > gate_thresh)) v = 0.0f;
        out[k] = v;
    }
    return 0;
}
const char* nb200_last_error(void) { return nb::g_err; }
}
