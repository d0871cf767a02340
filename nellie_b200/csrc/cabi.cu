// Library-level entry points of the C ABI (include/nellie_b200.h): version, error text, device facts.
#include <stdarg.h>

#include "common.cuh"

namespace nb {

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
}

}  // namespace nb

extern "C" {

int nb200_abi_version(void) { return NB200_ABI_VERSION; }

const char* nb200_last_error(void) { return nb::last_error_buf(); }

int nb200_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

}  // extern "C"
