// Library identification and error text of the eaQHM B200 C ABI.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void eaqhm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int eaqhm_version(void) { return EAQHM_VERSION; }
extern "C" const char* eaqhm_last_error(void) { return g_err; }
extern "C" int eaqhm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
