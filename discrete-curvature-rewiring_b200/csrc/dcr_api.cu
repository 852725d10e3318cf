// dcr_api.cu — error plumbing and version of libdcr.
#include <stdarg.h>
#include <string.h>

#include "dcr_common.cuh"

namespace dcr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return 100 + (int)e;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;  // B200
    }
    return n;
}

}  // namespace dcr

extern "C" const char* dcr_last_error(void) { return dcr::g_err; }
extern "C" int dcr_version(void) { return DCR_VERSION; }
