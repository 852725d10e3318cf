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

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
    return dev;
}

int sm_count() {
    static int cache[MAX_DEVICES] = {0};
    const int dev = current_device();
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;  // B200
        cache[dev] = n;
    }
    return cache[dev];
}

}  // namespace dcr

// In-band SM clock probe: one thread spins for ~100 us and reports cycles / nanosecond = the SM clock it ran at.
// bench.py launches it on a side stream while the timed kernels run; NVML's clock query was observed to stall GPU
// work for up to 200 ms on the shared hosts, which a measurement taken under load cannot afford.
__global__ void sm_clock_probe_kernel(float* out) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < 100000ull);
    const long long c1 = clock64();
    *out = (float)((double)(c1 - c0) * 1000.0 / (double)(t1 - t0));   // MHz
}

extern "C" int dcr_sm_clock_probe(float* out_mhz, void* stream) {
    if (!out_mhz) { dcr::set_error("dcr_sm_clock_probe: out_mhz is NULL"); return 1; }
    sm_clock_probe_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out_mhz);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" const char* dcr_last_error(void) { return dcr::g_err; }
extern "C" int dcr_version(void) { return DCR_VERSION; }
