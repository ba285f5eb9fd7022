// echoseal_b200/csrc/api.cu — C-ABI plumbing shared by all kernels (error text, device info).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace es {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= ES_MAX_DEVICES) return 0;
    return dev;
}

int sm_count()
{
    static int n[ES_MAX_DEVICES] = {0};
    const int dev = current_device();
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

}  // namespace es

extern "C" {

int es_version(void) { return 100; }

const char* es_last_error(void) { return es::g_err; }

int es_device_sm_count(void) { return es::sm_count(); }

}
