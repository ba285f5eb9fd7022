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

int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace es

extern "C" {

int es_version(void) { return 100; }

const char* es_last_error(void) { return es::g_err; }

int es_device_sm_count(void) { return es::sm_count(); }

}
