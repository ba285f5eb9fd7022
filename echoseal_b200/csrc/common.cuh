// echoseal_b200/csrc/common.cuh — shared helpers for the sm_100a kernels and the C-ABI glue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace es {

// last error text returned by es_last_error()
void set_error(const char* fmt, ...);

#define ES_CUDA_OK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            es::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -(int)_e - 1000;                                                       \
        }                                                                                 \
    } while (0)

constexpr int ES_OK = 0;
constexpr int ES_EINVAL = -1;
constexpr int ES_ENOTREADY = -2;

constexpr int POLAR_N = 1024;
constexpr int POLAR_NLOG = 10;

// Per-device state: constant tables, function attributes and SM counts are cached PER DEVICE (index = cudaGetDevice()),
// never process-wide.  A table upload whose content differs from what the device already holds first drains the
// device (cudaDeviceSynchronize), so kernels in flight on other streams keep reading consistent tables.
constexpr int ES_MAX_DEVICES = 64;
int current_device();          // cudaGetDevice(), clamped to [0, ES_MAX_DEVICES)
int sm_count();                // of the current device

// tx.cu keeps its own constant-memory copy of the polar code layout
int tx_set_code(const uint16_t* pos, int K);

}  // namespace es
