// Microbenchmark: FP64 DFMA dependent-issue latency and throughput on one SM sub-partition.
// chains = independent accumulators per thread, warps per block varies; 1 block on 1 SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, long long* cyc, int iters, double a, double b)
{
    double x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = a + c + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = fma(x[c], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int C> void run(int warps)
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
    const int iters = 2000;
    k<C><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    k<C><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (iters * 8.0 * C);
    printf("chains=%d warps/SM=%2d (per SMSP %.1f): %.2f cycles per DFMA per warp, SMSP DFMA rate %.3f /cycle\n", C, warps, warps / 4.0,
           per, (warps / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
