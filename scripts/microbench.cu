// Pipe-throughput microbenchmarks used to size the step kernel (not product code).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float a, double da)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double d0 = threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int acc = 0;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { x0 = fmaf(x0, a, 1.f); x1 = fmaf(x1, a, 1.f); x2 = fmaf(x2, a, 1.f); x3 = fmaf(x3, a, 1.f); }
        if (MODE == 1) { d0 = fma(d0, da, 1.0); d1 = fma(d1, da, 1.0); d2 = fma(d2, da, 1.0); d3 = fma(d3, da, 1.0); }
        if (MODE == 2) { d0 = (double)x0; x0 = (float)(d0 + da); d1 = (double)x1; x1 = (float)(d1 + da); }
        if (MODE == 3) { if ((threadIdx.x & 7) == 0) acc += atomicAdd(&cnt, 1); }
        if (MODE == 4) { x0 = __frcp_rn(x0 + a); x1 = __frcp_rn(x1 + a); x2 = __frcp_rn(x2 + a); x3 = __frcp_rn(x3 + a); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + (float)(d0 + d1 + d2 + d3) + acc;
}
int main()
{
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 100000;
    const char *names[] = {"FFMA x4", "DFMA x4", "F2F round trip x2 (+2 DADD)", "smem atomicAdd same addr (1 of 8 lanes)", "frcp_rn x4"};
    for (int mode = 0; mode < 5; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 3, 256>>>(out, iters, 1.0001f, 1.0001);
            if (mode == 1) k<1><<<148 * 3, 256>>>(out, iters, 1.0001f, 1.0001);
            if (mode == 2) k<2><<<148 * 3, 256>>>(out, iters, 1.0001f, 1.0001);
            if (mode == 3) k<3><<<148 * 3, 256>>>(out, iters / 10, 1.0001f, 1.0001);
            if (mode == 4) k<4><<<148 * 3, 256>>>(out, iters, 1.0001f, 1.0001);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1) {
                const double it = (mode == 3) ? iters / 10 : iters;
                const double ops_per_thread_iter = (mode == 2) ? 2 : (mode == 3 ? 1.0 / 8 : 4);
                const double ops = it * ops_per_thread_iter * 148.0 * 3 * 256;
                printf("%-42s %8.3f ms  -> %.1f ops/clk/SM @1.9GHz\n", names[mode], ms, ops / (ms * 1e-3) / 148 / 1.9e9);
            }
        }
    }
    return 0;
}
