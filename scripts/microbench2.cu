// Shared-memory pipe microbenchmarks that size the round-2 step kernel (not product code):
// ATOMS.ADD (s32) scatter, LDS.128 / LDS.32 / LDS.U16 gathers with random and grouped addresses.
// One 1024-thread CTA per SM; prints SM cycles per warp instruction (whole SM, all 32 warps issuing).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NW = 16384;          // 64 KB of words
__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(unsigned long long *cyc, int *sink, int iters)
{
    extern __shared__ int sm[];
    for (int i = threadIdx.x; i < NW; i += blockDim.x) sm[i] = i;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
    // precomputed addresses (8 per thread) so that the loop body is the memory op + little else
    int a[8];
    for (int q = 0; q < 8; q++) {
        uint32_t r = lcg(s);
        if (MODE == 0 || MODE == 1) a[q] = r % NW;                                   // ATOMS random word
        if (MODE == 2) a[q] = (r % (NW / 4)) * 4;                                     // LDS.128 random
        if (MODE == 3) { uint32_t g = __shfl_sync(0xffffffffu, r, threadIdx.x & 24); a[q] = (g % (NW / 4)) * 4; }   // LDS.128, 8 lanes share
        if (MODE == 4) a[q] = r % NW;                                                 // LDS.32 random
        if (MODE == 5) a[q] = r % (NW * 2);                                           // LDS.U16 random
        if (MODE == 6) a[q] = (r % 32) * 12;                                          // LDS.128 from a 32-entry table of 48 B records
        if (MODE == 7) a[q] = ((threadIdx.x & 31) * 4 + q * 128) % NW;                // LDS.128 conflict-free (reference)
        if (MODE == 8) a[q] = (r % 1400) * 3;                                         // 3 x ATOMS to x,y,z of a random bead (AoS stride 3)
        if (MODE == 9) a[q] = r % 1400;                                               // 3 x ATOMS to SoA arrays
    }
    int acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (MODE == 0) atomicAdd(&sm[a[q]], i);
            if (MODE == 1) acc += atomicAdd(&sm[a[q]], i);
            if (MODE == 2 || MODE == 3 || MODE == 6 || MODE == 7) { const int4 v = *reinterpret_cast<const int4 *>(&sm[a[q]]); acc += v.x ^ v.y ^ v.z ^ v.w; }
            if (MODE == 4) acc += sm[a[q]];
            if (MODE == 5) acc += reinterpret_cast<const unsigned short *>(sm)[a[q]];
            if (MODE == 8) { atomicAdd(&sm[a[q]], i); atomicAdd(&sm[a[q] + 1], i); atomicAdd(&sm[a[q] + 2], i); }
            if (MODE == 9) { atomicAdd(&sm[a[q]], i); atomicAdd(&sm[a[q] + 1408], i); atomicAdd(&sm[a[q] + 2816], i); }
            a[q] = (a[q] + ((MODE == 2 || MODE == 3 || MODE == 7) ? 4 * 37 : 0)) & (NW - 1) ;
            if (MODE == 6) a[q] = (a[q] + 12 * 7) % 384;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[threadIdx.x];
}
int main()
{
    unsigned long long *cyc; int *sink;
    cudaMallocManaged(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
    const char *names[] = {"ATOMS.ADD s32 random word, no return", "ATOMS.ADD s32 random word, with return", "LDS.128 random", "LDS.128 8 lanes share an address",
                           "LDS.32 random", "LDS.U16 random", "LDS.128 32-entry table of 48 B records", "LDS.128 conflict-free", "3 x ATOMS.ADD to a random bead (AoS)", "3 x ATOMS.ADD to a random bead (SoA)"};
    const int iters = 2000;
    for (int mode = 0; mode < 10; mode++) {
        for (int rep = 0; rep < 2; rep++) {
#define L(M) if (mode == M) { cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, NW * 4); k<M><<<148, 1024, NW * 4>>>(cyc, sink, iters); }
            L(0) L(1) L(2) L(3) L(4) L(5) L(6) L(7) L(8) L(9)
            cudaDeviceSynchronize();
        }
        const double ops = (double)iters * 8 * 32 * ((mode == 8 || mode == 9) ? 3 : 1);     // warp instructions per SM
        printf("%-44s %7.2f cycles per warp instruction (SM-wide)  err=%s\n", names[mode], (double)cyc[0] / ops, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
