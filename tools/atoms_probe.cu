// Shared-memory atomic throughput probe (sm_100a): how many 32-bit atomicAdd lanes per clock an SM retires when every
// lane hits a pseudo-random word of an 8192-word array -- the inner operation of the inverted-file scorer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atoms_probe tools/atoms_probe.cu && tools/atoms_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, int words, unsigned int *out)
{
    extern __shared__ int acc[];
    for (int t = threadIdx.x; t < words; t += 256) acc[t] = 0;
    __syncthreads();
    unsigned int x = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = x * 1664525u + 1013904223u;
            const int a = (x >> 10) & (words - 1);
            if (MODE == 0) atomicAdd(&acc[a], 1);
            else if (MODE == 1) acc[a] += 1;                                // racy read-modify-write: LDS + STS
            else if (MODE == 2) atomicAdd(&acc[(a & ~31) | (threadIdx.x & 31)], 1);   // conflict-free banks
            else if (MODE == 3) atomicAdd(reinterpret_cast<unsigned long long *>(acc) + (a >> 1), 1ull);
            else if (MODE == 4) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&acc[a])), "r"(1));
        }
    }
    __syncthreads();
    unsigned int s = 0;
    for (int t = threadIdx.x; t < words; t += 256) s += acc[t];
    if (s == 0xdeadbeef) out[0] = s;
}
template <int MODE> void run(const char *name, int ctas_per_sm, int words)
{
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    unsigned int *out; cudaMalloc(&out, 4);
    const int iters = 4096;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<sm * ctas_per_sm, 256, words * 4>>>(16, words, out);
    cudaEventRecord(a);
    probe<MODE><<<sm * ctas_per_sm, 256, words * 4>>>(iters, words, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)sm * ctas_per_sm * 256 * iters * 8;
    printf("%-44s %d CTAs/SM: %7.1f G lane-ops/s = %.2f lanes/clk/SM (at %.2f GHz)  %s\n", name, ctas_per_sm, ops / ms / 1e6,
           ops / ms / 1e6 / sm / (khz / 1e6), khz / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    for (int c : {1, 2, 4, 6}) run<0>("atomicAdd u32, random word of 8192", c, 8192);
    for (int c : {4}) run<1>("ld + add + st (racy), random word", c, 8192);
    for (int c : {4}) run<2>("atomicAdd u32, random row, own bank", c, 8192);
    for (int c : {4}) run<3>("atomicAdd u64, random", c, 8192);
    for (int c : {4}) run<4>("red.shared.add.u32, random", c, 8192);
    return 0;
}
