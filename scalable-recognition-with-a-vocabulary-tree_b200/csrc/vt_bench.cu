// vt_bench.cu -- measurement helper: FP32 FFMA microbenchmark.  The quantiser's roofline is the
// FP32 pipe; this measures what that pipe delivers on the box at the clocks
// it actually runs at, so bench.py does not have to trust a nominal 148 x 128 x 1.965 GHz figure.
#include "vt_common.cuh"

namespace {
__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

// best-of-`reps` FP32 throughput in TFLOP/s (2 flop per FFMA); d_scratch: >= sm_count*8*256 floats
extern "C" int vt_measure_fp32_peak(float *d_scratch, int iters, int reps, double *tflops, double *ms_best, void *stream)
{
    VT_CHECK_ARG(d_scratch && tflops && iters > 0 && reps > 0, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = vt_sm_count() * 8;
    cudaEvent_t e0, e1;
    VT_CUDA(cudaEventCreate(&e0));
    VT_CUDA(cudaEventCreate(&e1));
    double best = 1e30;
    for (int r = 0; r < reps + 1; ++r) {
        VT_CUDA(cudaEventRecord(e0, s));
        ffma_kernel<<<blocks, 256, 0, s>>>(d_scratch, iters, 0.999f, 0.001f);
        VT_CUDA(cudaEventRecord(e1, s));
        VT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        VT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;                     // first launch is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    VT_LAUNCH_CHECK();
    const double flop = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    *tflops = flop / (best * 1e-3) / 1e12;
    if (ms_best) *ms_best = best;
    return VT_OK;
}
