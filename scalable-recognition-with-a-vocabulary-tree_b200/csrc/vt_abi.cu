// vt_abi.cu -- error plumbing, device info and descriptor staging kernels of the C ABI.
#include "vt_common.cuh"
#include <stdarg.h>
#include <string.h>
#include <thread>
#include <vector>

static thread_local char g_err[512] = "";

void vt_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" int vt_abi_version(void) { return VT_ABI_VERSION; }
extern "C" const char *vt_last_error(void) { return g_err; }

int vt_sm_count()
{
    // per device: a process may drive several GPUs (VocabularyTree(device="cuda:1") while device 0 is current)
    static int cached[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached[dev] = n;
        else
            return 148;
    }
    return cached[dev];
}

extern "C" int vt_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0, n = 0, ma = 0, mi = 0;
    VT_CUDA(cudaGetDevice(&dev));
    VT_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    VT_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    VT_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    if (ma != 10) {
        vt_set_error("vt_device_info: device is sm_%d%d; this library only carries sm_100a code", ma, mi);
        return VT_EUNSUPPORTED;
    }
    return VT_OK;
}

// ---- host staging: many small host arrays -> one (pinned) host buffer ----------------------------------------------
// The images of a batch are separate host arrays; the upload wants them back to back in pinned memory.  One memcpy
// thread moves ~10 GB/s, the upload that follows 55 GB/s, so the copy is spread over n_threads threads, each taking a
// contiguous range of arrays of about equal bytes.  Pure host code: no device is touched.
extern "C" int vt_host_pack(const void *const *h_src, const int64_t *h_bytes, int64_t n, void *h_dst, int n_threads)
{
    VT_CHECK_ARG(n >= 0 && (n == 0 || (h_src && h_bytes && h_dst)) && n_threads >= 1, "bad arguments");
    if (n == 0) return VT_OK;
    std::vector<int64_t> off((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        VT_CHECK_ARG(h_bytes[i] >= 0 && (h_bytes[i] == 0 || h_src[i]), "negative size or null source");
        off[(size_t)i + 1] = off[(size_t)i] + h_bytes[i];
    }
    const int64_t total = off[(size_t)n];
    auto copy_range = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i)
            if (h_bytes[i]) memcpy((char *)h_dst + off[(size_t)i], h_src[i], (size_t)h_bytes[i]);
    };
    if (n_threads > 64) n_threads = 64;
    if (n_threads == 1 || total < (8 << 20)) { copy_range(0, n); return VT_OK; }
    std::vector<std::thread> pool;
    int64_t lo = 0;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t want = total / n_threads * (t + 1);
        int64_t hi = lo;
        while (hi < n && (t + 1 == n_threads || off[(size_t)hi + 1] <= want)) ++hi;
        if (hi > lo) pool.emplace_back(copy_range, lo, hi);
        lo = hi;
    }
    for (auto &th : pool) th.join();
    return VT_OK;
}

// ---- staging ------------------------------------------------------------------------------------
__global__ void stage_f32_to_u8_kernel(const float *__restrict__ src, int64_t n, int d, uint8_t *__restrict__ dst,
                                       int dp, int *__restrict__ flag)
{
    const int64_t total = n * (int64_t)dp;
    bool bad = false;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / dp;
        const int c = (int)(t - r * dp);
        uint8_t out = 0;
        if (c < d) {
            const float v = src[r * d + c];
            const float q = rintf(v);
            if (!(q == v) || v < 0.f || v > 255.f) bad = true;
            out = (uint8_t)fminf(fmaxf(q, 0.f), 255.f);
        }
        dst[t] = out;
    }
    if (__any_sync(VT_FULL, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}

// unpadded rows (d == dp, a multiple of 16): a thread converts 16 consecutive values -- four 16-byte loads in flight,
// one 16-byte store; HBM bound (5 B moved per value) instead of one 4-byte load and a 64-bit division per byte
__global__ void __launch_bounds__(256)
stage_f32_to_u8_x16_kernel(const float4 *__restrict__ src, int64_t n16, uint4 *__restrict__ dst, int *__restrict__ flag)
{
    bool bad = false;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n16; t += (int64_t)gridDim.x * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldcs(src + 4 * t + i);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float f[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            uint32_t pack = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float q = rintf(f[j]);
                if (!(q == f[j]) || f[j] < 0.f || f[j] > 255.f) bad = true;
                pack |= (uint32_t)fminf(fmaxf(q, 0.f), 255.f) << (8 * j);
            }
            w[i] = pack;
        }
        dst[t] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (__any_sync(VT_FULL, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}

extern "C" int vt_stage_f32_to_u8(const float *d_src, int64_t n, int d, uint8_t *d_dst, int dp, int *d_flag, void *stream)
{
    VT_CHECK_ARG(n >= 0 && d > 0 && dp >= d, "bad shape");
    if (n == 0) return VT_OK;
    const int64_t total = n * (int64_t)dp;
    if (d == dp && d % 16 == 0 && ((uintptr_t)d_src & 15) == 0 && ((uintptr_t)d_dst & 15) == 0) {
        const int64_t n16 = total / 16, want = (n16 + 255) / 256, cap = (int64_t)vt_sm_count() * 16;
        stage_f32_to_u8_x16_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
            (const float4 *)d_src, n16, (uint4 *)d_dst, d_flag);
        VT_LAUNCH_CHECK();
        return VT_OK;
    }
    const int blocks = (int)((total + 255) / 256 < (int64_t)vt_sm_count() * 16 ? (total + 255) / 256 : (int64_t)vt_sm_count() * 16);
    stage_f32_to_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, n, d, d_dst, dp, d_flag);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

template <typename E>
__global__ void stage_pad_kernel(const E *__restrict__ src, int64_t n, int d, E *__restrict__ dst, int dp)
{
    const int64_t total = n * (int64_t)dp;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / dp;
        const int c = (int)(t - r * dp);
        dst[t] = c < d ? src[r * d + c] : (E)0;
    }
}

extern "C" int vt_stage_pad(const void *d_src, int64_t n, int d, void *d_dst, int dp, int elem_size, void *stream)
{
    VT_CHECK_ARG(n >= 0 && d > 0 && dp >= d, "bad shape");
    VT_CHECK_ARG(elem_size == 1 || elem_size == 4, "elem_size must be 1 or 4");
    if (n == 0) return VT_OK;
    const int64_t total = n * (int64_t)dp;
    const int blocks = (int)((total + 255) / 256 < (int64_t)vt_sm_count() * 16 ? (total + 255) / 256 : (int64_t)vt_sm_count() * 16);
    if (elem_size == 1)
        stage_pad_kernel<uint8_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)d_src, n, d, (uint8_t *)d_dst, dp);
    else
        stage_pad_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)d_src, n, d, (float *)d_dst, dp);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
