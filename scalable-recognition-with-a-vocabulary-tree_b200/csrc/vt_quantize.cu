// vt_quantize.cu -- fused L-level greedy tree descent (VocabularyTree.propagate_feature,
// cbir/encoders/vocabulary.py:104-124) for a whole batch of descriptors.
//
// Mapping: 8 lanes per descriptor (4 descriptors per warp); a lane keeps DP/8 dimensions of its
// descriptor in registers for the whole descent, so the descriptor is read from HBM exactly once
// (DP bytes for u8 input) and only the 4-byte word id is written.  Child centroids are streamed
// through L1/L2 as coalesced 128 B rows.  Algorithmic work per descriptor: 2*k*DP*L float32
// lane-slots (sub + fma), DP*sizeof(T) + 4 bytes.
#include "vt_common.cuh"
#include <stdlib.h>

template <typename T, int DP>
__global__ void __launch_bounds__(256, 3)
quantize_kernel(const T *__restrict__ desc, int64_t n, const float *__restrict__ cent,
                const uint8_t *__restrict__ internal, const int32_t *__restrict__ ref_id, int k, int depth,
                int64_t n_heap, int32_t start, int32_t *__restrict__ leaf, int32_t *__restrict__ word,
                unsigned long long *__restrict__ stats)
{
    constexpr int VEC = DP / 32;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned int rechecks = 0;

    for (int64_t base = warp * 4; base < n; base += nwarps * 4) {       // warp-uniform trip count
        const int64_t i = base + g;
        const bool valid = i < n;
        const T *xrow = desc + (size_t)(valid ? i : 0) * DP;
        float4 x[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[v] = Elem<T>::load4(xrow, j + 8 * v);

        int64_t node = start;
        bool active = valid && (node * k + k < n_heap) && internal[node];
        for (int lvl = 0; lvl < depth; ++lvl) {
            if (!__any_sync(VT_FULL, active)) break;
            const float *children = cent + (size_t)(node * k + 1) * DP;
            const int c = descend_one_level<T, DP>(x, xrow, children, k, active, lane, rechecks);
            if (active) {
                node = node * k + 1 + c;
                active = (node * k + k < n_heap) && internal[node];
            }
        }
        if (valid && j == 0) {
            if (leaf) leaf[i] = (int32_t)node;
            if (word) word[i] = ref_id[node];
        }
    }
    if (stats && rechecks) atomicAdd(stats, (unsigned long long)rechecks);
}

static int64_t heap_nodes(int k, int depth)
{
    int64_t n = 0, p = 1;
    for (int l = 0; l <= depth; ++l) { n += p; p *= k; }
    return n;
}

template <typename T, int DP>
static int launch_quantize(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal,
                           const int32_t *ref_id, int k, int depth, int32_t start, int32_t *leaf, int32_t *word,
                           uint64_t *stats, cudaStream_t s)
{
    const int64_t want = (n + 31) / 32;                       // 32 descriptors per 256-thread block pass
    const int64_t cap = (int64_t)vt_sm_count() * 8;
    const int blocks = (int)(want < cap ? want : cap);
    quantize_kernel<T, DP><<<blocks, 256, 0, s>>>((const T *)d_desc, n, cent, internal, ref_id, k, depth,
                                                  heap_nodes(k, depth), start, leaf, word,
                                                  (unsigned long long *)stats);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

int vt_quantize_dispatch(const void *d_desc, int dtype, int64_t n, int dp, const float *cent, const uint8_t *internal,
                         const int32_t *ref_id, int k, int depth, int32_t start, int32_t *leaf, int32_t *word,
                         uint64_t *stats, cudaStream_t s)
{
#define VT_Q(T, DPV) return launch_quantize<T, DPV>(d_desc, n, cent, internal, ref_id, k, depth, start, leaf, word, stats, s)
    if (dtype == VT_U8) {
        if (dp == 32) VT_Q(uint8_t, 32);
        if (dp == 64) VT_Q(uint8_t, 64);
        if (dp == 128) VT_Q(uint8_t, 128);
    } else if (dtype == VT_F32) {
        if (dp == 32) VT_Q(float, 32);
        if (dp == 64) VT_Q(float, 64);
        if (dp == 128) VT_Q(float, 128);
    }
#undef VT_Q
    vt_set_error("vt_quantize: unsupported dtype %d / dp %d (dp must be 32, 64 or 128)", dtype, dp);
    return VT_EUNSUPPORTED;
}

extern "C" int vt_quantize(const void *d_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                           const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth, int32_t start,
                           int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, void *stream)
{
    VT_CHECK_ARG(n >= 0, "n < 0");
    VT_CHECK_ARG(k >= 1 && k <= 255, "k out of range");
    VT_CHECK_ARG(depth >= 0 && depth <= 16, "depth out of range");
    VT_CHECK_ARG(d_centroids && d_internal, "null tree");
    VT_CHECK_ARG(d_word == nullptr || d_ref_id != nullptr, "d_word needs d_ref_id");
    VT_CHECK_ARG(start >= 0 && start < heap_nodes(k, depth), "start node outside the heap");
    if (n == 0) return VT_OK;
    return vt_quantize_dispatch(d_desc, dtype, n, dp, d_centroids, d_internal, d_ref_id, k, depth, start,
                                d_leaf_heap, d_word, d_stats, (cudaStream_t)stream);
}

// ---- host-buffer pipeline (end-to-end path) ---------------------------------------------------------
int vt_quantize_sorted_impl(const void *d_desc, int64_t n, int dp, const float *d_centroids, const uint8_t *d_internal,
                            const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap, int32_t *d_word,
                            uint64_t *d_stats, void *d_work, cudaStream_t s, cudaEvent_t *ev);
int vt_mma_prepare(const float *cent, int k, int depth, void *planes_and_norms, uint64_t *stats, cudaStream_t s);
int64_t vt_mma_prepared_bytes(int k, int depth);
int vt_mma_levels(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id, int k,
                  int depth, int32_t *leaf, int32_t *word, uint64_t *stats, const void *prepared, void *sortwork,
                  cudaStream_t s, cudaEvent_t *ev, int *n_ev_io);
// Staging state of vt_quantize_host, kept between calls (allocating and freeing ~2 GB of device
// buffers per call costs more than the copies it overlaps).  Released by vt_host_release().
namespace {
constexpr int HP_NS = 3;
struct HostPipe {
    bool valid = false;
    int device = -1;
    int64_t chunk_rows = 0, row_bytes = 0, work_bytes = 0, prepared_bytes = 0;
    cudaStream_t st[HP_NS] = {};
    void *dbuf[HP_NS] = {};
    void *work[HP_NS] = {};
    int32_t *wbuf[HP_NS] = {};
    void *prepared = nullptr;
    unsigned long long *d_flag = nullptr;
};
HostPipe g_pipe;

void host_pipe_free()
{
    for (int s = 0; s < HP_NS; ++s) {
        if (g_pipe.st[s]) { cudaStreamSynchronize(g_pipe.st[s]); cudaStreamDestroy(g_pipe.st[s]); }
        if (g_pipe.dbuf[s]) cudaFree(g_pipe.dbuf[s]);
        if (g_pipe.wbuf[s]) cudaFree(g_pipe.wbuf[s]);
        if (g_pipe.work[s]) cudaFree(g_pipe.work[s]);
    }
    if (g_pipe.prepared) cudaFree(g_pipe.prepared);
    if (g_pipe.d_flag) cudaFree(g_pipe.d_flag);
    g_pipe = HostPipe();
}

int host_pipe_prepare(int64_t chunk_rows, int64_t row_bytes, int64_t work_bytes, int64_t prepared_bytes)
{
    int dev = 0;
    VT_CUDA(cudaGetDevice(&dev));
    if (g_pipe.valid && g_pipe.device == dev && g_pipe.chunk_rows == chunk_rows && g_pipe.row_bytes == row_bytes &&
        g_pipe.work_bytes >= work_bytes && g_pipe.prepared_bytes >= prepared_bytes)
        return VT_OK;
    host_pipe_free();
    bool ok = true;
    for (int s = 0; s < HP_NS && ok; ++s) {
        ok = cudaStreamCreateWithFlags(&g_pipe.st[s], cudaStreamNonBlocking) == cudaSuccess &&
             cudaMalloc(&g_pipe.dbuf[s], (size_t)(chunk_rows * row_bytes)) == cudaSuccess &&
             cudaMalloc((void **)&g_pipe.wbuf[s], (size_t)chunk_rows * sizeof(int32_t)) == cudaSuccess &&
             (work_bytes == 0 || cudaMalloc(&g_pipe.work[s], (size_t)work_bytes) == cudaSuccess);
    }
    ok = ok && (prepared_bytes == 0 || cudaMalloc(&g_pipe.prepared, (size_t)prepared_bytes) == cudaSuccess);
    ok = ok && cudaMalloc((void **)&g_pipe.d_flag, 4 * sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) {
        vt_set_error("vt_quantize_host: staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        host_pipe_free();
        return VT_ENOMEM;
    }
    g_pipe.valid = true;
    g_pipe.device = dev;
    g_pipe.chunk_rows = chunk_rows;
    g_pipe.row_bytes = row_bytes;
    g_pipe.work_bytes = work_bytes;
    g_pipe.prepared_bytes = prepared_bytes;
    return VT_OK;
}
}  // namespace

extern "C" int vt_host_release(void)
{
    host_pipe_free();
    return VT_OK;
}

extern "C" int vt_quantize_host(const void *h_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                                const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth,
                                int32_t *h_word, int64_t chunk_rows)
{
    VT_CHECK_ARG(n >= 0 && h_desc && h_word, "bad host buffers");
    VT_CHECK_ARG(dtype == VT_U8 || dtype == VT_F32, "bad dtype");
    VT_CHECK_ARG(d_ref_id != nullptr, "d_ref_id required");
    if (n == 0) return VT_OK;
    const int64_t esz = dtype == VT_U8 ? 1 : 4;
    if (chunk_rows <= 0) chunk_rows = (int64_t)1 << 22;
    if (chunk_rows > n) chunk_rows = n;
    // large byte chunks go through the level-synchronous descent (same ids, far less DRAM traffic);
    // 128-byte rows and k <= 16: distance step on the INT8 tensor pipe (planes prepared once per call)
    const bool sorted = dtype == VT_U8 && depth >= 2 && chunk_rows >= ((int64_t)1 << 18) &&
                        path_bits(k) <= 8 && (depth - 1) * path_bits(k) <= 32;
    const bool mma = sorted && dp == 128 && k <= 16 && getenv("VT_NO_MMA") == nullptr;
    int rc = host_pipe_prepare(chunk_rows, dp * esz, sorted ? vt_quantize_sorted_work_bytes(chunk_rows) : 0,
                               mma ? vt_mma_prepared_bytes(k, depth) : 0);
    if (rc != VT_OK) return rc;
    HostPipe &hp = g_pipe;
    bool use_mma = false;
    if (mma) {
        // the tensor path needs every child centroid inside the 8.16 fixed-point range: the prepare
        // kernel counts violations and the chunks fall back to the FP32 kernels if there is one
        unsigned long long h_flag[4] = {0, 0, 0, 0};
        if (cudaMemsetAsync(hp.d_flag, 0, sizeof h_flag, hp.st[0]) != cudaSuccess) rc = VT_ECUDA;
        if (rc == VT_OK) rc = vt_mma_prepare(d_centroids, k, depth, hp.prepared, (uint64_t *)hp.d_flag, hp.st[0]);
        if (rc == VT_OK && (cudaMemcpyAsync(h_flag, hp.d_flag, sizeof h_flag, cudaMemcpyDeviceToHost, hp.st[0]) != cudaSuccess ||
                            cudaStreamSynchronize(hp.st[0]) != cudaSuccess)) rc = VT_ECUDA;
        use_mma = rc == VT_OK && h_flag[1] == 0;
    }
    int64_t done = 0;
    for (int c = 0; done < n && rc == VT_OK; ++c, done += chunk_rows) {
        const int s = c % HP_NS;
        const int64_t rows = (n - done) < chunk_rows ? (n - done) : chunk_rows;
        const char *src = (const char *)h_desc + (size_t)(done * dp * esz);
        if (cudaMemcpyAsync(hp.dbuf[s], src, (size_t)(rows * dp * esz), cudaMemcpyHostToDevice, hp.st[s]) != cudaSuccess) {
            vt_set_error("vt_quantize_host: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = VT_ECUDA;
            break;
        }
        if (use_mma && rows >= ((int64_t)1 << 16))
            rc = vt_mma_levels(hp.dbuf[s], rows, d_centroids, d_internal, d_ref_id, k, depth, nullptr, hp.wbuf[s], nullptr,
                               hp.prepared, hp.work[s], hp.st[s], nullptr, nullptr);
        else if (sorted && rows >= ((int64_t)1 << 16))
            rc = vt_quantize_sorted_impl(hp.dbuf[s], rows, dp, d_centroids, d_internal, d_ref_id, k, depth, nullptr,
                                         hp.wbuf[s], nullptr, hp.work[s], hp.st[s], nullptr);
        else
            rc = vt_quantize_dispatch(hp.dbuf[s], dtype, rows, dp, d_centroids, d_internal, d_ref_id, k, depth, 0, nullptr,
                                      hp.wbuf[s], nullptr, hp.st[s]);
        if (rc != VT_OK) break;
        if (cudaMemcpyAsync(h_word + done, hp.wbuf[s], (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToHost, hp.st[s]) !=
            cudaSuccess) {
            vt_set_error("vt_quantize_host: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = VT_ECUDA;
        }
    }
    for (int s = 0; s < HP_NS; ++s)
        if (cudaStreamSynchronize(hp.st[s]) != cudaSuccess && rc == VT_OK) {
            vt_set_error("vt_quantize_host: stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = VT_ECUDA;
        }
    return rc;
}
