// vt_kmeans.cu -- level-synchronous hierarchical k-means (VocabularyTree.fit,
// cbir/encoders/vocabulary.py:36-76; the per-node k-means at :62-63 is the seeded exact Lloyd
// of DESIGN.md section 4 instead of sklearn's unseeded MiniBatchKMeans).
//
// All nodes of a tree level are built by the same launches.  Members are addressed through a
// permutation grouped by node (stable), so a warp reads 4 member rows and the 10 child centroids
// of each member's node (5 KB, L1/L2 resident and shared by neighbouring members).  Centroid
// sums are exact integers (uint64), which makes the result independent of summation order and of
// the number of GPUs (sums are all-reduced between vt_kmeans_assign and vt_kmeans_update).
#include "vt_common.cuh"

namespace {

// ---- initial centres ---------------------------------------------------------------------------
template <typename T>
__global__ void kmeans_init_kernel(const T *__restrict__ desc, int dp, const int32_t *__restrict__ perm,
                                   const int64_t *__restrict__ seg_begin, const int64_t *__restrict__ seg_count,
                                   const int64_t *__restrict__ seg_count_global,
                                   const int64_t *__restrict__ seg_rank0, int64_t n_nodes, int k,
                                   float *__restrict__ child_cent)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t idx = warp; idx < n_nodes * k; idx += nwarps) {
        const int64_t j = idx / k;
        const int c = (int)(idx - j * k);
        const int64_t m_local = seg_count[j];
        const int64_t m = seg_count_global ? seg_count_global[j] : m_local;
        if (m < k) continue;                                    // leaf: vocabulary.py:55
        const int64_t r = ((int64_t)c * m) / k - (seg_rank0 ? seg_rank0[j] : 0);
        if (r < 0 || r >= m_local) continue;                    // another rank owns that member
        const T *row = desc + (size_t)perm[seg_begin[j] + r] * dp;
        float *out = child_cent + (size_t)idx * dp;
        for (int d = lane; d < dp; d += 32) out[d] = (float)row[d];
    }
}

// ---- assignment + exact accumulation -----------------------------------------------------------
// SMEM_ACC: the whole level has few child rows (top levels) -> per-block uint32 partial sums in
// shared memory, flushed with one uint64 atomic per touched cell.  Otherwise uint64 atomics go
// straight to L2 (deep levels: ~10^5..10^6 rows, negligible contention).
template <typename T, int DP, bool SMEM_ACC>
__global__ void __launch_bounds__(256, 3)
kmeans_assign_kernel(const T *__restrict__ desc, const int32_t *__restrict__ perm, const int32_t *__restrict__ parent,
                     int64_t n_members, const uint8_t *__restrict__ node_internal, int k, int n_rows,
                     const float *__restrict__ child_cent, uint8_t *__restrict__ label,
                     unsigned long long *__restrict__ sums, unsigned long long *__restrict__ counts,
                     unsigned long long *__restrict__ changed, unsigned long long *__restrict__ stats)
{
    constexpr int VEC = DP / 32;
    extern __shared__ uint32_t s_acc[];                      // [n_rows*DP] sums then [n_rows] counts
    if (SMEM_ACC) {
        for (int t = threadIdx.x; t < n_rows * (DP + 1); t += blockDim.x) s_acc[t] = 0u;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, j = lane & 7;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned int rechecks = 0, n_changed = 0;

    for (int64_t base = warp * 4; base < n_members; base += nwarps * 4) {
        const int64_t p = base + g;
        const bool valid = p < n_members;
        const int32_t jp = valid ? parent[p] : 0;
        const bool active = valid && node_internal[jp];
        const T *xrow = desc + (size_t)(valid ? perm[p] : 0) * DP;
        float4 x[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[v] = Elem<T>::load4(xrow, j + 8 * v);
        const float *children = child_cent + (size_t)jp * k * DP;
        const int c = descend_one_level<T, DP>(x, xrow, children, k, active, lane, rechecks);
        if (active) {
            const int64_t row = (int64_t)jp * k + c;
            if (j == 0) {
                if (label[p] != (uint8_t)c) { ++n_changed; label[p] = (uint8_t)c; }
                if (SMEM_ACC) atomicAdd(&s_acc[n_rows * DP + row], 1u);
                else atomicAdd(&counts[row], 1ull);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int d0 = 4 * (j + 8 * v);
                const float xv[4] = {x[v].x, x[v].y, x[v].z, x[v].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned int q = (unsigned int)xv[e];
                    if (q) {
                        if (SMEM_ACC) atomicAdd(&s_acc[row * DP + d0 + e], q);
                        else atomicAdd(&sums[row * DP + d0 + e], (unsigned long long)q);
                    }
                }
            }
        }
    }
    if (SMEM_ACC) {
        __syncthreads();
        for (int t = threadIdx.x; t < n_rows * DP; t += blockDim.x)
            if (s_acc[t]) atomicAdd(&sums[t], (unsigned long long)s_acc[t]);
        for (int t = threadIdx.x; t < n_rows; t += blockDim.x)
            if (s_acc[n_rows * DP + t]) atomicAdd(&counts[t], (unsigned long long)s_acc[n_rows * DP + t]);
    }
    // warp-aggregate the two statistics
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
        n_changed += __shfl_xor_sync(VT_FULL, n_changed, o);
    }
    if (lane == 0) {
        if (n_changed && changed) atomicAdd(changed, (unsigned long long)n_changed);
        if (rechecks && stats) atomicAdd(stats, (unsigned long long)rechecks);
    }
}

__global__ void kmeans_update_kernel(const unsigned long long *__restrict__ sums,
                                     const unsigned long long *__restrict__ counts, int64_t n_rows, int dp,
                                     float *__restrict__ cent)
{
    const int64_t total = n_rows * dp;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long cnt = counts[t / dp];
        if (cnt) cent[t] = __double2float_rn(__ddiv_rn((double)sums[t], (double)cnt));   // empty cluster keeps its centre
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
column_sums_kernel(const T *__restrict__ desc, int64_t n, int dp, unsigned long long *__restrict__ sums)
{
    // thread t owns column t % dp of rows (t / dp), strided; dp divides 256 (dp in {32,64,128})
    const int col = threadIdx.x % dp;
    const int rows_per_pass = blockDim.x / dp;
    const int rsub = threadIdx.x / dp;
    unsigned long long acc = 0;
    for (int64_t r = (int64_t)blockIdx.x * rows_per_pass + rsub; r < n; r += (int64_t)gridDim.x * rows_per_pass)
        acc += (unsigned long long)desc[(size_t)r * dp + col];
    if (acc) atomicAdd(&sums[col], acc);
}

__global__ void partition_keys_kernel(const int32_t *__restrict__ parent, const uint8_t *__restrict__ label,
                                      int64_t n, const uint8_t *__restrict__ node_internal, int k, uint32_t sentinel,
                                      uint32_t *__restrict__ keys)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t jp = parent[p];
        keys[p] = node_internal[jp] ? (uint32_t)jp * (uint32_t)k + label[p] : sentinel;
    }
}

inline int grid_for(int64_t work_items, int per_block, int cap_per_sm)
{
    const int64_t want = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)vt_sm_count() * cap_per_sm;
    const int64_t g = want < cap ? want : cap;
    return (int)(g > 0 ? g : 1);
}
}  // namespace

extern "C" int vt_kmeans_init(const void *d_desc, int dtype, int dp, const int32_t *d_perm, const int64_t *d_seg_begin,
                              const int64_t *d_seg_count, const int64_t *d_seg_count_global, const int64_t *d_seg_rank0,
                              int64_t n_nodes, int k, float *d_child_centroids, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8, "tree build takes byte descriptors (stage float input with vt_stage_f32_to_u8)");
    VT_CHECK_ARG(dp == 32 || dp == 64 || dp == 128, "dp must be 32, 64 or 128");
    VT_CHECK_ARG(n_nodes >= 0 && k >= 1 && k <= 255, "bad n_nodes / k");
    if (n_nodes == 0) return VT_OK;
    kmeans_init_kernel<uint8_t><<<grid_for(n_nodes * k, 8, 16), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t *)d_desc, dp, d_perm, d_seg_begin, d_seg_count, d_seg_count_global, d_seg_rank0, n_nodes, k,
        d_child_centroids);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

template <int DP>
static int launch_assign(const void *d_desc, const int32_t *perm, const int32_t *parent, int64_t n_members,
                         const uint8_t *node_internal, int64_t n_nodes, int k, const float *cent, uint8_t *label,
                         uint64_t *sums, uint64_t *counts, uint64_t *changed, uint64_t *stats, cudaStream_t s)
{
    const int64_t n_rows = n_nodes * k;
    const size_t smem = (size_t)n_rows * (DP + 1) * sizeof(uint32_t);
    // uint32 partial sums hold 255 * members_per_block: keep members_per_block below 2^24
    const bool use_smem = smem <= 64 * 1024 && n_members / ((int64_t)vt_sm_count() * 2) < ((int64_t)1 << 24);
    if (use_smem) {
        auto kern = kmeans_assign_kernel<uint8_t, DP, true>;
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        const int blocks = grid_for(n_members, 32, 2) < vt_sm_count() * 2 ? grid_for(n_members, 32, 2) : vt_sm_count() * 2;
        kern<<<blocks, 256, smem, s>>>((const uint8_t *)d_desc, perm, parent, n_members, node_internal, k, (int)n_rows,
                                       cent, label, (unsigned long long *)sums, (unsigned long long *)counts,
                                       (unsigned long long *)changed, (unsigned long long *)stats);
    } else {
        kmeans_assign_kernel<uint8_t, DP, false><<<grid_for(n_members, 32, 6), 256, 0, s>>>(
            (const uint8_t *)d_desc, perm, parent, n_members, node_internal, k, (int)n_rows, cent, label,
            (unsigned long long *)sums, (unsigned long long *)counts, (unsigned long long *)changed,
            (unsigned long long *)stats);
    }
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_kmeans_assign(const void *d_desc, int dtype, int dp, const int32_t *d_perm, const int32_t *d_parent,
                                int64_t n_members, const uint8_t *d_node_internal, int64_t n_nodes, int k,
                                const float *d_child_centroids, uint8_t *d_label, uint64_t *d_sums, uint64_t *d_counts,
                                uint64_t *d_changed, uint64_t *d_stats, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8, "tree build takes byte descriptors (stage float input with vt_stage_f32_to_u8)");
    VT_CHECK_ARG(n_members >= 0 && n_nodes >= 1 && k >= 1 && k <= 255, "bad sizes");
    VT_CHECK_ARG(n_nodes * k < ((int64_t)1 << 31), "too many child rows for one level");
    if (n_members == 0) return VT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dp) {
    case 32: return launch_assign<32>(d_desc, d_perm, d_parent, n_members, d_node_internal, n_nodes, k, d_child_centroids, d_label, d_sums, d_counts, d_changed, d_stats, s);
    case 64: return launch_assign<64>(d_desc, d_perm, d_parent, n_members, d_node_internal, n_nodes, k, d_child_centroids, d_label, d_sums, d_counts, d_changed, d_stats, s);
    case 128: return launch_assign<128>(d_desc, d_perm, d_parent, n_members, d_node_internal, n_nodes, k, d_child_centroids, d_label, d_sums, d_counts, d_changed, d_stats, s);
    }
    vt_set_error("vt_kmeans_assign: dp must be 32, 64 or 128 (got %d)", dp);
    return VT_EUNSUPPORTED;
}

extern "C" int vt_kmeans_update(const uint64_t *d_sums, const uint64_t *d_counts, int64_t n_rows, int dp,
                                float *d_child_centroids, void *stream)
{
    VT_CHECK_ARG(n_rows >= 0 && dp > 0, "bad shape");
    if (n_rows == 0) return VT_OK;
    kmeans_update_kernel<<<grid_for(n_rows * dp, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)d_sums, (const unsigned long long *)d_counts, n_rows, dp, d_child_centroids);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_column_sums(const void *d_desc, int dtype, int64_t n, int dp, uint64_t *d_sums, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8, "byte descriptors only");
    VT_CHECK_ARG(dp == 32 || dp == 64 || dp == 128, "dp must be 32, 64 or 128");
    if (n <= 0) return VT_OK;
    column_sums_kernel<uint8_t><<<grid_for(n, 256 / dp * 64, 8), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t *)d_desc, n, dp, (unsigned long long *)d_sums);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_partition_keys(const int32_t *d_parent, const uint8_t *d_label, int64_t n_members,
                                 const uint8_t *d_node_internal, int k, uint32_t sentinel, uint32_t *d_keys, void *stream)
{
    if (n_members <= 0) return VT_OK;
    partition_keys_kernel<<<grid_for(n_members, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        d_parent, d_label, n_members, d_node_internal, k, sentinel, d_keys);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
