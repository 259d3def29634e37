// vt_weight.cu -- TF-IDF weighting, L1/L2 normalisation and the weighted inverted-file scorer, all in
// binary64 (the weighting cbir/encoders/vocabulary.py:131,137 leaves commented out: w_i = ln(N / N_i) of
// Nister & Stewenius; opt-in, the reference's own scoring is plain tf -> vt_score.cu).
//
// Nothing weighted is stored: the inverted file keeps the exact integer tf postings of the tf mode and the
// weights enter on the query side,
//   L2:  key(q, d) = sum_n (idf_n^2 tf_q[n] / |w_q|_2) * tf_d[n]  *  1 / |w_d|_2          score = sqrt(2 - 2 key)
//   L1:  key(q, d) = sum_{n common} (q_n + d_n - |q_n - d_n|),  q_n = idf_n tf_q[n] / |w_q|_1,
//                                                               d_n = idf_n tf_d[n] / |w_d|_1      score = 2 - key
// with per-image norms |w_d| computed once at index time.  Products of a binary64 weight and an integer count
// accumulate in binary64 shared-memory cells, so keys carry ~1e-15 relative noise instead of the ~1e-7 of
// float32 weights (which sqrt(2 - 2 key) amplifies near 0).  One index serves tf, tf-idf/L2 and tf-idf/L1.
#include "vt_common.cuh"

namespace {
constexpr int WS_THREADS = 256;
constexpr int WS_WARPS = WS_THREADS / 32;
constexpr int WS_MAXK = 64;
constexpr uint32_t WS_LONG = 24;

__global__ void idf_kernel(const int32_t *__restrict__ df, int64_t n_ref, double n_total, double *__restrict__ idf)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_ref; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t f = df[t];
        idf[t] = f > 0 ? log(n_total / (double)f) : 0.0;         // nodes no image visits: weight 0
    }
}

// warp per image: |w|_2 or |w|_1 of w_n = idf_n tf_n, fixed summation order (lane-strided, then butterfly)
__global__ void weighted_rnorm_kernel(const int64_t *__restrict__ row_off, const int32_t *__restrict__ node,
                                      const int32_t *__restrict__ count, int64_t n_img, const double *__restrict__ idf,
                                      int l1, double *__restrict__ rnorm)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t img = warp; img < n_img; img += n_warps) {
        const int64_t b = row_off[img], e = row_off[img + 1];
        double s = 0.0;
        for (int64_t t = b + lane; t < e; t += 32) {
            const double w = idf[node[t]] * (double)count[t];
            s += l1 ? fabs(w) : w * w;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(VT_FULL, s, o);
        if (lane == 0) rnorm[img] = s > 0.0 ? (l1 ? 1.0 / s : 1.0 / sqrt(s)) : -1.0;   // all-zero vector: 0/0 -> NaN -> 1e6
    }
}

// per query entry: the value that multiplies tf_d in the L2 key, or (q_n, idf_n) for the L1 key
__global__ void query_weight_kernel(const int64_t *__restrict__ row_off, const int32_t *__restrict__ node,
                                    const int32_t *__restrict__ count, int64_t n_q, const double *__restrict__ idf,
                                    const double *__restrict__ q_rnorm, int l1, double *__restrict__ qv,
                                    double *__restrict__ qi)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp; q < n_q; q += n_warps) {
        const int64_t b = row_off[q], e = row_off[q + 1];
        const double rn = q_rnorm[q] > 0.0 ? q_rnorm[q] : 0.0;
        for (int64_t t = b + lane; t < e; t += 32) {
            const double w = idf[node[t]];
            const double x = w * (double)count[t] * rn;
            qv[t] = l1 ? x : x * w;
            qi[t] = w;
        }
    }
}

__device__ __forceinline__ bool better(double ka, int32_t ia, double kb, int32_t ib)
{
    return ka > kb || (ka == kb && ia < ib);
}

// posting = (tf << 13) | local image index (the packed tf postings of vt_score.cu)
template <int L1>
__device__ __forceinline__ void accumulate(double *s_acc, const double *s_rn, const uint32_t pe, double qv, double qi)
{
    const uint32_t img = pe & 8191u;
    const double tf = (double)(pe >> 13);
    if (L1) {
        const double d = qi * tf * s_rn[img];
        atomicAdd(&s_acc[img], qv + d - fabs(qv - d));
    } else {
        atomicAdd(&s_acc[img], qv * tf);
    }
}

// CTA per (query, shard of R images); lanes own query words, long posting runs go to the whole warp
template <int L1, int R>
__global__ void __launch_bounds__(WS_THREADS)
score_shard_w_kernel(const uint32_t *__restrict__ post_off, const uint32_t *__restrict__ postings,
                     const double *__restrict__ img_rnorm, int64_t n_img, int n_shards,
                     const int64_t *__restrict__ q_off, const int32_t *__restrict__ q_node, const double *__restrict__ qv,
                     const double *__restrict__ qi, const double *__restrict__ q_rnorm, int64_t n_query, int topk,
                     double *__restrict__ out_key, int32_t *__restrict__ out_id, double *__restrict__ all_key)
{
    constexpr int PER = R / WS_THREADS;
    extern __shared__ unsigned char s_raw[];
    double *s_acc = reinterpret_cast<double *>(s_raw);                         // [R] sums, then rank keys
    double *s_rn = s_acc + R;                                                  // [R] 1 / |w_d|_1 (L1 only)
    __shared__ double r_key[WS_WARPS];
    __shared__ int32_t r_id[WS_WARPS];
    __shared__ int r_owner[WS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int64_t pair = blockIdx.x; pair < n_query * n_shards; pair += gridDim.x) {
        const int64_t q = pair / n_shards;
        const int shard = (int)(pair - q * n_shards);
        const int64_t img0 = (int64_t)shard * R;
        const int n_here = (int)((n_img - img0) < R ? (n_img - img0) : R);
        for (int t = threadIdx.x; t < R; t += WS_THREADS) {
            s_acc[t] = 0.0;
            if (L1) s_rn[t] = t < n_here ? fmax(img_rnorm[img0 + t], 0.0) : 0.0;
        }
        __syncthreads();
        const int64_t qb = q_off[q], qe = q_off[q + 1];
        const uint32_t *off = post_off + shard;                 // row of (node, shard) = node * n_shards + shard
        for (int64_t w0 = qb + (int64_t)warp * 32; w0 < qe; w0 += WS_THREADS) {
            const int64_t w = w0 + lane;
            const int32_t nd = w < qe ? q_node[w] : -1;
            const double v = w < qe ? qv[w] : 0.0, vi = w < qe ? qi[w] : 0.0;
            const uint32_t b = nd >= 0 ? off[(size_t)nd * n_shards] : 0u;
            const uint32_t e = nd >= 0 ? off[(size_t)nd * n_shards + 1] : 0u;
            const bool is_long = (e - b) > WS_LONG;
            if (!is_long)
                for (uint32_t p = b; p < e; ++p) accumulate<L1>(s_acc, s_rn, postings[p], v, vi);
            unsigned int m = __ballot_sync(VT_FULL, is_long);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t lb = __shfl_sync(VT_FULL, b, src), le = __shfl_sync(VT_FULL, e, src);
                const double lv = __shfl_sync(VT_FULL, v, src), li = __shfl_sync(VT_FULL, vi, src);
                for (uint32_t p = lb + lane; p < le; p += 32) accumulate<L1>(s_acc, s_rn, postings[p], lv, li);
            }
        }
        __syncthreads();
        // rank keys (in place) + per-thread best; an empty query or image has key -1 (NaN -> 1e6 in the reference)
        const bool q_ok = q_rnorm[q] > 0.0;
        double bkey = -2.0; int32_t bid = 0x7fffffff;
#pragma unroll 4
        for (int u = 0; u < PER; ++u) {
            const int t = u * WS_THREADS + threadIdx.x;
            double key = -2.0;                                                  // padding slots never win
            if (t < n_here) {
                const double rn = img_rnorm[img0 + t];
                key = (rn < 0.0 || !q_ok) ? -1.0 : (L1 ? s_acc[t] : s_acc[t] * rn);
                if (all_key) all_key[q * n_img + img0 + t] = key;
            }
            s_acc[t] = key;
            if (better(key, (int32_t)(img0 + t), bkey, bid)) { bkey = key; bid = (int32_t)(img0 + t); }
        }
        __syncthreads();
        // k rounds of block argmax; only the winning thread rescans its slice
        for (int r = 0; r < topk; ++r) {
            double key = bkey; int32_t id = bid; int owner = threadIdx.x;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const double k2 = __shfl_xor_sync(VT_FULL, key, o);
                const int32_t i2 = __shfl_xor_sync(VT_FULL, id, o);
                const int o2 = __shfl_xor_sync(VT_FULL, owner, o);
                if (better(k2, i2, key, id)) { key = k2; id = i2; owner = o2; }
            }
            if (lane == 0) { r_key[warp] = key; r_id[warp] = id; r_owner[warp] = owner; }
            __syncthreads();
            key = r_key[0]; id = r_id[0]; owner = r_owner[0];
#pragma unroll
            for (int w = 1; w < WS_WARPS; ++w)
                if (better(r_key[w], r_id[w], key, id)) { key = r_key[w]; id = r_id[w]; owner = r_owner[w]; }
            __syncthreads();
            if (threadIdx.x == 0) {
                const size_t o = ((size_t)q * n_shards + shard) * topk + r;
                out_key[o] = key;
                out_id[o] = key > -2.0 ? id : -1;
            }
            if (threadIdx.x == owner) {
                if (key > -2.0) s_acc[id - img0] = -2.0;
                bkey = -2.0; bid = 0x7fffffff;
                for (int u = 0; u < PER; ++u) {
                    const int t = u * WS_THREADS + threadIdx.x;
                    const double k2 = s_acc[t];
                    if (better(k2, (int32_t)(img0 + t), bkey, bid)) { bkey = k2; bid = (int32_t)(img0 + t); }
                }
            }
        }
        __syncthreads();
    }
}
}  // namespace

// idf_n = ln(N / df_n) in binary64 (0 for nodes no image visits)
extern "C" int vt_idf(const int32_t *d_df, int64_t n_ref, int64_t n_total, double *d_idf, void *stream)
{
    VT_CHECK_ARG(d_df && d_idf && n_ref >= 0 && n_total >= 0, "bad arguments");
    if (n_ref == 0) return VT_OK;
    const int64_t want = (n_ref + 255) / 256, cap = (int64_t)vt_sm_count() * 8;
    idf_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(d_df, n_ref, (double)(n_total > 0 ? n_total : 1), d_idf);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// 1 / |idf * tf|_2 (norm_l1 == 0) or 1 / |idf * tf|_1 per CSR row; -1 for all-zero rows
extern "C" int vt_weighted_rnorm(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_img,
                                 const double *d_idf, int norm_l1, double *d_rnorm, void *stream)
{
    VT_CHECK_ARG(d_row_off && d_idf && d_rnorm && n_img >= 0, "bad arguments");
    if (n_img == 0) return VT_OK;
    const int64_t want = (n_img * 32 + 255) / 256, cap = (int64_t)vt_sm_count() * 16;
    weighted_rnorm_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(d_row_off, d_node, d_count, n_img,
                                                                                            d_idf, norm_l1, d_rnorm);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// query-side factors of the weighted keys (see the header comment); d_q_rnorm from vt_weighted_rnorm on the queries
extern "C" int vt_query_weights(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_query,
                                const double *d_idf, const double *d_q_rnorm, int norm_l1, double *d_qv, double *d_qi,
                                void *stream)
{
    VT_CHECK_ARG(d_row_off && d_idf && d_q_rnorm && d_qv && d_qi && n_query >= 0, "bad arguments");
    if (n_query == 0) return VT_OK;
    const int64_t want = (n_query * 32 + 255) / 256, cap = (int64_t)vt_sm_count() * 16;
    query_weight_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(d_row_off, d_node, d_count, n_query,
                                                                                          d_idf, d_q_rnorm, norm_l1, d_qv, d_qi);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// weighted scoring over the TF inverted file (d_postings: the packed 4-byte tf postings, (tf << 13) | local image);
// per-shard top-k lists like vt_score_topk, merged by vt_topk_merge with mode VT_SCORE_W_L2 / VT_SCORE_W_L1
extern "C" int vt_score_topk_weighted(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                                      int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                                      const double *d_qv, const double *d_qi, const double *d_q_rnorm, int64_t n_query,
                                      int mode, int topk, double *d_shard_key, int32_t *d_shard_id, double *d_all_key,
                                      void *stream)
{
    VT_CHECK_ARG(n_img > 0 && n_ref > 0 && n_query >= 0, "bad sizes");
    VT_CHECK_ARG(topk >= 1 && topk <= WS_MAXK, "topk out of range");
    VT_CHECK_ARG(mode == VT_SCORE_W_L2 || mode == VT_SCORE_W_L1, "mode must be VT_SCORE_W_L2 or VT_SCORE_W_L1");
    VT_CHECK_ARG(d_post_off && d_postings && d_img_rnorm && d_qv && d_qi && d_q_rnorm, "null buffer");
    if (n_query == 0) return VT_OK;
    constexpr int R = 8192;
    const int n_shards = (int)((n_img + R - 1) / R);
    const int64_t pairs = n_query * n_shards, cap = (int64_t)vt_sm_count() * 8;
    const int blocks = (int)(pairs < cap ? pairs : cap);
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == VT_SCORE_W_L1) {
        auto kern = score_shard_w_kernel<1, R>;
        const size_t smem = (size_t)R * 16;
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, WS_THREADS, smem, s>>>(d_post_off, (const uint32_t *)d_postings, d_img_rnorm, n_img, n_shards, d_q_off,
                                              d_q_node, d_qv, d_qi, d_q_rnorm, n_query, topk, d_shard_key, d_shard_id, d_all_key);
    } else {
        auto kern = score_shard_w_kernel<0, R>;
        const size_t smem = (size_t)R * 8;
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, WS_THREADS, smem, s>>>(d_post_off, (const uint32_t *)d_postings, d_img_rnorm, n_img, n_shards, d_q_off,
                                              d_q_node, d_qv, d_qi, d_q_rnorm, n_query, topk, d_shard_key, d_shard_id, d_all_key);
    }
    VT_LAUNCH_CHECK();
    return VT_OK;
}
