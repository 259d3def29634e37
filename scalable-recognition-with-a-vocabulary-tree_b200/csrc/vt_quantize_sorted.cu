// vt_quantize_sorted.cu -- level-synchronous greedy descent for LARGE batches
// (VocabularyTree.propagate_feature, cbir/encoders/vocabulary.py:104-124).
//
// Why a second kernel: the fused warp-per-descriptor descent (vt_quantize.cu) gathers 10 x 512 B of
// child centroids per descriptor per level; below level 4 of a million-leaf tree those gathers
// miss L2 (ncu, profiles/r01_*: 62 GB DRAM read for a 1 GB batch).  Here every level is one pass
// over the batch SORTED BY CURRENT NODE, so the 32 descriptors of a warp share their node's 5 KB of
// child centroids: one lane owns one descriptor (its 128 bytes live in 32 registers), the
// centroid float4 is a warp-uniform load served by L1 as a broadcast, and each loaded value feeds
// 32 descriptors.  Between levels the (node, row) pairs are re-sorted with the stable radix sort
// (vt_sort.cu).  DRAM traffic per level: 128 B row gather + 12 B of keys/permutation per
// descriptor + the level's centroids once.
//
// Results are bit-identical to vt_quantize (same canonical contract: float32 pass, margin check,
// binary64 re-evaluation of contested rows).
#include "vt_common.cuh"

int vt_sort_pairs_impl(uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_tmp_keys, uint32_t *d_tmp_vals, int64_t n,
                       int n_bits, void *d_hist, int *result_in_tmp, cudaStream_t s);

int vt_quantize_sorted_impl(const void *d_desc, int64_t n, int dp, const float *d_centroids, const uint8_t *d_internal,
                            const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap, int32_t *d_word,
                            uint64_t *d_stats, void *d_work, cudaStream_t s, cudaEvent_t *ev);

namespace {
constexpr int QS_CH = 10;          // children evaluated per register chunk
constexpr int QS_R = 4;            // descriptors per lane (register tile: one centroid load feeds 32*R rows)
constexpr int QS_PW = 4;           // packed words (16 dimensions) consumed per unrolled part

// float32 squared distances of ONE descriptor (packed bytes in global memory) to up to QS_CH child
// rows starting at `children` -- generic path, per-lane centroid pointer
template <int DP>
__device__ __forceinline__ void dist_one(const uint8_t *__restrict__ xrow, const float *__restrict__ children, int nc,
                                         bool on, float (&acc)[QS_CH])
{
    constexpr int NW = DP / 4;
#pragma unroll
    for (int u = 0; u < QS_CH; ++u) acc[u] = 0.f;
    for (int part = 0; part < NW / QS_PW; ++part) {
        const uint4 q = on ? __ldg(reinterpret_cast<const uint4 *>(xrow) + part) : make_uint4(0, 0, 0, 0);
        const uint32_t xp[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int w = 0; w < QS_PW; ++w) {
            const uint32_t b = xp[w];
            const float x0 = (float)(b & 0xffu), x1 = (float)((b >> 8) & 0xffu), x2 = (float)((b >> 16) & 0xffu),
                        x3 = (float)(b >> 24);
#pragma unroll
            for (int u = 0; u < QS_CH; ++u) {
                if (u < nc) {
                    const float4 c = __ldg(reinterpret_cast<const float4 *>(children + (size_t)u * DP) + part * QS_PW + w);
                    float d;
                    d = c.x - x0; acc[u] = fmaf(d, d, acc[u]);
                    d = c.y - x1; acc[u] = fmaf(d, d, acc[u]);
                    d = c.z - x2; acc[u] = fmaf(d, d, acc[u]);
                    d = c.w - x3; acc[u] = fmaf(d, d, acc[u]);
                }
            }
        }
    }
}

// float32 squared distances of the R descriptors of a lane to the nc (<= QS_CH) child rows at
// `children` (warp-uniform pointer): every centroid float4 is loaded once and used for R rows.
template <int DP, bool FULL>
__device__ __forceinline__ void tile_dists(const uint8_t *const (&xrow)[QS_R], const bool (&on)[QS_R],
                                           const float *__restrict__ children, int nc, float (&acc)[QS_R][QS_CH])
{
    constexpr int NW = DP / 4;
    constexpr int R = QS_R;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u < QS_CH; ++u) acc[r][u] = 0.f;
    uint4 q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] = on[r] ? __ldg(reinterpret_cast<const uint4 *>(xrow[r])) : make_uint4(0, 0, 0, 0);
    for (int part = 0; part < NW / QS_PW; ++part) {
        uint32_t xp[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { xp[r][0] = q[r].x; xp[r][1] = q[r].y; xp[r][2] = q[r].z; xp[r][3] = q[r].w; }
        if (part + 1 < NW / QS_PW) {                                // software pipeline: next 16 bytes of every row
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (on[r]) q[r] = __ldg(reinterpret_cast<const uint4 *>(xrow[r]) + part + 1);
        }
        const float *cpart = children + part * (QS_PW * 4);
#pragma unroll
        for (int w = 0; w < QS_PW; ++w) {
            float xf[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t b = xp[r][w];
                xf[r][0] = (float)(b & 0xffu); xf[r][1] = (float)((b >> 8) & 0xffu);
                xf[r][2] = (float)((b >> 16) & 0xffu); xf[r][3] = (float)(b >> 24);
            }
#pragma unroll
            for (int u = 0; u < QS_CH; ++u) {
                if (FULL || u < nc) {                                  // warp-uniform
                    const float4 c = __ldg(reinterpret_cast<const float4 *>(cpart + (size_t)u * DP) + w);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float d;
                        d = c.x - xf[r][0]; acc[r][u] = fmaf(d, d, acc[r][u]);
                        d = c.y - xf[r][1]; acc[r][u] = fmaf(d, d, acc[r][u]);
                        d = c.z - xf[r][2]; acc[r][u] = fmaf(d, d, acc[r][u]);
                        d = c.w - xf[r][3]; acc[r][u] = fmaf(d, d, acc[r][u]);
                    }
                }
            }
        }
    }
}

template <int DP>
__global__ void __launch_bounds__(256, 2)
level_kernel(const uint8_t *__restrict__ desc, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ parent,
             int64_t n, uint32_t n_nodes, int64_t level_off, int64_t n_heap, int k, int is_last,
             const float *__restrict__ cent, const uint8_t *__restrict__ internal, const int32_t *__restrict__ ref_id,
             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ perm_out, int32_t *__restrict__ leaf_out,
             int32_t *__restrict__ word_out, unsigned long long *__restrict__ stats)
{
    constexpr int NW = DP / 4;                                   // packed 4-byte words per row
    constexpr int R = QS_R;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t sentinel = n_nodes * (uint32_t)k;
    const float inf = __int_as_float(0x7f800000);
    unsigned int rechecks = 0;

    for (int64_t base = warp * (32 * R); base < n; base += nwarps * (32 * R)) {
        uint32_t row[R], jp[R];
        int64_t node[R];
        bool valid[R], is_int[R];
        const uint8_t *xrow[R];
        bool any_int = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t p = base + r * 32 + lane;
            valid[r] = p < n;
            row[r] = valid[r] ? (perm ? perm[p] : (uint32_t)p) : 0u;
            jp[r] = valid[r] ? (parent ? parent[p] : 0u) : sentinel;
            const bool alive = valid[r] && jp[r] < n_nodes;      // not yet stopped at a leaf
            node[r] = level_off + (alive ? jp[r] : 0u);
            is_int[r] = alive && (node[r] * k + k < n_heap) && internal[node[r]];
            xrow[r] = desc + (size_t)row[r] * DP;
            if (alive && !is_int[r]) {                           // ragged tree: this node is a leaf
                if (leaf_out) leaf_out[row[r]] = (int32_t)node[r];
                if (word_out) word_out[row[r]] = ref_id[node[r]];
            }
            if (is_int[r]) {                                     // pull the row towards the SM early
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow[r]));
                if (DP > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow[r] + 32));
                if (DP > 64) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow[r] + 64));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow[r] + 96));
                }
            }
            any_int |= is_int[r];
        }
        int arg[R];
        bool contested[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { arg[r] = 0; contested[r] = false; }

        if (__any_sync(VT_FULL, any_int)) {
            // warp-uniform node?  (the batch is sorted by node, so this is the common case)
            int64_t mine = -1;
#pragma unroll
            for (int r = 0; r < R; ++r) if (is_int[r] && mine < 0) mine = node[r];
            const unsigned int have = __ballot_sync(VT_FULL, mine >= 0);
            const int64_t n0 = __shfl_sync(VT_FULL, mine, __ffs(have) - 1);
            bool same = true;
#pragma unroll
            for (int r = 0; r < R; ++r) same = same && (!is_int[r] || node[r] == n0);
            const bool uniform = __all_sync(VT_FULL, same);

            float best[R], second[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { best[r] = inf; second[r] = inf; }

            if (uniform) {
                // ---- register-tiled fast path: one centroid float4 load feeds 32*R descriptors ----
                const float *children = cent + (size_t)(n0 * k + 1) * DP;
                for (int c0 = 0; c0 < k; c0 += QS_CH) {
                    const int nc = k - c0 < QS_CH ? k - c0 : QS_CH;
                    float acc[R][QS_CH];
                    if (nc == QS_CH) tile_dists<DP, true>(xrow, is_int, children + (size_t)c0 * DP, nc, acc);
                    else tile_dists<DP, false>(xrow, is_int, children + (size_t)c0 * DP, nc, acc);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int u = 0; u < QS_CH; ++u)
                            if (u < nc) {
                                const float s = acc[r][u];
                                if (s < best[r]) { second[r] = best[r]; best[r] = s; arg[r] = c0 + u; }
                                else if (s < second[r]) second[r] = s;
                            }
                }
            } else {
                // ---- mixed nodes inside the tile (segment boundary): per-descriptor centroid pointers ----
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (!__any_sync(VT_FULL, is_int[r])) continue;
                    const float *children = cent + (size_t)((is_int[r] ? node[r] : n0) * k + 1) * DP;
                    for (int c0 = 0; c0 < k; c0 += QS_CH) {
                        const int nc = k - c0 < QS_CH ? k - c0 : QS_CH;
                        float acc[QS_CH];
                        dist_one<DP>(xrow[r], children + (size_t)c0 * DP, nc, is_int[r], acc);
#pragma unroll
                        for (int u = 0; u < QS_CH; ++u)
                            if (u < nc) {
                                const float s = acc[u];
                                if (s < best[r]) { second[r] = best[r]; best[r] = s; arg[r] = c0 + u; }
                                else if (s < second[r]) second[r] = s;
                            }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                contested[r] = is_int[r] && (k > 1) && (second[r] <= best[r] + best[r] * contest_margin<DP>());
        }
        // binary64 re-evaluation of contested rows, one row at a time by the whole warp
#pragma unroll
        for (int r = 0; r < R; ++r) {
            unsigned int m = __ballot_sync(VT_FULL, contested[r]);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long xq = __shfl_sync(VT_FULL, (unsigned long long)xrow[r], src);
                const long long nq = __shfl_sync(VT_FULL, (long long)node[r], src);
                const int res = canon_argmin_warp<uint8_t, DP>(reinterpret_cast<const uint8_t *>(xq),
                                                               cent + (size_t)(nq * k + 1) * DP, k, lane);
                if (lane == src) { arg[r] = res; ++rechecks; }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t p = base + r * 32 + lane;
            uint32_t key = sentinel;
            if (is_int[r]) {
                const int64_t child = node[r] * k + 1 + arg[r];
                if (is_last || !((child * k + k < n_heap) && internal[child])) {
                    if (leaf_out) leaf_out[row[r]] = (int32_t)child;      // the child is a leaf: stop here
                    if (word_out) word_out[row[r]] = ref_id[child];
                } else {
                    key = jp[r] * (uint32_t)k + (uint32_t)arg[r];          // index of the child inside the next level
                }
            }
            if (valid[r] && keys_out) { keys_out[p] = key; perm_out[p] = row[r]; }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
    if (stats && lane == 0 && rechecks) atomicAdd(stats, (unsigned long long)rechecks);
}

int bits_for(uint64_t max_value)
{
    int b = 0;
    while (max_value) { ++b; max_value >>= 1; }
    return b;
}

template <int DP>
int run_sorted(const uint8_t *desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id, int k,
               int depth, int32_t *leaf, int32_t *word, uint64_t *stats, void *work, cudaStream_t s,
               cudaEvent_t *ev = nullptr)        // ev: optional 2*depth events, one before every stage + one at the end
{
    int n_ev = 0;
#define VT_MARK() do { if (ev) VT_CUDA(cudaEventRecord(ev[n_ev++], s)); } while (0)
    int64_t n_heap = 0, pw = 1;
    for (int l = 0; l <= depth; ++l) { n_heap += pw; pw *= k; }
    uint32_t *kA = (uint32_t *)work, *kB = kA + n, *pA = kB + n, *pB = pA + n;
    void *hist = (void *)(pB + n);
    const int64_t want = (n + 256 * QS_R - 1) / (256 * QS_R), cap = (int64_t)vt_sm_count() * 16;
    const int blocks = (int)(want < cap ? want : cap);
    if (depth == 0) {                                            // the root is the only node
        level_kernel<DP><<<blocks, 256, 0, s>>>(desc, nullptr, nullptr, n, 1u, 0, n_heap, k, 1, cent, internal, ref_id,
                                                nullptr, nullptr, leaf, word, (unsigned long long *)stats);
        VT_LAUNCH_CHECK();
        return VT_OK;
    }
    uint32_t *keys = kA, *perm = pA, *tkeys = kB, *tperm = pB;
    const uint32_t *cur_perm = nullptr, *cur_parent = nullptr;
    int64_t level_off = 0;
    uint64_t n_nodes = 1;
    for (int level = 0; level < depth; ++level) {
        const int is_last = level + 1 == depth;
        // level 0 writes the identity permutation; later levels write new keys beside the sorted perm
        uint32_t *keys_out = is_last ? nullptr : (level == 0 ? keys : tkeys);
        uint32_t *perm_out = is_last ? nullptr : (level == 0 ? perm : tperm);
        VT_MARK();
        level_kernel<DP><<<blocks, 256, 0, s>>>(desc, cur_perm, cur_parent, n, (uint32_t)n_nodes, level_off, n_heap, k,
                                                is_last, cent, internal, ref_id, keys_out, perm_out, leaf, word,
                                                (unsigned long long *)stats);
        VT_LAUNCH_CHECK();
        VT_MARK();
        if (is_last) break;
        if (level > 0) { uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t; }
        int flip = 0;
        const int rc = vt_sort_pairs_impl(keys, perm, tkeys, tperm, n, bits_for(n_nodes * (uint64_t)k), hist, &flip, s);
        if (rc != VT_OK) return rc;
        if (flip) { uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t; }
        cur_parent = keys;
        cur_perm = perm;
        level_off += (int64_t)n_nodes;
        n_nodes *= (uint64_t)k;
    }
#undef VT_MARK
    return VT_OK;
}
}  // namespace

extern "C" int64_t vt_quantize_sorted_work_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return 16 * n + vt_sort_scratch_bytes(n);
}

extern "C" int vt_quantize_sorted(const void *d_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                                  const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth,
                                  int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, void *d_work,
                                  int64_t work_bytes, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8, "the sorted descent takes byte descriptors (use vt_quantize for float rows)");
    VT_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "n out of range");
    VT_CHECK_ARG(k >= 1 && k <= 255 && depth >= 0 && depth <= 16, "k / depth out of range");
    VT_CHECK_ARG(d_centroids && d_internal, "null tree");
    VT_CHECK_ARG(d_word == nullptr || d_ref_id != nullptr, "d_word needs d_ref_id");
    VT_CHECK_ARG(d_work && work_bytes >= vt_quantize_sorted_work_bytes(n), "workspace too small");
    {
        double rows = 1.0;
        for (int l = 0; l < depth; ++l) rows *= k;
        VT_CHECK_ARG(rows < 4.0e9, "k^depth must fit 32 bits");
    }
    return vt_quantize_sorted_impl(d_desc, n, dp, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word,
                                   d_stats, d_work, (cudaStream_t)stream, nullptr);
}

int vt_quantize_sorted_impl(const void *d_desc, int64_t n, int dp, const float *d_centroids, const uint8_t *d_internal,
                            const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap, int32_t *d_word,
                            uint64_t *d_stats, void *d_work, cudaStream_t s, cudaEvent_t *ev)
{
    if (n == 0) return VT_OK;
    switch (dp) {
    case 32: return run_sorted<32>((const uint8_t *)d_desc, n, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word, d_stats, d_work, s, ev);
    case 64: return run_sorted<64>((const uint8_t *)d_desc, n, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word, d_stats, d_work, s, ev);
    case 128: return run_sorted<128>((const uint8_t *)d_desc, n, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word, d_stats, d_work, s, ev);
    }
    vt_set_error("vt_quantize_sorted: dp must be 32, 64 or 128 (got %d)", dp);
    return VT_EUNSUPPORTED;
}

// Measurement variant (bench.py): same work, plus CUDA events around every stage.  h_level_ms[depth]
// receives the duration of each level kernel, h_sort_ms[depth] of each re-sort (last entry 0).
// Synchronises the stream before returning.
extern "C" int vt_quantize_sorted_profile(const void *d_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                                          const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth,
                                          int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, void *d_work,
                                          int64_t work_bytes, float *h_level_ms, float *h_sort_ms, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8 && n > 0 && depth >= 1 && depth <= 16, "bad arguments");
    VT_CHECK_ARG(d_work && work_bytes >= vt_quantize_sorted_work_bytes(n), "workspace too small");
    VT_CHECK_ARG(h_level_ms && h_sort_ms, "null output");
    cudaEvent_t ev[40];
    const int n_ev = 2 * depth;
    for (int i = 0; i < n_ev; ++i) VT_CUDA(cudaEventCreate(&ev[i]));
    int rc = vt_quantize_sorted_impl(d_desc, n, dp, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word,
                                     d_stats, d_work, (cudaStream_t)stream, ev);
    if (rc == VT_OK && cudaEventSynchronize(ev[n_ev - 1]) != cudaSuccess) rc = VT_ECUDA;
    for (int l = 0; l < depth && rc == VT_OK; ++l) {
        cudaEventElapsedTime(&h_level_ms[l], ev[2 * l], ev[2 * l + 1]);
        h_sort_ms[l] = 0.f;
        if (l + 1 < depth) cudaEventElapsedTime(&h_sort_ms[l], ev[2 * l + 1], ev[2 * l + 2]);
    }
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
    return rc;
}
