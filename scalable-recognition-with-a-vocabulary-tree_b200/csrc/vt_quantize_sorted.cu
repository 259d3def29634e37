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
#include <stdlib.h>

int vt_sort_pairs_impl(uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_tmp_keys, uint32_t *d_tmp_vals, int64_t n,
                       int n_bits, void *d_hist, int *result_in_tmp, cudaStream_t s);

int vt_quantize_sorted_impl(const void *d_desc, int64_t n, int dp, const float *d_centroids, const uint8_t *d_internal,
                            const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap, int32_t *d_word,
                            uint64_t *d_stats, void *d_work, cudaStream_t s, cudaEvent_t *ev);

namespace {
constexpr int QS_CH = 10;          // children evaluated per register chunk
constexpr int QS_PW = 4;           // packed words (16 dimensions) consumed per unrolled part

// float32 squared distances of the R descriptors of a lane to the nc (<= QS_CH) child rows at
// `children` (warp-uniform pointer): every centroid float4 is loaded once and used for R rows.
template <int DP, int R, bool FULL>
__device__ __forceinline__ void tile_dists(const uint8_t *const (&xrow)[R], const bool (&on)[R],
                                           const float *__restrict__ children, int nc, float (&acc)[R][QS_CH])
{
    constexpr int NW = DP / 4;
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

// One tree level for the whole (sorted) batch.  A warp owns a tile of 32*R consecutive positions;
// lane l owns positions l, l+32, ...  Tiles whose rows all sit in one node (the common case after
// sorting) take one pass; tiles that straddle segment boundaries take one pass per distinct node,
// each pass at full speed with the other rows masked off.
template <int DP, int R>
__global__ void __launch_bounds__(256, R >= 4 ? 2 : (R == 2 ? 3 : 4))
level_kernel(const uint8_t *__restrict__ desc, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ parent,
             int64_t n, int level, int bits, int64_t n_heap, int k, int is_last,
             const float *__restrict__ cent, const uint8_t *__restrict__ internal, const int32_t *__restrict__ ref_id,
             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ perm_out, int32_t *__restrict__ leaf_out,
             int32_t *__restrict__ word_out, unsigned long long *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t dmask = (1u << bits) - 1u;                    // all-ones digit = "stopped at a leaf"
    const float inf = __int_as_float(0x7f800000);
    const long long none = 0x7fffffffffffffffLL;
    unsigned int rechecks = 0;

    for (int64_t base = warp * (32 * R); base < n; base += nwarps * (32 * R)) {
        uint32_t row[R], jp[R];
        long long node[R];
        bool valid[R], is_int[R];
        const uint8_t *xrow[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t p = base + r * 32 + lane;
            valid[r] = p < n;
            row[r] = valid[r] ? (perm ? perm[p] : (uint32_t)p) : 0u;
            jp[r] = valid[r] && parent ? parent[p] : 0u;         // packed path from the root
            const bool alive = valid[r] && (level == 0 || (jp[r] & dmask) != dmask);   // not yet stopped at a leaf
            node[r] = alive ? path_to_heap(jp[r], level, bits, k) : 0;
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
        }
        int arg[R];
        float best[R], second[R];
        bool pending[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { arg[r] = 0; best[r] = inf; second[r] = inf; pending[r] = is_int[r]; }

        // one pass per distinct node present in the tile (ascending node order)
        while (true) {
            long long mine = none;
#pragma unroll
            for (int r = 0; r < R; ++r) if (pending[r] && node[r] < mine) mine = node[r];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const long long other = __shfl_xor_sync(VT_FULL, mine, o);
                mine = other < mine ? other : mine;
            }
            if (mine == none) break;                             // warp-uniform
            bool on[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { on[r] = pending[r] && node[r] == mine; pending[r] = pending[r] && !on[r]; }
            const float *children = cent + (size_t)(mine * k + 1) * DP;
            for (int c0 = 0; c0 < k; c0 += QS_CH) {
                const int nc = k - c0 < QS_CH ? k - c0 : QS_CH;
                float acc[R][QS_CH];
                if (nc == QS_CH) tile_dists<DP, R, true>(xrow, on, children + (size_t)c0 * DP, nc, acc);
                else tile_dists<DP, R, false>(xrow, on, children + (size_t)c0 * DP, nc, acc);
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int u = 0; u < QS_CH; ++u)
                        if (u < nc && on[r]) {
                            const float s = acc[r][u];
                            if (s < best[r]) { second[r] = best[r]; best[r] = s; arg[r] = c0 + u; }
                            else if (s < second[r]) second[r] = s;
                        }
            }
        }
        // binary64 re-evaluation of contested rows, one row at a time by the whole warp
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool contested = is_int[r] && (k > 1) && (second[r] <= best[r] + best[r] * contest_margin<DP>());
            unsigned int m = __ballot_sync(VT_FULL, contested);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long xq = __shfl_sync(VT_FULL, (unsigned long long)xrow[r], src);
                const long long nq = __shfl_sync(VT_FULL, node[r], src);
                const int res = canon_argmin_warp<uint8_t, DP>(reinterpret_cast<const uint8_t *>(xq),
                                                               cent + (size_t)(nq * k + 1) * DP, k, lane);
                if (lane == src) { arg[r] = res; ++rechecks; }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t p = base + r * 32 + lane;
            uint32_t key = (jp[r] << bits) | dmask;                         // stopped
            if (is_int[r]) {
                const long long child = node[r] * k + 1 + arg[r];
                if (is_last || !((child * k + k < n_heap) && internal[child])) {
                    if (leaf_out) leaf_out[row[r]] = (int32_t)child;      // the child is a leaf: stop here
                    if (word_out) word_out[row[r]] = ref_id[child];
                } else {
                    key = (jp[r] << bits) | (uint32_t)arg[r];              // path of the child
                }
            }
            if (valid[r] && keys_out) { keys_out[p] = key; perm_out[p] = row[r]; }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
    if (stats && lane == 0 && rechecks) atomicAdd(stats, (unsigned long long)rechecks);
}

template <int DP>
void launch_level(int r_tile, int blocks_cap_per_sm, const uint8_t *desc, const uint32_t *perm, const uint32_t *parent,
                  int64_t n, int level, int bits, int64_t n_heap, int k, int is_last, const float *cent,
                  const uint8_t *internal, const int32_t *ref_id, uint32_t *keys_out, uint32_t *perm_out, int32_t *leaf,
                  int32_t *word, unsigned long long *stats, cudaStream_t s)
{
    const int64_t cap = (int64_t)vt_sm_count() * blocks_cap_per_sm;
#define VT_LAUNCH_R(RV)                                                                                         \
    {                                                                                                           \
        const int64_t want = (n + 256 * RV - 1) / (256 * RV);                                                   \
        level_kernel<DP, RV><<<(int)(want < cap ? want : cap), 256, 0, s>>>(desc, perm, parent, n, level,       \
            bits, n_heap, k, is_last, cent, internal, ref_id, keys_out, perm_out, leaf, word, stats);           \
    }
    if (r_tile >= 4) VT_LAUNCH_R(4)
    else if (r_tile == 2) VT_LAUNCH_R(2)
    else VT_LAUNCH_R(1)
#undef VT_LAUNCH_R
}

// rows per lane for a level: large tiles only pay off while segments are much longer than a tile
int pick_r(int64_t n, uint64_t n_nodes)
{
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("VT_QS_R");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    const double seg = (double)n / (double)n_nodes;
    return seg >= 192.0 ? 2 : 1;        // measured on B200: R=2 (24 warps/SM) beats R=4 (16 warps/SM); VT_QS_R overrides
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
    if (depth == 0) {                                            // the root is the only node
        launch_level<DP>(1, 16, desc, nullptr, nullptr, n, 0, path_bits(k), n_heap, k, 1, cent, internal, ref_id, nullptr,
                         nullptr, leaf, word, (unsigned long long *)stats, s);
        VT_LAUNCH_CHECK();
        return VT_OK;
    }
    uint32_t *keys = kA, *perm = pA, *tkeys = kB, *tperm = pB;
    const uint32_t *cur_perm = nullptr, *cur_parent = nullptr;
    const int bits = path_bits(k);
    uint64_t n_nodes = 1;
    for (int level = 0; level < depth; ++level) {
        const int is_last = level + 1 == depth;
        // level 0 writes the identity permutation; later levels write new keys beside the sorted perm
        uint32_t *keys_out = is_last ? nullptr : (level == 0 ? keys : tkeys);
        uint32_t *perm_out = is_last ? nullptr : (level == 0 ? perm : tperm);
        VT_MARK();
        launch_level<DP>(pick_r(n, n_nodes), 16, desc, cur_perm, cur_parent, n, level, bits, n_heap, k,
                         is_last, cent, internal, ref_id, keys_out, perm_out, leaf, word, (unsigned long long *)stats, s);
        VT_LAUNCH_CHECK();
        VT_MARK();
        if (is_last) break;
        if (level > 0) { uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t; }
        int flip = 0;
        // ONE stable pass over the low 8 bits of the packed paths groups the rows by their new node
        const int rc = vt_sort_pairs_impl(keys, perm, tkeys, tperm, n, 8, hist, &flip, s);
        if (rc != VT_OK) return rc;
        if (flip) { uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t; }
        cur_parent = keys;
        cur_perm = perm;
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
    VT_CHECK_ARG(path_bits(k) <= 8 && (depth - 1) * path_bits(k) <= 32, "k / depth too large for 32-bit packed paths");
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
