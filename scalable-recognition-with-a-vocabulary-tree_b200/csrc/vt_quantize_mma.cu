// vt_quantize_mma.cu -- tcgen05 (5th-gen tensor core) variant of the level-synchronous descent
// (VocabularyTree.propagate_feature, cbir/encoders/vocabulary.py:104-124) for byte descriptors, and of the
// k-means assignment of the tree build (VocabularyTree.fit, vocabulary.py:62-66).
//
// After the batch is grouped by node, the 128 descriptors of a tile meet the SAME <= 16 child
// centroids: the distance step really is a dense contraction  X[128 x 128] . C^T[128 x k].
// It is done in integers on the INT8 tensor pipe:
//   * descriptors are bytes -> the A operand is the gathered rows as they lie in HBM (cp.async
//     into a 128B-swizzled K-major tile, no conversion);
//   * centroids are rounded to 8.16 fixed point (|delta| <= 2^-17) and split into three byte
//     planes -> B operand = 3 x KP rows of bytes per node (KP = 10 child slots when k <= 10, else 16),
//     prepared once per tree;
//   * tcgen05.mma kind::i8, M=128, N=32 or 48, K=32 x 4, int32 accumulators in TMEM (exact: 128 * 255 * 255
//     < 2^31); tcgen05.ld brings one accumulator row to the thread that owns that descriptor;
//   * s_c = |c|^2 - 2 x.c is compared in 32 bits, units 2^-7 (score_children): each score is within one
//     unit of the exact score of the ROUNDED centroid.
// Two bounded approximations, then: the centroid rounding, |s_c(rounded) - s_c(float32 centroid)| <=
// 2^-16 * sqrt(D) * ||c - x|| + D * 2^-34, and the 2^-7 score units.  Rows whose runner-up is inside the
// margin that covers both (contest_test) are re-evaluated in canonical binary64 exactly like in the
// SIMT kernels.  Results are therefore bit-identical to vt_quantize / vt_quantize_sorted and the oracle.
//
// The last level of large batches does not write its ids straight to the (row-indexed) output -- 4-byte
// writes all over a 400 MB array -- but emits (row, child) pairs in tile order, buckets them by row range
// with one radix pass and writes the ids bucket by bucket (deferred_output_kernel).
//
// Preconditions: dp == 128, k <= 16, every child centroid in [0, 255.99] (true for trees built from
// byte descriptors; the prepare kernel reports violations in d_stats[1]).
#include "vt_mma_common.cuh"
#include <stdlib.h>

int vt_sort_pairs_impl(uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_tmp_keys, uint32_t *d_tmp_vals, int64_t n,
                       int n_bits, void *d_hist, int *result_in_tmp, cudaStream_t s);
int vt_sort_pass_impl(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, int64_t n, int shift,
                      uint32_t mask, void *d_hist, cudaStream_t s);

namespace {
// ---- per-call preparation: byte planes + fixed-point squared norms of every node's children --------
template <int KP>
__global__ void __launch_bounds__(128)
prep_planes_kernel(const float *__restrict__ cent, int64_t n_int, int k, uint8_t *__restrict__ bimg,
                   unsigned long long *__restrict__ cn, int *__restrict__ cn32, unsigned long long *__restrict__ stats)
{
    __shared__ unsigned long long s_part[4];
    const int d = threadIdx.x;
    for (int64_t h = blockIdx.x; h < n_int; h += gridDim.x) {
        uint8_t *img = bimg + (size_t)h * PlaneCfg<KP>::B_BYTES;
        for (int r = 3 * KP; r < PlaneCfg<KP>::N; ++r) img[sw128_offset(r, d)] = 0;      // unused accumulator columns
        bool bad = false;
        for (int c = 0; c < 16; ++c) {
            uint32_t fx = 0;
            if (c < k) {
                const float v = cent[(size_t)(h * k + 1 + c) * 128 + d];
                if (!(v >= 0.f) || v > 255.99f) bad = true;
                fx = (uint32_t)rintf(fminf(fmaxf(v, 0.f), 255.99f) * 65536.f);
            }
            if (c < KP) {                                        // (k <= KP: the other slots do not exist)
                img[sw128_offset(c, d)] = (uint8_t)(fx >> 16);
                img[sw128_offset(KP + c, d)] = (uint8_t)(fx >> 8);
                img[sw128_offset(2 * KP + c, d)] = (uint8_t)fx;
            }
            unsigned long long sq = (unsigned long long)fx * fx;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(VT_FULL, sq, o);
            if ((d & 31) == 0) s_part[d >> 5] = sq;
            __syncthreads();
            if (d == 0) {
                const unsigned long long tot = s_part[0] + s_part[1] + s_part[2] + s_part[3];
                if (c < k) cn[h * k + 1 + c] = tot;
                cn32[h * 16 + c] = c < k ? (int)(tot >> MM_S_SHIFT) : 0x3fffffff;     // |c|^2 in units 2^-7, rounded down
            }
            __syncthreads();
        }
        if (bad && stats) atomicAdd(stats + 1, 1ull);
    }
}

// ---- one tree level -------------------------------------------------------------------------------
template <int MINB, int KP>
__global__ void __launch_bounds__(MM_TILE, MINB)
level_mma_kernel(const uint8_t *__restrict__ desc, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ parent,
                 int64_t n, int level, int bits, int64_t n_heap, int k, int is_last,
                 const float *__restrict__ cent, const uint8_t *__restrict__ internal, const int32_t *__restrict__ ref_id,
                 const uint8_t *__restrict__ bimg, const int *__restrict__ cn32,
                 uint32_t *__restrict__ keys_out, uint32_t *__restrict__ perm_out, int32_t *__restrict__ leaf_out,
                 int32_t *__restrict__ word_out, unsigned long long *__restrict__ stats)
{
    using Cfg = PlaneCfg<KP>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the swizzled tiles need a 1024 B aligned base (the launch reserves 1 KB of slack for this)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    uint8_t *sB = smem + MM_A_BYTES;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_row[MM_TILE];
    __shared__ long long s_min[4];
    __shared__ int s_max[4];
    __shared__ __align__(16) int s_cn[16];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t dmask = (1u << bits) - 1u;                    // all-ones digit = "stopped at a leaf"
    const int none = 0x7fffffff;
    unsigned int rechecks = 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const uint64_t a_desc = make_desc(smem_u32(sA));
    const uint64_t b_desc = make_desc(smem_u32(sB));
    uint32_t phase = 0;

    // the permutation / path of the NEXT tile is loaded while the current one is processed: one DRAM round
    // trip less on the per-tile dependency chain (prefetching the rows themselves into L2 measured slower)
    uint32_t nx_row = 0, nx_jp = 0;
    {
        const int64_t p0 = (int64_t)blockIdx.x * MM_TILE + tid;
        if (p0 < n) { nx_row = perm ? perm[p0] : (uint32_t)p0; nx_jp = parent ? parent[p0] : 0u; }
    }
    for (int64_t base = (int64_t)blockIdx.x * MM_TILE; base < n; base += (int64_t)gridDim.x * MM_TILE) {
        const int64_t p = base + tid;
        const bool valid = p < n;
        const uint32_t row = valid ? nx_row : 0u;
        const uint32_t jp = valid ? nx_jp : 0u;                  // packed path from the root
        {
            const int64_t pn = p + (int64_t)gridDim.x * MM_TILE;
            nx_row = 0; nx_jp = 0;
            if (pn < n) { nx_row = perm ? perm[pn] : (uint32_t)pn; nx_jp = parent ? parent[pn] : 0u; }
        }
        const bool alive = valid && (level == 0 || (jp & dmask) != dmask);
        const long long node = alive ? (long long)path_to_heap32(jp, level, bits, (uint32_t)k) : 0;
        const bool is_int = alive && (node * k + k < n_heap) && internal[node];
        const uint8_t *xrow = desc + (size_t)row * 128;
        if (alive && !is_int) {                                  // ragged tree: this node is a leaf
            if (leaf_out) leaf_out[row] = (int32_t)node;
            if (word_out) word_out[row] = ref_id[node];
        }
        s_row[tid] = is_int ? row : 0xffffffffu;
        __syncthreads();
        // A tile: 8 lanes copy one 128 B row (coalesced), 16 rows per pass, swizzled destination
        {
            // row r = (tid >> 3) + 16 i: its swizzle phase (r & 7) does not depend on i, so the destination
            // advances by a constant 2 KB per pass
            const int r0 = tid >> 3, j = tid & 7;
            const uint32_t dst0 = smem_u32(sA) + (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
            const uint8_t *col = desc + j * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t rr = s_row[r0 + 16 * i];
                const bool ok = rr != 0xffffffffu;
                const uint8_t *src = col + (size_t)(ok ? rr : 0u) * 128;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 2048u * i), "l"(src),
                             "r"(ok ? 16 : 0));                 // zero fill for rows that take no part
            }
        }
        int best = none, second = none, arg = 0;
        bool pending = is_int;

        // one MMA pass per distinct node present in the tile (sorted batch: almost always one)
        for (int pass = 0;; ++pass) {
            // heap / level indices fit 31 bits (ids are int32 across the ABI): 32-bit warp minimum
            // (the maximum rides along: when it equals the minimum this is the tile's only node and the loop
            // ends after this pass without another block-wide reduction)
            int mine32 = pending ? (int)node : 0x7fffffff;
            int max32 = pending ? (int)node : -1;
            mine32 = __reduce_min_sync(VT_FULL, mine32);
            max32 = __reduce_max_sync(VT_FULL, max32);
            if (lane == 0) { s_min[warp] = mine32; s_max[warp] = max32; }
            __syncthreads();
            long long mine = s_min[0];
            int maxn = s_max[0];
#pragma unroll
            for (int w = 1; w < 4; ++w) {
                mine = s_min[w] < mine ? s_min[w] : mine;
                maxn = s_max[w] > maxn ? s_max[w] : maxn;
            }
            if (mine == 0x7fffffff) break;                            // block-uniform
            const bool only_node = (long long)maxn == mine;           // block-uniform
            const bool on = pending && node == mine;
            pending = pending && !on;
            // B tile of this node: 384 x 16 B, already in swizzled image form
            const uint8_t *bsrc = bimg + (size_t)mine * Cfg::B_BYTES;
#pragma unroll
            for (int i = 0; i < Cfg::B_BYTES / 16 / MM_TILE; ++i) {
                const int c = tid + MM_TILE * i;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB) + c * 16), "l"(bsrc + c * 16));
            }
            if (tid < 4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_cn) + tid * 16), "l"(cn32 + mine * 16 + tid * 4));
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                // K = 4 x 32 bytes
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                        ::"r"(tmem), "l"(a_desc + 2 * ks), "l"(b_desc + 2 * ks), "r"(Cfg::IDESC), "r"(acc), "r"(0u));
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
            {   // wait for the accumulator
                uint32_t done;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
                } while (!done);
                phase ^= 1u;
            }
            asm volatile("tcgen05.fence::after_thread_sync;");
            uint32_t v[Cfg::N];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int g = 0; g < Cfg::N / 16; ++g) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[16 * g + 0]), "=r"(v[16 * g + 1]), "=r"(v[16 * g + 2]), "=r"(v[16 * g + 3]),
                               "=r"(v[16 * g + 4]), "=r"(v[16 * g + 5]), "=r"(v[16 * g + 6]), "=r"(v[16 * g + 7]),
                               "=r"(v[16 * g + 8]), "=r"(v[16 * g + 9]), "=r"(v[16 * g + 10]), "=r"(v[16 * g + 11]),
                               "=r"(v[16 * g + 12]), "=r"(v[16 * g + 13]), "=r"(v[16 * g + 14]), "=r"(v[16 * g + 15])
                             : "r"(taddr + 16 * g));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (on) score_children<KP, Cfg::N>(v, s_cn, k, best, second, arg);
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();                                    // TMEM and the B tile may be reused
            if (only_node) break;
        }

        // provable margin for the centroid rounding; contested rows go to canonical binary64
        const bool contested = contested_row(is_int && k > 1, best, second, sA, tid);
        unsigned int m = __ballot_sync(VT_FULL, contested);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned long long xq = __shfl_sync(VT_FULL, (unsigned long long)xrow, src);
            const long long nq = __shfl_sync(VT_FULL, node, src);
            const int res = canon_argmin_warp<uint8_t, 128>(reinterpret_cast<const uint8_t *>(xq),
                                                            cent + (size_t)(nq * k + 1) * 128, k, lane);
            if (lane == src) { arg = res; ++rechecks; }
        }
        uint32_t key = (jp << bits) | dmask;                     // stopped
        const long long child = node * k + 1 + arg;
        if (is_last == 2) {
            // deferred output: (row, child) pairs leave in tile order (coalesced); vt_mma_levels buckets them
            // by row range and writes the ids from there, instead of 4-byte writes all over the output here
            if (valid) { keys_out[p] = is_int ? row : 0xffffffffu; perm_out[p] = (uint32_t)child; }
        } else {
            if (is_int) {
                if (is_last || !((child * k + k < n_heap) && internal[child])) {
                    if (leaf_out) leaf_out[row] = (int32_t)child;
                    if (word_out) word_out[row] = ref_id[child];
                } else {
                    key = (jp << bits) | (uint32_t)arg;          // path of the child
                }
            }
            if (valid && keys_out) { keys_out[p] = key; perm_out[p] = row; }
        }
        __syncthreads();                                        // s_row / A tile reused by the next tile
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
    if (stats && lane == 0 && rechecks) atomicAdd(stats, (unsigned long long)rechecks);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS));
}

// deferred output of the last level: pairs bucketed by row range, so the 4-byte writes of neighbouring CTAs
// fall into a window of a few MB that the L2 turns into full lines
__global__ void __launch_bounds__(256)
deferred_output_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ child, int64_t n,
                       const int32_t *__restrict__ ref_id, int32_t *__restrict__ leaf_out, int32_t *__restrict__ word_out)
{
    // the partially written sectors of the current bucket must stay in L2 until their other words arrive
    // (evict_last); the pairs are read once (evict_first)
    unsigned long long keep, stream;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
    // (measured: four independent pairs per thread and step -- 4 x 256 consecutive pairs per CTA -- is slower, 2.7 against
    // 2.1 ms for the whole deferred output: the wider write window costs more than the overlapped chains gain)
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r, c;
        asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(rows + p), "l"(stream));
        if (r == 0xffffffffu) continue;
        asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(c) : "l"(child + p), "l"(stream));
        if (leaf_out) asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(leaf_out + r), "r"(c), "l"(keep) : "memory");
        if (word_out)
            asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(word_out + r), "r"(ref_id[c]), "l"(keep) : "memory");
    }
}
}  // namespace

// =====================================================================================================
// k-means assignment + exact accumulation of one tree level on the tensor pipe (VocabularyTree.fit,
// cbir/encoders/vocabulary.py:62-66 with the seeded exact Lloyd).  Same distance step as
// level_mma_kernel; the per-cluster sums are a SECOND integer MMA on the same shared-memory tile:
//   S^T[d][c] = sum_i X[i][d] * onehot[i][c]      (A = X read MN-major, B = one-hot labels, K = 128 rows)
// so the 128 x 128 bytes of a tile are added up by the tensor core instead of 16 K atomics.  Every
// thread then owns one dimension d of the 16 running cluster sums of the node it is working on and
// flushes them with uint64 atomics only when the node changes.
// Level-local indexing (as vt_kmeans_assign): parent[p] = node index inside the level, child rows
// j*k + c, bimg/cn prepared per level by vt_kmeans_planes.
// =====================================================================================================
namespace {
constexpr uint32_t MM_IDESC_SUM = (2u << 4) | (1u << 15) /* A is MN-major */ | ((uint32_t)(16 >> 3) << 17) |
                                  ((uint32_t)(MM_TILE >> 4) << 24);
constexpr int MM_H_BYTES = 16 * 128;           // one-hot tile: 16 clusters x 128 rows

template <int KP>
__global__ void __launch_bounds__(128)
prep_level_planes_kernel(const float *__restrict__ child_cent, int64_t n_nodes, int k, uint8_t *__restrict__ bimg,
                         int *__restrict__ cn32, unsigned long long *__restrict__ stats)
{
    __shared__ unsigned long long s_part[4];
    const int d = threadIdx.x;
    for (int64_t h = blockIdx.x; h < n_nodes; h += gridDim.x) {
        uint8_t *img = bimg + (size_t)h * PlaneCfg<KP>::B_BYTES;
        for (int r = 3 * KP; r < PlaneCfg<KP>::N; ++r) img[sw128_offset(r, d)] = 0;      // unused accumulator columns
        bool bad = false;
        for (int c = 0; c < 16; ++c) {
            uint32_t fx = 0;
            if (c < k) {
                const float v = child_cent[(size_t)(h * k + c) * 128 + d];
                if (!(v >= 0.f) || v > 255.99f) bad = true;
                fx = (uint32_t)rintf(fminf(fmaxf(v, 0.f), 255.99f) * 65536.f);
            }
            if (c < KP) {                                        // (k <= KP: the other slots do not exist)
                img[sw128_offset(c, d)] = (uint8_t)(fx >> 16);
                img[sw128_offset(KP + c, d)] = (uint8_t)(fx >> 8);
                img[sw128_offset(2 * KP + c, d)] = (uint8_t)fx;
            }
            unsigned long long sq = (unsigned long long)fx * fx;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(VT_FULL, sq, o);
            if ((d & 31) == 0) s_part[d >> 5] = sq;
            __syncthreads();
            if (d == 0)     // |c|^2 in units 2^-7, rounded down (see score_children)
                cn32[h * 16 + c] = c < k ? (int)((s_part[0] + s_part[1] + s_part[2] + s_part[3]) >> MM_S_SHIFT) : 0x3fffffff;
            __syncthreads();
        }
        if (bad && stats) atomicAdd(stats + 1, 1ull);
    }
}

template <int KP>
__global__ void __launch_bounds__(MM_TILE, 6)
kmeans_mma_kernel(const uint8_t *__restrict__ desc, const int32_t *__restrict__ perm, const int32_t *__restrict__ parent,
                  int64_t n, const uint8_t *__restrict__ node_internal, int k, const float *__restrict__ child_cent,
                  const uint8_t *__restrict__ bimg, const int *__restrict__ cn32,
                  uint8_t *__restrict__ label, unsigned long long *__restrict__ sums,
                  unsigned long long *__restrict__ counts, unsigned long long *__restrict__ changed,
                  unsigned long long *__restrict__ stats)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    using Cfg = PlaneCfg<KP>;
    uint8_t *sB = smem + MM_A_BYTES;                 // 4 or 6 KB centroid planes
    uint8_t *sH = smem + MM_A_BYTES + Cfg::B_BYTES;  // 2 KB one-hot labels (1024 B aligned)
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_row[MM_TILE];
    __shared__ long long s_min[4];
    __shared__ int s_cnt[16];
    __shared__ int s_max[4];
    __shared__ __align__(16) int s_cn[16];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int none = 0x7fffffff;
    unsigned int rechecks = 0, n_changed = 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(MM_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < 16) s_cnt[tid] = 0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const uint64_t a_desc = make_desc(smem_u32(sA));
    const uint64_t b_desc = make_desc(smem_u32(sB));
    const uint64_t h_desc = make_desc(smem_u32(sH));
    uint32_t phase = 0;

    // running sums of the node this CTA is working on: thread = dimension, 16 clusters
    int run[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) run[c] = 0;
    long long run_node = -1;
    int run_tiles = 0;

    auto flush = [&]() {
        if (run_node >= 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < k && run[c]) atomicAdd(&sums[(size_t)(run_node * k + c) * 128 + tid], (unsigned long long)run[c]);
            if (tid < k && s_cnt[tid]) atomicAdd(&counts[run_node * k + tid], (unsigned long long)s_cnt[tid]);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) run[c] = 0;
        __syncthreads();
        if (tid < 16) s_cnt[tid] = 0;
        __syncthreads();
        run_tiles = 0;
    };

    auto wait_mma = [&]() {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
        } while (!done);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;");
    };

    for (int64_t base = (int64_t)blockIdx.x * MM_TILE; base < n; base += (int64_t)gridDim.x * MM_TILE) {
        const int64_t p = base + tid;
        const bool valid = p < n;
        const uint32_t row = valid ? (uint32_t)perm[p] : 0u;
        const long long node = valid ? (long long)parent[p] : 0;
        const bool is_int = valid && node_internal[node];
        const uint8_t *xrow = desc + (size_t)row * 128;
        s_row[tid] = is_int ? row : 0xffffffffu;
        __syncthreads();
        {
            const int r0 = tid >> 3, j = tid & 7;
            const uint32_t dst0 = smem_u32(sA) + (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
            const uint8_t *col = desc + j * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t rr = s_row[r0 + 16 * i];
                const bool ok = rr != 0xffffffffu;
                const uint8_t *src = col + (size_t)(ok ? rr : 0u) * 128;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 2048u * i), "l"(src), "r"(ok ? 16 : 0));
            }
        }
        int best = none, second = none, arg = 0;
        bool pending = is_int;

        // ---- assignment: one distance MMA per distinct node of the tile --------------------------------
        // (the maximum rides along with the minimum: a tile with a single node -- the usual case -- leaves
        // both loops after one pass without another block-wide reduction)
        long long tile_node = -1;                               // the tile's only node, if it has just one
        for (int pass = 0;; ++pass) {
            // heap / level indices fit 31 bits (ids are int32 across the ABI): 32-bit warp minimum
            int mine32 = pending ? (int)node : 0x7fffffff;
            int max32 = pending ? (int)node : -1;
            mine32 = __reduce_min_sync(VT_FULL, mine32);
            max32 = __reduce_max_sync(VT_FULL, max32);
            if (lane == 0) { s_min[warp] = mine32; s_max[warp] = max32; }
            __syncthreads();
            long long mine = s_min[0];
            int maxn = s_max[0];
#pragma unroll
            for (int w = 1; w < 4; ++w) {
                mine = s_min[w] < mine ? s_min[w] : mine;
                maxn = s_max[w] > maxn ? s_max[w] : maxn;
            }
            if (mine == 0x7fffffff) break;
            const bool only_node = (long long)maxn == mine;           // block-uniform
            if (pass == 0 && only_node) tile_node = mine;
            const bool on = pending && node == mine;
            pending = pending && !on;
            const uint8_t *bsrc = bimg + (size_t)mine * Cfg::B_BYTES;
#pragma unroll
            for (int i = 0; i < Cfg::B_BYTES / 16 / MM_TILE; ++i) {
                const int c = tid + MM_TILE * i;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sB) + c * 16), "l"(bsrc + c * 16));
            }
            if (tid < 4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_cn) + tid * 16), "l"(cn32 + mine * 16 + tid * 4));
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                        ::"r"(tmem), "l"(a_desc + 2 * ks), "l"(b_desc + 2 * ks), "r"(Cfg::IDESC), "r"(acc), "r"(0u));
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
            wait_mma();
            uint32_t v[Cfg::N];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int g = 0; g < Cfg::N / 16; ++g) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[16 * g + 0]), "=r"(v[16 * g + 1]), "=r"(v[16 * g + 2]), "=r"(v[16 * g + 3]),
                               "=r"(v[16 * g + 4]), "=r"(v[16 * g + 5]), "=r"(v[16 * g + 6]), "=r"(v[16 * g + 7]),
                               "=r"(v[16 * g + 8]), "=r"(v[16 * g + 9]), "=r"(v[16 * g + 10]), "=r"(v[16 * g + 11]),
                               "=r"(v[16 * g + 12]), "=r"(v[16 * g + 13]), "=r"(v[16 * g + 14]), "=r"(v[16 * g + 15])
                             : "r"(taddr + 16 * g));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (on) score_children<KP, Cfg::N>(v, s_cn, k, best, second, arg);
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();
            if (only_node) break;
        }
        const bool contested = contested_row(is_int && k > 1, best, second, sA, tid);
        unsigned int m = __ballot_sync(VT_FULL, contested);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned long long xq = __shfl_sync(VT_FULL, (unsigned long long)xrow, src);
            const long long nq = __shfl_sync(VT_FULL, node, src);
            const int res = canon_argmin_warp<uint8_t, 128>(reinterpret_cast<const uint8_t *>(xq),
                                                            child_cent + (size_t)(nq * k) * 128, k, lane);
            if (lane == src) { arg = res; ++rechecks; }
        }
        if (is_int) {
            if (label[p] != (uint8_t)arg) { ++n_changed; label[p] = (uint8_t)arg; }
        }

        // ---- accumulation: one one-hot MMA per distinct node of the tile --------------------------------
        pending = is_int;
        for (;;) {
            long long mine = tile_node;
            if (tile_node < 0) {
                // heap / level indices fit 31 bits (ids are int32 across the ABI): 32-bit warp minimum
                int mine32 = pending ? (int)node : 0x7fffffff;
                mine32 = __reduce_min_sync(VT_FULL, mine32);
                if (lane == 0) s_min[warp] = mine32;
                __syncthreads();
                mine = s_min[0];
#pragma unroll
                for (int w = 1; w < 4; ++w) mine = s_min[w] < mine ? s_min[w] : mine;
                if (mine == 0x7fffffff) break;
            }
            const bool on = pending && node == mine;
            pending = pending && !on;
            if (mine != run_node || run_tiles >= 60000) { flush(); run_node = mine; }   // block-uniform
            ++run_tiles;
            // one-hot tile: row c (cluster), byte i (this thread's row): 1 iff label == c
#pragma unroll
            for (int c = 0; c < 16; ++c) sH[sw128_offset(c, tid)] = (on && arg == c) ? 1 : 0;
            if (on) atomicAdd(&s_cnt[arg], 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                // K = 4 x 32 rows; A advances 32 rows = 4096 B
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                        ::"r"(tmem + (uint32_t)Cfg::N), "l"(a_desc + 256 * ks), "l"(h_desc + 2 * ks), "r"(MM_IDESC_SUM), "r"(acc), "r"(0u));
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
            wait_mma();
            uint32_t sv[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)Cfg::N;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(sv[0]), "=r"(sv[1]), "=r"(sv[2]), "=r"(sv[3]), "=r"(sv[4]), "=r"(sv[5]), "=r"(sv[6]), "=r"(sv[7]),
                           "=r"(sv[8]), "=r"(sv[9]), "=r"(sv[10]), "=r"(sv[11]), "=r"(sv[12]), "=r"(sv[13]), "=r"(sv[14]), "=r"(sv[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) run[c] += (int)sv[c];
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();
            if (tile_node >= 0) break;
        }
        __syncthreads();
    }
    flush();
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
        n_changed += __shfl_xor_sync(VT_FULL, n_changed, o);
    }
    if (lane == 0) {
        if (n_changed && changed) atomicAdd(changed, (unsigned long long)n_changed);
        if (rechecks && stats) atomicAdd(stats, (unsigned long long)rechecks);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(MM_TMEM_COLS));
}
}  // namespace

// centroid byte planes + fixed-point norms of one LEVEL (node j -> image j, child rows j*k + c)
extern "C" int64_t vt_kmeans_planes_bytes(int64_t n_nodes, int k)
{
    const int64_t plane = k <= 10 ? PlaneCfg<10>::B_BYTES : PlaneCfg<16>::B_BYTES;
    return ((n_nodes * plane + 255) / 256) * 256 + n_nodes * 64;           // byte planes + 16 32-bit norms per node
}

extern "C" int vt_kmeans_planes(const float *d_child_centroids, int64_t n_nodes, int k, int dp, void *d_planes,
                                uint64_t *d_stats, void *stream)
{
    VT_CHECK_ARG(dp == 128 && k >= 1 && k <= 16 && n_nodes >= 1, "tensor-core k-means needs dp == 128 and k <= 16");
    uint8_t *bimg = (uint8_t *)d_planes;
    const int64_t plane = k <= 10 ? PlaneCfg<10>::B_BYTES : PlaneCfg<16>::B_BYTES;
    int *cn = (int *)(bimg + ((n_nodes * plane + 255) / 256) * 256);
    const int64_t cap = (int64_t)vt_sm_count() * 16;
    const int blocks = (int)(n_nodes < cap ? n_nodes : cap);
    if (k <= 10)
        prep_level_planes_kernel<10><<<blocks, 128, 0, (cudaStream_t)stream>>>(d_child_centroids, n_nodes, k, bimg, cn,
                                                                               (unsigned long long *)d_stats);
    else
        prep_level_planes_kernel<16><<<blocks, 128, 0, (cudaStream_t)stream>>>(d_child_centroids, n_nodes, k, bimg, cn,
                                                                               (unsigned long long *)d_stats);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// Same contract as vt_kmeans_assign (labels, exact uint64 sums / counts, changed counter), distance
// step and cluster sums on the INT8 tensor pipe.  d_planes from vt_kmeans_planes for the CURRENT centres.
extern "C" int vt_kmeans_assign_mma(const void *d_desc, int dtype, int dp, const int32_t *d_perm, const int32_t *d_parent,
                                    int64_t n_members, const uint8_t *d_node_internal, int64_t n_nodes, int k,
                                    const float *d_child_centroids, const void *d_planes, uint8_t *d_label,
                                    uint64_t *d_sums, uint64_t *d_counts, uint64_t *d_changed, uint64_t *d_stats,
                                    void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8 && dp == 128 && k >= 1 && k <= 16, "tensor-core k-means needs byte rows, dp == 128, k <= 16");
    VT_CHECK_ARG(n_members >= 0 && n_nodes >= 1 && d_planes, "bad arguments");
    if (n_members == 0) return VT_OK;
    const uint8_t *bimg = (const uint8_t *)d_planes;
    const int64_t plane = k <= 10 ? PlaneCfg<10>::B_BYTES : PlaneCfg<16>::B_BYTES;
    const int *cn = (const int *)(bimg + ((n_nodes * plane + 255) / 256) * 256);
    static bool attr_set = false;
    const int smem = MM_A_BYTES + (int)plane + MM_H_BYTES + 1024;
    if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(kmeans_mma_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     MM_A_BYTES + PlaneCfg<10>::B_BYTES + MM_H_BYTES + 1024));
        VT_CUDA(cudaFuncSetAttribute(kmeans_mma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     MM_A_BYTES + PlaneCfg<16>::B_BYTES + MM_H_BYTES + 1024));
        attr_set = true;
    }
    const int64_t want = (n_members + MM_TILE - 1) / MM_TILE, cap = (int64_t)vt_sm_count() * 6;
    const int blocks = (int)(want < cap ? want : cap);
    if (k <= 10)
        kmeans_mma_kernel<10><<<blocks, MM_TILE, smem, (cudaStream_t)stream>>>(
            (const uint8_t *)d_desc, d_perm, d_parent, n_members, d_node_internal, k, d_child_centroids, bimg, cn, d_label,
            (unsigned long long *)d_sums, (unsigned long long *)d_counts, (unsigned long long *)d_changed,
            (unsigned long long *)d_stats);
    else
        kmeans_mma_kernel<16><<<blocks, MM_TILE, smem, (cudaStream_t)stream>>>(
            (const uint8_t *)d_desc, d_perm, d_parent, n_members, d_node_internal, k, d_child_centroids, bimg, cn, d_label,
            (unsigned long long *)d_sums, (unsigned long long *)d_counts, (unsigned long long *)d_changed,
            (unsigned long long *)d_stats);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

int64_t vt_mma_prepared_bytes(int k, int depth);
// Batches of at least this many rows take TWO tree levels per launch (vt_quantize_pair.cu).  Off by default: the
// kernel halves the DRAM traffic of the descent (3 row reads instead of 6) but its per-chunk bookkeeping is latency
// bound and on one B200 it is still behind the one-level kernel at 100 M rows.  VT_MMA_PAIR_MIN
// in the environment or vt_quantize_mma_pair_min() switch it on.
static int64_t g_pair_min = (int64_t)1 << 62;
extern "C" int64_t vt_quantize_mma_pair_min(int64_t min_rows)
{
    const int64_t old = g_pair_min;
    if (min_rows != 0) g_pair_min = min_rows > 0 ? min_rows : ((int64_t)1 << 62);
    return old;
}
int vt_mma_levels_paired(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id,
                         int k, int depth, int32_t *leaf, int32_t *word, uint64_t *stats, const uint8_t *bimg,
                         const int *cn32, int *cninfo, void *sortwork, bool defer, cudaStream_t s, cudaEvent_t *ev,
                         int *n_ev_io);

// ids of (row, child) pairs that were bucketed by row range (the deferred output of the last level)
int vt_mma_deferred_output(const uint32_t *rows, const uint32_t *child, int64_t n, const int32_t *ref_id, int32_t *leaf,
                           int32_t *word, cudaStream_t s)
{
    // one resident wave, grid-stride: all CTAs sweep the buckets together, so the rows being written at
    // any moment lie in one or two buckets
    const int64_t w2 = (n + 255) / 256, c2 = (int64_t)vt_sm_count() * 8;
    deferred_output_kernel<<<(int)(w2 < c2 ? w2 : c2), 256, 0, s>>>(rows, child, n, ref_id, leaf, word);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

static int64_t mma_int_nodes(int k, int depth)
{
    int64_t n = 0, p = 1;
    for (int l = 0; l < depth; ++l) { n += p; p *= k; }
    return n;
}

extern "C" int64_t vt_quantize_mma_work_bytes(int64_t n, int k, int depth)
{
    if (n < 1) n = 1;
    int64_t n_heap = 0, p = 1;
    for (int l = 0; l <= depth; ++l) { n_heap += p; p *= k; }
    (void)n_heap;
    return vt_mma_prepared_bytes(k, depth) + 16 * n + vt_sort_scratch_bytes(n);
}

// byte planes + fixed-point norms for the whole tree (once per call / per tree)
static int mma_slots(int k) { return k <= 10 ? 10 : 16; }      // child slots per byte plane (PlaneCfg)
static int64_t mma_plane_bytes(int k) { return k <= 10 ? PlaneCfg<10>::B_BYTES : PlaneCfg<16>::B_BYTES; }

int vt_mma_prepare(const float *cent, int k, int depth, void *planes_and_norms, uint64_t *stats, cudaStream_t s)
{
    const int64_t n_int = mma_int_nodes(k, depth);
    uint8_t *bimg = (uint8_t *)planes_and_norms;
    const int64_t planes = ((n_int * mma_plane_bytes(k) + 255) / 256) * 256;
    unsigned long long *cn = (unsigned long long *)(bimg + planes);
    int64_t n_heap = 0, pw = 1;
    for (int l = 0; l <= depth; ++l) { n_heap += pw; pw *= k; }
    int *cn32 = (int *)(bimg + planes + ((n_heap * 8 + 255) / 256) * 256);
    const int64_t cap = (int64_t)vt_sm_count() * 16;
    const int blocks = (int)(n_int < cap ? n_int : cap);
    if (mma_slots(k) == 10)
        prep_planes_kernel<10><<<blocks, 128, 0, s>>>(cent, n_int, k, bimg, cn, cn32, (unsigned long long *)stats);
    else
        prep_planes_kernel<16><<<blocks, 128, 0, s>>>(cent, n_int, k, bimg, cn, cn32, (unsigned long long *)stats);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

int64_t vt_mma_prepared_bytes(int k, int depth)
{
    int64_t n_heap = 0, p = 1;
    for (int l = 0; l <= depth; ++l) { n_heap += p; p *= k; }
    return ((mma_int_nodes(k, depth) * mma_plane_bytes(k) + 255) / 256) * 256 + ((n_heap * 8 + 255) / 256) * 256 +
           ((mma_int_nodes(k, depth) * 64 + 255) / 256) * 256 +  // planes, 64-bit norms, 32-bit norms (16 per node)
           mma_int_nodes(k, depth) * 80;                         // 80-byte node records (vt_quantize_pair.cu)
}

// the level loop; `prepared` from vt_mma_prepare, `sortwork` = 16 n + vt_sort_scratch_bytes(n) bytes
int vt_mma_levels(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id, int k,
                  int depth, int32_t *leaf, int32_t *word, uint64_t *stats, const void *prepared, void *sortwork,
                  cudaStream_t s, cudaEvent_t *ev, int *n_ev_io)
{
    int n_ev = n_ev_io ? *n_ev_io : 0;
#define VT_MARK() do { if (ev) VT_CUDA(cudaEventRecord(ev[n_ev++], s)); } while (0)
    int64_t n_heap = 0, pw = 1;
    for (int l = 0; l <= depth; ++l) { n_heap += pw; pw *= k; }
    VT_CHECK_ARG(n_heap < ((int64_t)1 << 31), "tree too large for 32-bit node ids");
    const int64_t n_int = mma_int_nodes(k, depth);
    const uint8_t *bimg = (const uint8_t *)prepared;
    const int64_t planes = ((n_int * mma_plane_bytes(k) + 255) / 256) * 256;
    const int *cn = (const int *)(bimg + planes + ((n_heap * 8 + 255) / 256) * 256);     // 32-bit norms
    uint32_t *kA = (uint32_t *)sortwork, *kB = kA + n, *pA = kB + n, *pB = pA + n;
    void *hist = (void *)(pB + n);
    const uint8_t *desc = (const uint8_t *)d_desc;
    static bool attr_set = false;
    static int ctas10 = 8, ctas16 = 7;                           // resident CTAs per SM the grid is sized for (k <= 10 / k <= 16)
    static int64_t defer_min = (int64_t)1 << 22;                 // batches from this size on defer the last level's output
    const bool narrow = mma_slots(k) == 10;
    const int smem = MM_A_BYTES + (int)mma_plane_bytes(k) + 1024;
    if (!attr_set) {
        const char *e = getenv("VT_MMA_CTAS");
        if (e && atoi(e) >= 1 && atoi(e) <= 8) { ctas10 = atoi(e); ctas16 = atoi(e) < 7 ? atoi(e) : 7; }
        e = getenv("VT_MMA_DEFER_MIN");
        if (e) defer_min = atoll(e);
        e = getenv("VT_MMA_PAIR_MIN");                            // <= 0 switches the two-level kernel off
        if (e) g_pair_min = atoll(e) > 0 ? atoll(e) : ((int64_t)1 << 62);
        VT_CUDA(cudaFuncSetAttribute(level_mma_kernel<7, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     MM_A_BYTES + PlaneCfg<16>::B_BYTES + 1024));
        VT_CUDA(cudaFuncSetAttribute(level_mma_kernel<8, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     MM_A_BYTES + PlaneCfg<10>::B_BYTES + 1024));
        attr_set = true;
    }
    if (n >= g_pair_min && depth >= 2) {
        // two levels per launch (vt_quantize_pair.cu): every row is read from DRAM once per PAIR of levels
        // (the per-node records live behind the 32-bit norms of the prepared buffer; they are rewritten on every call, so the
        // buffer only has to be large enough -- vt_mma_prepared_bytes)
        int *cninfo = (int *)((uint8_t *)const_cast<int *>(cn) + ((n_int * 64 + 255) / 256) * 256);
        const int rc = vt_mma_levels_paired(d_desc, n, cent, internal, ref_id, k, depth, leaf, word, stats, bimg, cn, cninfo,
                                            sortwork, n >= defer_min, s, ev, &n_ev);
        if (n_ev_io) *n_ev_io = n_ev;
        return rc;
    }
    const int64_t want = (n + MM_TILE - 1) / MM_TILE, cap = (int64_t)vt_sm_count() * (narrow ? ctas10 : ctas16);
    const int blocks = (int)(want < cap ? want : cap);
    uint32_t *keys = kA, *perm = pA, *tkeys = kB, *tperm = pB;
    const uint32_t *cur_perm = nullptr, *cur_parent = nullptr;
    const int bits = path_bits(k);
    const bool defer = n >= defer_min;
    const uint32_t dmask = (1u << bits) - 1u;                    // bits <= 5 for k <= 16
    for (int level = 0; level < depth; ++level) {
        const bool last = level + 1 == depth;
        const int is_last = last ? (defer ? 2 : 1) : 0;
        uint32_t *keys_out = last && !defer ? nullptr : (level == 0 ? keys : tkeys);
        uint32_t *perm_out = last && !defer ? nullptr : (level == 0 ? perm : tperm);
        if (level > 0) VT_MARK();
        if (narrow)
            level_mma_kernel<8, 10><<<blocks, MM_TILE, smem, s>>>(desc, cur_perm, cur_parent, n, level, bits, n_heap, k,
                                                                  is_last, cent, internal, ref_id, bimg, cn, keys_out,
                                                                  perm_out, leaf, word, (unsigned long long *)stats);
        else
            level_mma_kernel<7, 16><<<blocks, MM_TILE, smem, s>>>(desc, cur_perm, cur_parent, n, level, bits, n_heap, k,
                                                                  is_last, cent, internal, ref_id, bimg, cn, keys_out,
                                                                  perm_out, leaf, word, (unsigned long long *)stats);
        VT_LAUNCH_CHECK();
        if (last && defer) {
            VT_MARK();                                           // (the deferred output is timed as the last "regrouping" stage)
            // (row, child) pairs -> 256 buckets of consecutive rows -> ids written bucket by bucket
            uint32_t *ok = level == 0 ? tkeys : keys, *op = level == 0 ? tperm : perm;
            int nb = 0;
            while (((int64_t)1 << nb) < n) ++nb;
            const int rc = vt_sort_pass_impl(keys_out, perm_out, ok, op, n, nb > 8 ? nb - 8 : 0, 0xffu, hist, s);
            if (rc != VT_OK) return rc;
            // one resident wave, grid-stride: all CTAs sweep the buckets together, so the rows being written at
            // any moment lie in one or two buckets
            const int64_t w2 = (n + 255) / 256, c2 = (int64_t)vt_sm_count() * 8;
            deferred_output_kernel<<<(int)(w2 < c2 ? w2 : c2), 256, 0, s>>>(ok, op, n, ref_id, leaf, word);
            VT_LAUNCH_CHECK();
        }
        VT_MARK();
        if (last) break;
        if (level > 0) { uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t; }
        // The rows are grouped by node already, so ONE stable pass over the newest digit of the packed paths
        // regroups them by child node (digit-major: the rows of one (node, child) pair stay contiguous).
        // Few long runs per tile -> long coalesced segments in the scatter.
        const int rc = vt_sort_pass_impl(keys, perm, tkeys, tperm, n, 0, dmask, hist, s);
        if (rc != VT_OK) return rc;
        { uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t; }
        cur_parent = keys;
        cur_perm = perm;
    }
#undef VT_MARK
    if (n_ev_io) *n_ev_io = n_ev;
    return VT_OK;
}

int vt_quantize_mma_impl(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id,
                         int k, int depth, int32_t *leaf, int32_t *word, uint64_t *stats, void *work, cudaStream_t s,
                         cudaEvent_t *ev, int *n_ev_out = nullptr)
{
    int n_ev = 0;
    int rc = vt_mma_prepare(cent, k, depth, work, stats, s);
    if (rc != VT_OK) return rc;
    if (ev) VT_CUDA(cudaEventRecord(ev[n_ev++], s));           // stages start after the plane preparation
    rc = vt_mma_levels(d_desc, n, cent, internal, ref_id, k, depth, leaf, word, stats, work,
                       (uint8_t *)work + vt_mma_prepared_bytes(k, depth), s, ev, &n_ev);
    if (n_ev_out) *n_ev_out = n_ev;
    return rc;
}

// Same contract as vt_quantize_sorted (bit-identical ids); tensor-core distance step.
extern "C" int vt_quantize_mma(const void *d_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                               const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth,
                               int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, void *d_work,
                               int64_t work_bytes, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8 && dp == 128, "the tensor-core descent takes 128-byte descriptor rows");
    VT_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "n out of range");
    VT_CHECK_ARG(k >= 1 && k <= 16 && depth >= 1 && depth <= 16, "k must be <= 16 and depth >= 1");
    VT_CHECK_ARG(d_centroids && d_internal, "null tree");
    VT_CHECK_ARG(d_word == nullptr || d_ref_id != nullptr, "d_word needs d_ref_id");
    VT_CHECK_ARG(d_work && work_bytes >= vt_quantize_mma_work_bytes(n, k, depth), "workspace too small");
    VT_CHECK_ARG((depth - 1) * path_bits(k) <= 32, "k / depth too large for 32-bit packed paths");
    if (n == 0) return VT_OK;
    return vt_quantize_mma_impl(d_desc, n, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word, d_stats, d_work,
                                (cudaStream_t)stream, nullptr);
}

// measurement variant: h_level_ms[depth] (the level kernel launches alone; the per-call plane preparation is not timed),
// h_sort_ms[depth] (re-sorts)
extern "C" int vt_quantize_mma_profile(const void *d_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                                       const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth,
                                       int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, void *d_work,
                                       int64_t work_bytes, float *h_level_ms, float *h_sort_ms, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8 && dp == 128 && n > 0 && k <= 16 && depth >= 1 && depth <= 16, "bad arguments");
    VT_CHECK_ARG(d_work && work_bytes >= vt_quantize_mma_work_bytes(n, k, depth), "workspace too small");
    VT_CHECK_ARG(h_level_ms && h_sort_ms, "null output");
    cudaEvent_t ev[40];
    const int n_ev = 2 * depth + 1;
    for (int i = 0; i < n_ev; ++i) VT_CUDA(cudaEventCreate(&ev[i]));
    int n_used = 0;
    int rc = vt_quantize_mma_impl(d_desc, n, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word, d_stats,
                                  d_work, (cudaStream_t)stream, ev, &n_used);
    // (the two-levels-per-launch path records one stage per launch: entries beyond its stages stay 0)
    if (rc == VT_OK) {
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            vt_set_error("vt_quantize_mma_profile: %s", cudaGetErrorString(e));
            rc = VT_ECUDA;
        }
    }
    for (int l = 0; l < depth && rc == VT_OK; ++l) {
        h_level_ms[l] = 0.f;
        h_sort_ms[l] = 0.f;
        if (2 * l + 1 < n_used) cudaEventElapsedTime(&h_level_ms[l], ev[2 * l], ev[2 * l + 1]);
        if (2 * l + 2 < n_used) cudaEventElapsedTime(&h_sort_ms[l], ev[2 * l + 1], ev[2 * l + 2]);
    }
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

// ---- planes prepared once per tree (they only depend on the centroids) -----------------------------------
extern "C" int64_t vt_quantize_mma_planes_bytes(int k, int depth) { return vt_mma_prepared_bytes(k, depth); }

extern "C" int vt_quantize_mma_prepare(const float *d_centroids, int k, int depth, void *d_planes, uint64_t *d_stats,
                                       void *stream)
{
    VT_CHECK_ARG(d_centroids && d_planes && k >= 1 && k <= 16 && depth >= 1 && depth <= 16, "bad arguments");
    return vt_mma_prepare(d_centroids, k, depth, d_planes, d_stats, (cudaStream_t)stream);
}

// vt_quantize_mma with planes from vt_quantize_mma_prepare; d_work: 16 n + vt_sort_scratch_bytes(n) bytes
extern "C" int vt_quantize_mma_prepared(const void *d_desc, int dtype, int64_t n, int dp, const float *d_centroids,
                                        const uint8_t *d_internal, const int32_t *d_ref_id, int k, int depth,
                                        int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, const void *d_planes,
                                        void *d_work, int64_t work_bytes, void *stream)
{
    VT_CHECK_ARG(dtype == VT_U8 && dp == 128, "the tensor-core descent takes 128-byte descriptor rows");
    VT_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "n out of range");
    VT_CHECK_ARG(k >= 1 && k <= 16 && depth >= 1 && depth <= 16, "k must be <= 16 and depth >= 1");
    VT_CHECK_ARG(d_centroids && d_internal && d_planes, "null tree / planes");
    VT_CHECK_ARG(d_word == nullptr || d_ref_id != nullptr, "d_word needs d_ref_id");
    VT_CHECK_ARG(d_work && work_bytes >= 16 * (n > 0 ? n : 1) + vt_sort_scratch_bytes(n), "workspace too small");
    VT_CHECK_ARG((depth - 1) * path_bits(k) <= 32, "k / depth too large for 32-bit packed paths");
    if (n == 0) return VT_OK;
    return vt_mma_levels(d_desc, n, d_centroids, d_internal, d_ref_id, k, depth, d_leaf_heap, d_word, d_stats, d_planes,
                         d_work, (cudaStream_t)stream, nullptr, nullptr);
}
