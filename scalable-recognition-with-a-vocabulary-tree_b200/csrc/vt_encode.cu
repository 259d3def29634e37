// vt_encode.cu -- bag-of-words encoding of a batch of images (VocabularyTree.propagate +
// embedding, cbir/encoders/vocabulary.py:78-137) and inverted-file construction.
//
// One CTA per image: the image's stop-node ids (reference DFS ids) are sorted in shared memory
// (bitonic), run-length encoded into (node, tf) pairs, and -- in the reference's "all nodes"
// semantics (vocabulary.py:92-100 counts every node on the path) -- lifted level by level:
// because reference ids are DFS pre-order, the parents of an id-sorted list are still sorted, so
// every level is one parent lookup + one adjacent-merge.  HBM traffic per image is the 4 B/word
// read plus the 8 B/entry CSR write; everything else stays in shared memory.
#include "vt_common.cuh"

namespace {
constexpr int ENC_THREADS = 256;
constexpr int ENC_WARPS = ENC_THREADS / 32;
constexpr int ENC_MAX = 8192;

struct ScanScratch { uint32_t wsum[ENC_WARPS]; };

// in-place exclusive scan of a[0..len) (uint32, shared memory); returns the total to every thread
__device__ uint32_t block_excl_scan(uint32_t *a, int len, ScanScratch &sc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int base = 0; base < len; base += ENC_THREADS) {
        const int idx = base + threadIdx.x;
        const uint32_t v = idx < len ? a[idx] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sc.wsum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < ENC_WARPS; ++w) {
            const uint32_t t = sc.wsum[w];
            if (w < warp) woff += t;
            total += t;
        }
        if (idx < len) a[idx] = carry + woff + incl - v;
        carry += total;
        __syncthreads();
    }
    return carry;
}

__device__ void bitonic_sort(uint32_t *key, int p2)
{
    for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (p2 >> 1); t += ENC_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));        // index with bit `stride` clear
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint32_t a = key[lo], b = key[hi];
                if ((a > b) == up) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// Same network for p2 == E * ENC_THREADS keys, E keys per thread in registers (thread t owns keys t*E .. t*E+E-1):
// strides below E are compare-exchanges between registers, strides below 32 E one shuffle per key, only the strides that
// cross warps go through shared memory (6 of the 55 stages at 1024 keys) -- 5.7 k instead of 24 k warp instructions per
// 1024-word image (ncu: the shared-memory network was 70 % of encode_kernel's instructions).  The sorted keys end up in
// `key`; `tmp` is a second array of p2 words.
template <int E>
__device__ void bitonic_sort_regs(uint32_t *key, uint32_t *tmp)
{
    constexpr int P2 = E * ENC_THREADS;
    const int t = threadIdx.x;
    uint32_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = key[t * E + r];
    uint32_t *src = key, *dst = tmp;
#pragma unroll 1
    for (int size = 2; size <= P2; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride >= 32 * E; stride >>= 1) {          // partner in another warp
#pragma unroll
            for (int r = 0; r < E; ++r) dst[t * E + r] = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int i = t * E + r;
                const uint32_t o = dst[i ^ stride];
                const bool keep_min = ((i & stride) == 0) == ((i & size) == 0);
                v[r] = keep_min ? min(v[r], o) : max(v[r], o);
            }
            uint32_t *sw = src; src = dst; dst = sw;                            // (alternate buffers: one barrier per stage)
        }
        for (int stride = min(size >> 1, 16 * E); stride >= E; stride >>= 1) {  // partner in another lane
            const int lm = stride / E;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int i = t * E + r;
                const uint32_t o = __shfl_xor_sync(VT_FULL, v[r], lm);
                const bool keep_min = ((i & stride) == 0) == ((i & size) == 0);
                v[r] = keep_min ? min(v[r], o) : max(v[r], o);
            }
        }
#pragma unroll
        for (int stride = E >> 1; stride > 0; stride >>= 1) {                   // partner in a register
            if (stride < size) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & stride) == 0) {
                        const int i = t * E + r;
                        const bool up = (i & size) == 0;
                        const uint32_t a = v[r], b = v[r | stride];
                        const uint32_t lo = min(a, b), hi = max(a, b);
                        v[r] = up ? lo : hi;
                        v[r | stride] = up ? hi : lo;
                    }
                }
            }
        }
    }
    __syncthreads();                                                            // everyone is done reading both buffers
#pragma unroll
    for (int r = 0; r < E; ++r) key[t * E + r] = v[r];
    __syncthreads();
}

// CAP: shared-memory capacity in words per array (power of two)
template <int CAP>
__global__ void __launch_bounds__(ENC_THREADS)
encode_kernel(const int32_t *__restrict__ word, const int64_t *__restrict__ img_off, int64_t n_img, int all_nodes,
              const int32_t *__restrict__ parent_ref, const uint8_t *__restrict__ level_ref, int depth,
              int32_t *__restrict__ nnz_out, const int64_t *__restrict__ row_off, int32_t *__restrict__ out_node,
              int32_t *__restrict__ out_count, unsigned long long *__restrict__ sumsq, uint32_t *__restrict__ df)
{
    extern __shared__ uint32_t smem[];
    uint32_t *A = smem;                 // ids
    uint32_t *C = smem + CAP;           // counts
    uint32_t *A2 = smem + 2 * CAP;      // scratch: flags / positions / merged ids
    uint32_t *C2 = smem + 3 * CAP;      // scratch: merged counts
    __shared__ ScanScratch sc;
    __shared__ unsigned long long s_sq;
    const bool fill = out_node != nullptr;

    for (int64_t img = blockIdx.x; img < n_img; img += gridDim.x) {
        const int64_t beg = img_off[img];
        const int n = (int)(img_off[img + 1] - beg);
        int64_t out_base = fill ? row_off[img] : 0;
        uint32_t emitted = 0;
        if (threadIdx.x == 0) s_sq = 0ull;
        if (n == 0) {
            if (threadIdx.x == 0) { if (nnz_out) nnz_out[img] = 0; if (fill && sumsq) sumsq[img] = 0ull; }
            __syncthreads();
            continue;
        }
        int p2 = 32;
        while (p2 < n) p2 <<= 1;
        for (int t = threadIdx.x; t < p2; t += ENC_THREADS) A[t] = t < n ? (uint32_t)word[beg + t] : 0xffffffffu;
        __syncthreads();
        if (p2 == 4 * ENC_THREADS) bitonic_sort_regs<4>(A, A2);
        else if (p2 == 8 * ENC_THREADS && CAP >= 8 * ENC_THREADS) bitonic_sort_regs<8>(A, A2);
        else bitonic_sort(A, p2);
        // run-length encode A[0..n) -> (A2 ids, C counts), m distinct
        for (int t = threadIdx.x; t < n; t += ENC_THREADS) C2[t] = (t == 0 || A[t] != A[t - 1]) ? 1u : 0u;
        __syncthreads();
        int m = (int)block_excl_scan(C2, n, sc);                       // C2[t] = output slot of t if head
        for (int t = threadIdx.x; t < n; t += ENC_THREADS)
            if (t == 0 || A[t] != A[t - 1]) { A2[C2[t]] = A[t]; C[C2[t]] = (uint32_t)t; }   // C = run start
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += ENC_THREADS) {
            const uint32_t start = C[t];
            const uint32_t end = t + 1 < m ? C[t + 1] : (uint32_t)n;
            C2[t] = end - start;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += ENC_THREADS) { A[t] = A2[t]; C[t] = C2[t]; }
        __syncthreads();

        // emit levels from the deepest up; without all_nodes emit everything at once.  all_nodes = m >= 1
        // keeps the levels >= m-1 only (level mask: the levels above carry weight 0)
        unsigned long long sq = 0ull;
        const int first_level = all_nodes ? depth : 0;
        const int min_level = all_nodes ? all_nodes - 1 : 0;
        for (int dd = first_level; dd >= 0; --dd) {
            // flag entries that leave the working list at this step
            for (int t = threadIdx.x; t < m; t += ENC_THREADS)
                A2[t] = (!all_nodes || level_ref[A[t]] == dd) ? 1u : 0u;
            __syncthreads();
            const uint32_t n_emit = block_excl_scan(A2, m, sc);
            for (int t = threadIdx.x; t < m; t += ENC_THREADS) {
                const bool mine = !all_nodes || level_ref[A[t]] == dd;
                if (mine) {
                    const uint32_t cnt = C[t];
                    sq += (unsigned long long)cnt * cnt;
                    if (fill) {
                        const int64_t o = out_base + emitted + A2[t];
                        out_node[o] = (int32_t)A[t];
                        out_count[o] = (int32_t)cnt;
                        if (df) atomicAdd(&df[A[t]], 1u);
                    }
                }
            }
            emitted += n_emit;
            __syncthreads();
            if (!all_nodes || dd <= min_level) break;
            // lift: entries of level dd move to their parent, then merge adjacent equal ids
            for (int t = threadIdx.x; t < m; t += ENC_THREADS)
                if (level_ref[A[t]] == dd) A[t] = (uint32_t)parent_ref[A[t]];
            __syncthreads();
            for (int t = threadIdx.x; t < m; t += ENC_THREADS) A2[t] = (t == 0 || A[t] != A[t - 1]) ? 1u : 0u;
            __syncthreads();
            const int m2 = (int)block_excl_scan(A2, m, sc);            // A2[t]: slot of the run t starts
            for (int t = threadIdx.x; t < m2; t += ENC_THREADS) C2[t] = 0u;
            __syncthreads();
            for (int t = threadIdx.x; t < m; t += ENC_THREADS) {
                const bool head = (t == 0 || A[t] != A[t - 1]);
                const uint32_t slot = head ? A2[t] : A2[t] - 1u;       // exclusive scan: non-heads see slot+1
                atomicAdd(&C2[slot], C[t]);
            }
            __syncthreads();
            // C is consumed: reuse it to stage the merged id list, then swap both lists in
            for (int t = threadIdx.x; t < m; t += ENC_THREADS)
                if (t == 0 || A[t] != A[t - 1]) C[A2[t]] = A[t];
            __syncthreads();
            for (int t = threadIdx.x; t < m2; t += ENC_THREADS) { A[t] = C[t]; C[t] = C2[t]; }
            __syncthreads();
            m = m2;
        }
        // per-image statistics: warp sums first -- a 64-bit shared-memory atomicAdd is a compare-and-swap loop, and 256
        // threads spinning on one word were 25 % of the kernel's instructions (ncu)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(VT_FULL, sq, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_sq, sq);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (nnz_out) nnz_out[img] = (int32_t)emitted;
            if (fill && sumsq) sumsq[img] = s_sq;
        }
        __syncthreads();
    }
}

// ---- inverted file ------------------------------------------------------------------------------
// tf postings without a gather: the sort value carries the posting itself, (count << shard_bits) | local image
__global__ void posting_pack_kernel(const int64_t *__restrict__ row_off, const int32_t *__restrict__ node,
                                    const int32_t *__restrict__ count, int64_t n_img, int shard_bits, int64_t n_ref,
                                    uint32_t *__restrict__ keys, uint32_t *__restrict__ packed)
{
    const uint32_t shard_size = 1u << shard_bits;
    const uint32_t n_shards = (uint32_t)((n_img + shard_size - 1) >> shard_bits);
    for (int64_t img = blockIdx.x; img < n_img; img += gridDim.x) {
        const int64_t b = row_off[img], e = row_off[img + 1];
        const uint32_t shard = (uint32_t)(img >> shard_bits), local = (uint32_t)img & (shard_size - 1u);
        for (int64_t t = b + threadIdx.x; t < e; t += blockDim.x) {
            keys[t] = (uint32_t)node[t] * n_shards + shard;
            packed[t] = ((uint32_t)count[t] << shard_bits) | local;
        }
    }
}

// row pointers of the sorted keys: every key in (previous key, key] starts at p.  The sorted packed values are the
// postings themselves.
__global__ void posting_offsets_kernel(const uint32_t *__restrict__ skeys, int64_t nnz, int64_t n_keys,
                                       uint32_t *__restrict__ post_off)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t key = skeys[p];
        const int64_t prev = p == 0 ? -1 : (int64_t)skeys[p - 1];
        for (int64_t kk = prev + 1; kk <= (int64_t)key; ++kk) post_off[kk] = (uint32_t)p;
        if (p == nnz - 1)
            for (int64_t kk = (int64_t)key + 1; kk <= n_keys; ++kk) post_off[kk] = (uint32_t)nnz;
    }
}

// rows written at loose offsets (in_off, lengths nnz) -> packed CSR (row_off)
__global__ void csr_compact_kernel(const int64_t *__restrict__ in_off, const int32_t *__restrict__ nnz,
                                   const int64_t *__restrict__ row_off, int64_t n_img, const int32_t *__restrict__ node_in,
                                   const int32_t *__restrict__ count_in, int32_t *__restrict__ node_out,
                                   int32_t *__restrict__ count_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t img = warp; img < n_img; img += n_warps) {
        const int64_t src = in_off[img], dst = row_off[img];
        const int len = nnz[img];
        for (int t = lane; t < len; t += 32) {
            node_out[dst + t] = node_in[src + t];
            count_out[dst + t] = count_in[src + t];
        }
    }
}

__global__ void fill_u32_kernel(uint32_t *a, int64_t n, uint32_t v)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) a[t] = v;
}
}  // namespace

extern "C" int vt_encode_max_words(void) { return ENC_MAX; }

template <int CAP>
static int launch_encode(const int32_t *word, const int64_t *img_off, int64_t n_img, int all_nodes,
                         const int32_t *parent_ref, const uint8_t *level_ref, int depth, int32_t *nnz,
                         const int64_t *row_off, int32_t *out_node, int32_t *out_count, uint64_t *sumsq, uint32_t *df,
                         cudaStream_t s)
{
    auto kern = encode_kernel<CAP>;
    const size_t smem = (size_t)4 * CAP * sizeof(uint32_t);
    VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)vt_sm_count() * (CAP <= 2048 ? 6 : 1);
    const int blocks = (int)(n_img < cap ? n_img : cap);
    kern<<<blocks, ENC_THREADS, smem, s>>>(word, img_off, n_img, all_nodes, parent_ref, level_ref, depth, nnz, row_off,
                                           out_node, out_count, (unsigned long long *)sumsq, df);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_encode(const int32_t *d_word, const int64_t *d_img_off, int64_t n_img, int32_t max_words,
                         int all_nodes, const int32_t *d_parent_ref, const uint8_t *d_level_ref, int depth,
                         int32_t *d_nnz, const int64_t *d_row_off, int32_t *d_out_node, int32_t *d_out_count,
                         uint64_t *d_sumsq, uint32_t *d_df, void *stream)
{
    VT_CHECK_ARG(n_img >= 0 && d_img_off, "bad image offsets");
    VT_CHECK_ARG(max_words >= 0 && max_words <= ENC_MAX, "an image has more descriptors than vt_encode_max_words()");
    VT_CHECK_ARG(all_nodes >= 0 && all_nodes <= depth + 1, "all_nodes must be 0 (stop nodes) or 1 + first kept level");
    VT_CHECK_ARG(!all_nodes || (d_parent_ref && d_level_ref), "all_nodes needs parent/level tables");
    VT_CHECK_ARG((d_out_node == nullptr) == (d_out_count == nullptr), "out_node/out_count must come together");
    VT_CHECK_ARG(d_out_node != nullptr || d_nnz != nullptr, "count pass needs d_nnz");
    VT_CHECK_ARG(d_out_node == nullptr || d_row_off != nullptr, "fill pass needs d_row_off");
    if (n_img == 0) return VT_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define VT_E(CAPV) return launch_encode<CAPV>(d_word, d_img_off, n_img, all_nodes, d_parent_ref, d_level_ref, depth, d_nnz, d_row_off, d_out_node, d_out_count, d_sumsq, d_df, s)
    if (max_words <= 1024) VT_E(1024);
    if (max_words <= 2048) VT_E(2048);
    VT_E(8192);
#undef VT_E
}

// The inverted file: node-major, sharded by image range.  The value sorted along with the key (node * n_shards + shard)
// IS the posting, (count << log2(shard_size)) | local image id, so building the inverted file needs no gather through
// entry indices and a posting is 4 bytes.  shard_size must be a power of two and counts must fit the remaining bits
// (vt_encode_max_words() = 8192 descriptors per image do).
extern "C" int vt_posting_pack(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_img,
                               int32_t shard_size, int64_t n_ref, uint32_t *d_keys, uint32_t *d_packed, void *stream)
{
    VT_CHECK_ARG(n_img >= 0 && shard_size > 0 && (shard_size & (shard_size - 1)) == 0 && n_ref > 0, "bad sizes");
    int shard_bits = 0;
    while ((1 << shard_bits) < shard_size) ++shard_bits;
    VT_CHECK_ARG(ENC_MAX < (1 << (32 - shard_bits)), "counts do not fit beside the local image id");
    const int64_t n_shards = (n_img + shard_size - 1) / shard_size;
    VT_CHECK_ARG(n_shards * n_ref < ((int64_t)1 << 32), "shards * nodes must fit 32 bits");
    if (n_img == 0) return VT_OK;
    VT_CHECK_ARG(d_row_off && d_node && d_count && d_keys && d_packed, "null buffer");
    const int64_t cap = (int64_t)vt_sm_count() * 16;
    posting_pack_kernel<<<(int)(n_img < cap ? n_img : cap), 128, 0, (cudaStream_t)stream>>>(
        d_row_off, d_node, d_count, n_img, shard_bits, n_ref, d_keys, d_packed);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// row pointers d_post_off [n_keys + 1] of the sorted keys; the sorted packed values are the posting array as they are
extern "C" int vt_posting_offsets(const uint32_t *d_sorted_keys, int64_t nnz, int64_t n_keys, uint32_t *d_post_off,
                                  void *stream)
{
    VT_CHECK_ARG(nnz >= 0 && nnz < ((int64_t)1 << 32) && n_keys >= 0 && d_post_off, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (nnz == 0) {
        const int64_t cap = (int64_t)vt_sm_count() * 16, want = (n_keys + 1 + 255) / 256;
        fill_u32_kernel<<<(int)(want < cap ? want : cap), 256, 0, s>>>(d_post_off, n_keys + 1, 0u);
        VT_LAUNCH_CHECK();
        return VT_OK;
    }
    VT_CHECK_ARG(d_sorted_keys, "null buffer");
    const int64_t cap = (int64_t)vt_sm_count() * 16, want = (nnz + 255) / 256;
    posting_offsets_kernel<<<(int)(want < cap ? want : cap), 256, 0, s>>>(d_sorted_keys, nnz, n_keys, d_post_off);
    VT_LAUNCH_CHECK();
    return VT_OK;
}


// Single-pass encoding: vt_encode may be called ONCE with d_row_off = d_img_off (an image never has more entries
// than descriptors when only stop nodes are counted) and d_nnz set; this packs the loosely placed rows.
extern "C" int vt_csr_compact(const int64_t *d_in_off, const int32_t *d_nnz, const int64_t *d_row_off, int64_t n_img,
                              const int32_t *d_node_in, const int32_t *d_count_in, int32_t *d_node_out,
                              int32_t *d_count_out, void *stream)
{
    VT_CHECK_ARG(n_img >= 0, "bad sizes");
    if (n_img == 0) return VT_OK;
    VT_CHECK_ARG(d_in_off && d_nnz && d_row_off && d_node_in && d_count_in && d_node_out && d_count_out, "null buffer");
    const int64_t cap = (int64_t)vt_sm_count() * 8, want = (n_img * 32 + 255) / 256;
    csr_compact_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        d_in_off, d_nnz, d_row_off, n_img, d_node_in, d_count_in, d_node_out, d_count_out);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
