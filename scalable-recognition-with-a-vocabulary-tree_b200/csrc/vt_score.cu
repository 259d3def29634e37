// vt_score.cu -- inverted-file scoring + top-k (Database.score / Database.retrieve,
// cbir/database.py:59-81), batched over queries.
//
// The database is sharded by image range (shard_size images, the same rule that shards it across
// GPUs).  The kernels of THIS file give every (query, shard) pair a CTA of its own and write one top-k list per pair:
// lanes walk the posting lists of the query's words inside that shard (node-major CSR, 4 B per posting: tf << 13 | local
// image id), accumulate the sparse dot products of all shard images in shared memory, and select the shard's top-k --
// score_shard_tf_kernel through an exact integer bound from a histogram of the dots, score_shard_kernel exhaustively
// (and with the full key dump `retrieve()` needs).  Since round 2 the DEFAULT scorer is vt_score_query.cu (a CTA carries a
// query's top-k from shard to shard); these kernels remain for the full key dump, for indexes built without the shard
// bounds, and as the per-shard reference the carried kernels are tested against.  topk_merge_kernel merges the lists of a
// query (per shard or per group) and turns keys into the reference's score (database.py:66-70: L2 distance of the
// L2-normalised vectors, NaN -> 1e6).
// Measured on the round-1 version (8-byte postings; ncu, 1 M images, leaf words): 2.7x the algorithmic traffic -- a
// (word, shard) run is ~8 postings, so 32 B sectors around the runs and 8 B pointer pairs dominate -- at 3.2 TB/s, issue
// slots 27 % busy: bound by the dependent DRAM round trips per word, not by the shared-memory atomics.
//
// Rank key (larger is better), ties broken by smaller image index like the reference's stable
// ascending sort (database.py:79-80):
//   VT_SCORE_TF_L2 : dot(tf_q, tf_d) / ||tf_d||   (integer dot, exact)   score = sqrt(2 - 2 key/||tf_q||)
//   VT_SCORE_W_L2  : dot(w_q, w_d)   (weights pre-normalised)            score = sqrt(2 - 2 key)
//   VT_SCORE_W_L1  : -sum_common(|q-d| - q - d)                          score = 2 - key
// An image with no descriptors has key -1 (score 1e6).
#include "vt_common.cuh"

namespace {
constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_MAXK = 64;

__device__ __forceinline__ bool better(double ka, int32_t ia, double kb, int32_t ib)
{
    return ka > kb || (ka == kb && ia < ib);
}

// block-wide argmax of (key, id) over one candidate per thread; result broadcast to all threads
__device__ __forceinline__ void block_best(double &key, int32_t &id, int &owner, double *s_key, int32_t *s_id, int *s_owner)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    owner = threadIdx.x;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const double k2 = __shfl_xor_sync(VT_FULL, key, o);
        const int32_t i2 = __shfl_xor_sync(VT_FULL, id, o);
        const int o2 = __shfl_xor_sync(VT_FULL, owner, o);
        if (better(k2, i2, key, id)) { key = k2; id = i2; owner = o2; }
    }
    if (lane == 0) { s_key[warp] = key; s_id[warp] = id; s_owner[warp] = owner; }
    __syncthreads();
    key = s_key[0]; id = s_id[0]; owner = s_owner[0];
#pragma unroll
    for (int w = 1; w < SC_WARPS; ++w)
        if (better(s_key[w], s_id[w], key, id)) { key = s_key[w]; id = s_id[w]; owner = s_owner[w]; }
    __syncthreads();
}

constexpr uint32_t SC_LONG = 24;
constexpr int SC_UNROLL = 4;
constexpr int SC_FETCH = 8;         // postings fetched together by one lane

// A posting is ONE 32-bit word: (tf << 13) | local image index inside its shard of 8192 images -- exactly the value the
// index build sorts (vt_posting_pack), so the sorted array IS the posting array (4 B per posting instead of 8).
constexpr uint32_t SC_LOCAL_MASK = 8191u;
constexpr int SC_LOCAL_BITS = 13;

template <int MODE>
__device__ __forceinline__ void accumulate(int *s_acci, float *, const uint32_t pe, const float qv)
{
    static_assert(MODE == VT_SCORE_TF_L2, "weighted modes run in vt_weight.cu");
    atomicAdd(&s_acci[pe & SC_LOCAL_MASK], (int)qv * (int)(pe >> SC_LOCAL_BITS));
}

// one lane walks a short posting list: SC_FETCH postings are fetched before any is accumulated, so the
// DRAM round trips of a list overlap instead of serialising behind the shared-memory atomics
template <int MODE>
__device__ __forceinline__ void walk_short(int *s_acci, float *s_accf, const uint32_t *__restrict__ postings, uint32_t b,
                                           uint32_t e, float qv)
{
    for (uint32_t p = b; p < e; p += SC_FETCH) {
        uint32_t pe[SC_FETCH];
#pragma unroll
        for (int i = 0; i < SC_FETCH; ++i) pe[i] = p + i < e ? __ldg(postings + p + i) : 0u;
#pragma unroll
        for (int i = 0; i < SC_FETCH; ++i)
            if (p + i < e) accumulate<MODE>(s_acci, s_accf, pe[i], qv);
    }
}

// MODE: VT_SCORE_*.  R: images per shard.
template <int MODE, int R>
__global__ void __launch_bounds__(SC_THREADS)
score_shard_kernel(const uint32_t *__restrict__ post_off, const uint32_t *__restrict__ postings,
                   const double *__restrict__ img_rnorm, int64_t n_img, int64_t n_ref, int n_shards,
                   const int64_t *__restrict__ q_off, const int32_t *__restrict__ q_node,
                   const float *__restrict__ q_val, int64_t n_query, int topk, double *__restrict__ out_key,
                   int32_t *__restrict__ out_id, double *__restrict__ all_key, const uint8_t *__restrict__ shard_empty,
                   const int32_t *__restrict__ seed, int64_t seed_stride)
{
    constexpr int PER = R / SC_THREADS;
    extern __shared__ unsigned char s_raw[];
    double *s_keys = reinterpret_cast<double *>(s_raw);                        // [R]
    float *s_accf = reinterpret_cast<float *>(s_raw + (size_t)R * 8);          // [R] (aliased int/float)
    int *s_acci = reinterpret_cast<int *>(s_accf);
    __shared__ double r_key[SC_WARPS];
    __shared__ int32_t r_id[SC_WARPS];
    __shared__ int r_owner[SC_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int64_t pair = blockIdx.x; pair < n_query * n_shards; pair += gridDim.x) {
        const int64_t q = pair / n_shards;
        const int shard = (int)(pair - q * n_shards);
        const int64_t img0 = (int64_t)shard * R;
        const int n_here = (int)((n_img - img0) < R ? (n_img - img0) : R);
        // images the query does not touch have key 0 without looking at their norm -- unless the shard
        // holds images without descriptors (key -1, score 1e6), which only the norm reveals
        const bool lazy_norm = !all_key && !seed && shard_empty && !shard_empty[shard];
        // tf mode may be seeded with the dots over the dense top levels (vt_dense_dots): the postings add the rest
        for (int t = threadIdx.x; t < R; t += SC_THREADS)
            s_acci[t] = (seed && t < n_here) ? seed[(size_t)q * seed_stride + img0 + t] : 0;
        __syncthreads();
        // accumulate: one LANE per query word (32 independent row-pointer / posting fetches in flight
        // per warp -- the walk is latency bound, not bandwidth bound); posting lists longer than
        // SC_LONG are handed to the whole warp so that one lane never serialises a long list
        const int64_t qb = q_off[q], qe = q_off[q + 1];
        const uint32_t *off = post_off + shard;                 // row of (node, shard) = node * n_shards + shard
        // SC_UNROLL words per lane are in flight together: the q_node -> row pointer -> posting chain is
        // three dependent DRAM/L2 round trips, so independent chains are what hides the latency
        for (int64_t w0 = qb + (int64_t)warp * 32; w0 < qe; w0 += (int64_t)SC_THREADS * SC_UNROLL) {
            int32_t node[SC_UNROLL];
            float qv[SC_UNROLL];
            uint32_t b[SC_UNROLL], e[SC_UNROLL];
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                const int64_t w = w0 + (int64_t)u * SC_THREADS + lane;
                node[u] = w < qe ? __ldg(q_node + w) : -1;
                qv[u] = w < qe ? __ldg(q_val + w) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                b[u] = node[u] >= 0 ? __ldg(off + (size_t)node[u] * n_shards) : 0u;
                e[u] = node[u] >= 0 ? __ldg(off + (size_t)node[u] * n_shards + 1) : 0u;
            }
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                const bool is_long = (e[u] - b[u]) > SC_LONG;
                if (!is_long) walk_short<MODE>(s_acci, s_accf, postings, b[u], e[u], qv[u]);
                unsigned int m = __ballot_sync(VT_FULL, is_long);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t lb = __shfl_sync(VT_FULL, b[u], src), le = __shfl_sync(VT_FULL, e[u], src);
                    const float lq = __shfl_sync(VT_FULL, qv[u], src);
                    for (uint32_t p = lb + lane; p < le; p += 32) accumulate<MODE>(s_acci, s_accf, __ldg(postings + p), lq);
                }
            }
        }
        __syncthreads();
        // rank keys + per-thread best
        double bkey = -2.0; int32_t bid = 0x7fffffff;
#pragma unroll 4
        for (int u = 0; u < PER; ++u) {
            const int t = u * SC_THREADS + threadIdx.x;
            double key = -2.0;                                                  // padding slots never win
            if (t < n_here) {
                if (MODE == VT_SCORE_TF_L2) {
                    const int a = s_acci[t];
                    if (a == 0 && lazy_norm) {
                        key = 0.0;
                    } else {
                        const double rn = img_rnorm[img0 + t];
                        key = rn < 0.0 ? -1.0 : (double)a * rn;
                    }
                } else {
                    const double rn = img_rnorm ? img_rnorm[img0 + t] : 1.0;
                    key = rn < 0.0 ? -1.0 : (double)s_accf[t];
                }
                if (all_key) all_key[q * n_img + img0 + t] = key;
            }
            s_keys[t] = key;
            if (better(key, (int32_t)(img0 + t), bkey, bid)) { bkey = key; bid = (int32_t)(img0 + t); }
        }
        __syncthreads();
        // k rounds of block argmax; only the winning thread rescans its slice
        for (int r = 0; r < topk; ++r) {
            double key = bkey; int32_t id = bid; int owner;
            block_best(key, id, owner, r_key, r_id, r_owner);
            if (threadIdx.x == 0) {
                const size_t o = ((size_t)q * n_shards + shard) * topk + r;
                out_key[o] = key;
                out_id[o] = key > -2.0 ? id : -1;
            }
            if (threadIdx.x == owner) {
                if (key > -2.0) s_keys[id - img0] = -2.0;
                bkey = -2.0; bid = 0x7fffffff;
                for (int u = 0; u < PER; ++u) {
                    const int t = u * SC_THREADS + threadIdx.x;
                    const double k2 = s_keys[t];
                    if (better(k2, (int32_t)(img0 + t), bkey, bid)) { bkey = k2; bid = (int32_t)(img0 + t); }
                }
            }
        }
        __syncthreads();
    }
}

// ---- TF / L2 fast path ---------------------------------------------------------------------------
// Same walk as score_shard_kernel, but the selection never touches most images: the dot products
// are small exact integers, so an integer bound discards everything that cannot reach the top-k.
//   T  = the k-th largest dot in the shard (exact, from a 64-bin histogram of the dots >= 2);
//   an image with dot < T * rho  (rho = min norm^-1 / max norm^-1 of the shard) has key
//   dot/|d| < T * min|d|^-1 <= the key of each of the >= k images with dot >= T  ->  not in the top-k.
// Survivors (typically a few dozen of 8192) get their binary64 key and one warp picks the k best.
// Shards with images without descriptors (rho < 0), T == 0 or more than SC_CAND survivors fall back
// to the exhaustive selection (keys recomputed on demand).
constexpr int SC_CAND = 256;

// MINB: resident CTAs per SM the register budget is set for (4: the lane-per-word walk of short lists keeps its
// postings in registers; 6: long lists are walked by whole warps and want occupancy instead)
template <int R, int MINB>
__global__ void __launch_bounds__(SC_THREADS, MINB)
score_shard_tf_kernel(const uint32_t *__restrict__ post_off, const uint32_t *__restrict__ postings,
                      const double *__restrict__ img_rnorm, int64_t n_img, int64_t n_ref, int n_shards,
                      const int64_t *__restrict__ q_off, const int32_t *__restrict__ q_node,
                      const float *__restrict__ q_val, int64_t n_query, int topk, double *__restrict__ out_key,
                      int32_t *__restrict__ out_id, const double *__restrict__ shard_rho,
                      const int32_t *__restrict__ seed, int64_t seed_stride, double *__restrict__ all_key)
{
    constexpr int PER = R / SC_THREADS;
    extern __shared__ unsigned char s_raw[];
    int *s_acc = reinterpret_cast<int *>(s_raw);                               // [R]
    double *s_ckey = reinterpret_cast<double *>(s_raw + (size_t)R * 4);        // [SC_CAND]
    int *s_cslot = reinterpret_cast<int *>(s_raw + (size_t)R * 4 + SC_CAND * 8);   // [SC_CAND]
    __shared__ int s_hist[64];
    __shared__ int s_thr, s_ncand;
    __shared__ double r_key[SC_WARPS];
    __shared__ int32_t r_id[SC_WARPS];
    __shared__ int r_owner[SC_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TAKEN = (int)0x80000000;

    for (int64_t pair = blockIdx.x; pair < n_query * n_shards; pair += gridDim.x) {
        const int64_t q = pair / n_shards;
        const int shard = (int)(pair - q * n_shards);
        const int64_t img0 = (int64_t)shard * R;
        const int n_here = (int)((n_img - img0) < R ? (n_img - img0) : R);
        // seeded (all-nodes indexes): the accumulators start from the int32 dots over the dense top levels (vt_dense_dots)
        for (int t = threadIdx.x; t < R; t += SC_THREADS)
            s_acc[t] = (seed && t < n_here) ? seed[(size_t)q * seed_stride + img0 + t] : 0;
        if (threadIdx.x == 0) s_ncand = 0;
        if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
        __syncthreads();
        const int64_t qb = q_off[q], qe = q_off[q + 1];
        const uint32_t *off = post_off + shard;                 // row of (node, shard) = node * n_shards + shard
        for (int64_t w0 = qb + (int64_t)warp * 32; w0 < qe; w0 += (int64_t)SC_THREADS * SC_UNROLL) {
            int32_t node[SC_UNROLL];
            float qv[SC_UNROLL];
            uint32_t b[SC_UNROLL], e[SC_UNROLL];
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                const int64_t w = w0 + (int64_t)u * SC_THREADS + lane;
                node[u] = w < qe ? __ldg(q_node + w) : -1;
                qv[u] = w < qe ? __ldg(q_val + w) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                b[u] = node[u] >= 0 ? __ldg(off + (size_t)node[u] * n_shards) : 0u;
                e[u] = node[u] >= 0 ? __ldg(off + (size_t)node[u] * n_shards + 1) : 0u;
            }
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                const bool is_long = (e[u] - b[u]) > SC_LONG;
                if (!is_long) walk_short<VT_SCORE_TF_L2>(s_acc, nullptr, postings, b[u], e[u], qv[u]);
                unsigned int m = __ballot_sync(VT_FULL, is_long);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t lb = __shfl_sync(VT_FULL, b[u], src), le = __shfl_sync(VT_FULL, e[u], src);
                    const float lq = __shfl_sync(VT_FULL, qv[u], src);
                    for (uint32_t p = lb + lane; p < le; p += 32) accumulate<VT_SCORE_TF_L2>(s_acc, nullptr, __ldg(postings + p), lq);
                }
            }
        }
        __syncthreads();

        // ---- integer bound -------------------------------------------------------------------------
        // exact histogram of the dots >= 2 (clamped at 63); T = the k-th largest dot (or 63)
        const double rho = shard_rho ? shard_rho[shard] : -1.0;
        int mine_acc[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int a = s_acc[u * SC_THREADS + threadIdx.x];
            mine_acc[u] = a;
            if (a >= 2 && !seed) atomicAdd(&s_hist[a < 63 ? a : 63], 1);      // (seeded dots all land in the last bin: skip)
        }
        __syncthreads();
        if (warp == 0) {
            // lanes walk the bins from the top: lane l owns bins 63-l and 31-l; inclusive prefix = images
            // with a dot at least that large
            int hi = s_hist[63 - lane], lo = lane < 30 ? s_hist[31 - lane] : 0;   // bins 0 and 1 are not counted
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t1 = __shfl_up_sync(VT_FULL, hi, o), t2 = __shfl_up_sync(VT_FULL, lo, o);
                if (lane >= o) { hi += t1; lo += t2; }
            }
            lo += __shfl_sync(VT_FULL, hi, 31);
            const unsigned int mh = __ballot_sync(VT_FULL, hi >= topk), ml = __ballot_sync(VT_FULL, lo >= topk);
            int thr = mh ? 63 - (__ffs(mh) - 1) : (ml ? 31 - (__ffs(ml) - 1) : 0);
            int d_lo = 0;
            if (rho > 0.0 && thr >= 2) {
                d_lo = (int)ceil((double)thr * rho);     // dot < d_lo  =>  dot <= d_lo - 1 < thr * rho: cannot win
                if (d_lo < 2) d_lo = 2;                  // (a bound of 1 keeps every touched image: use the fallback)
            }
            if (lane == 0) s_thr = d_lo;                 // 0 => exhaustive fallback
        }
        __syncthreads();
        const int d_lo = s_thr;
        // (seeded: every image shares the top levels with the query, the dots are ~1e6 and the 64-bin bound says nothing:
        // straight to the exhaustive selection, which recomputes keys on demand and keeps the shared memory at 4 B per image)
        bool fast = d_lo > 0 && !seed;
        if (fast) {
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const bool keep = mine_acc[u] >= d_lo;
                const unsigned int m = __ballot_sync(VT_FULL, keep);
                if (m) {                                                         // one atomic per warp and pass
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_ncand, __popc(m));
                    base = __shfl_sync(VT_FULL, base, 0);
                    const int pos = base + __popc(m & ((1u << lane) - 1u));
                    if (keep && pos < SC_CAND) s_cslot[pos] = u * SC_THREADS + threadIdx.x;
                }
            }
            __syncthreads();
            fast = s_ncand <= SC_CAND;                                           // block-uniform
        }
        if (fast) {
            const int nc = s_ncand;
            for (int i = threadIdx.x; i < nc; i += SC_THREADS) {
                const int t = s_cslot[i];
                s_ckey[i] = (double)s_acc[t] * __ldg(img_rnorm + img0 + t);
            }
            __syncthreads();
            if (warp == 0) {                                                     // one warp selects the k best survivors
                for (int r = 0; r < topk; ++r) {
                    double bk = -2.0; int32_t bid = 0x7fffffff; int bslot = -1;
                    for (int i = lane; i < nc; i += 32) {
                        const double k2 = s_ckey[i];
                        const int32_t i2 = (int32_t)(img0 + s_cslot[i]);
                        if (k2 > -2.0 && better(k2, i2, bk, bid)) { bk = k2; bid = i2; bslot = i; }
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const double k2 = __shfl_xor_sync(VT_FULL, bk, o);
                        const int32_t i2 = __shfl_xor_sync(VT_FULL, bid, o);
                        const int s2 = __shfl_xor_sync(VT_FULL, bslot, o);
                        if (better(k2, i2, bk, bid)) { bk = k2; bid = i2; bslot = s2; }
                    }
                    if (lane == 0) {
                        const size_t o = ((size_t)q * n_shards + shard) * topk + r;
                        // fewer survivors than k: the rest of the list is images the query does not touch
                        // (key 0) -- they are not candidates of the fast path, so the slot is reported
                        // empty; with >= k images of dot >= T this cannot happen for r < k
                        out_key[o] = bk;
                        out_id[o] = bk > -2.0 ? bid : -1;
                        if (bslot >= 0) s_ckey[bslot] = -2.0;
                    }
                    __syncwarp();
                }
            }
        } else {
            // ---- exhaustive fallback: every slot is a candidate, keys recomputed on demand ------------
            auto key_of = [&](int t) -> double {
                if (t >= n_here) return -2.0;
                const int a = s_acc[t];
                if (a == TAKEN) return -2.0;
                const double rn = __ldg(img_rnorm + img0 + t);
                return rn < 0.0 ? -1.0 : (double)a * rn;
            };
            double bkey = -2.0; int32_t bid = 0x7fffffff;
            for (int u = 0; u < PER; ++u) {
                const int t = u * SC_THREADS + threadIdx.x;
                const double key = key_of(t);
                if (all_key && t < n_here) all_key[q * n_img + img0 + t] = key;
                if (better(key, (int32_t)(img0 + t), bkey, bid)) { bkey = key; bid = (int32_t)(img0 + t); }
            }
            for (int r = 0; r < topk; ++r) {
                double key = bkey; int32_t id = bid; int owner;
                block_best(key, id, owner, r_key, r_id, r_owner);
                if (threadIdx.x == 0) {
                    const size_t o = ((size_t)q * n_shards + shard) * topk + r;
                    out_key[o] = key;
                    out_id[o] = key > -2.0 ? id : -1;
                }
                if (threadIdx.x == owner) {
                    if (key > -2.0) s_acc[id - img0] = TAKEN;
                    bkey = -2.0; bid = 0x7fffffff;
                    for (int u = 0; u < PER; ++u) {
                        const int t = u * SC_THREADS + threadIdx.x;
                        const double k2 = key_of(t);
                        if (better(k2, (int32_t)(img0 + t), bkey, bid)) { bkey = k2; bid = (int32_t)(img0 + t); }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// merge n_lists sorted-or-not candidate lists of a query into its top-k, compute final scores
template <int MODE>
__global__ void __launch_bounds__(SC_THREADS)
topk_merge_kernel(const double *__restrict__ in_key, const int32_t *__restrict__ in_id, int64_t n_query, int n_cand,
                  int topk, const double *__restrict__ q_rnorm, const unsigned long long *__restrict__ q_sumsq,
                  const unsigned long long *__restrict__ img_sumsq, double *__restrict__ out_key,
                  int32_t *__restrict__ out_id, double *__restrict__ out_score)
{
    extern __shared__ unsigned char s_raw[];
    double *s_keys = reinterpret_cast<double *>(s_raw);                      // [n_cand]
    int32_t *s_ids = reinterpret_cast<int32_t *>(s_raw + (size_t)n_cand * 8);
    __shared__ double r_key[SC_WARPS];
    __shared__ int32_t r_id[SC_WARPS];
    __shared__ int r_owner[SC_WARPS];
    for (int64_t q = blockIdx.x; q < n_query; q += gridDim.x) {
        for (int t = threadIdx.x; t < n_cand; t += SC_THREADS) {
            const int32_t id = in_id[q * n_cand + t];
            s_ids[t] = id;
            s_keys[t] = id >= 0 ? in_key[q * n_cand + t] : -2.0;
        }
        __syncthreads();
        for (int r = 0; r < topk; ++r) {
            double bkey = -2.0; int32_t bid = 0x7fffffff; int bslot = -1;
            for (int t = threadIdx.x; t < n_cand; t += SC_THREADS) {
                const double k2 = s_keys[t];
                const int32_t i2 = k2 > -2.0 ? s_ids[t] : 0x7fffffff;
                if (better(k2, i2, bkey, bid)) { bkey = k2; bid = i2; bslot = t; }
            }
            double key = bkey; int32_t id = bid; int owner;
            block_best(key, id, owner, r_key, r_id, r_owner);
            if (threadIdx.x == owner && key > -2.0) s_keys[bslot] = -2.0;
            if (threadIdx.x == 0) {
                const bool ok = key > -2.0;
                out_key[q * topk + r] = key;
                out_id[q * topk + r] = ok ? id : -1;
                if (out_score) {
                    double s = 1e6;                                         // database.py:70
                    if (ok && key >= 0.0) {
                        if (MODE == VT_SCORE_TF_L2) {
                            const double rq = q_rnorm[q];
                            if (rq >= 0.0) {
                                const double c = key * rq;
                                s = sqrt(fmax(0.0, 2.0 - 2.0 * c));
                                // identical count vectors: dot == |d|^2 == |q|^2 -> exactly 0 like the
                                // reference's d/|d| - q/|q| (Cauchy-Schwarz equality)
                                if (q_sumsq && img_sumsq && img_sumsq[id] == q_sumsq[q]) {
                                    const double dot = key * sqrt((double)img_sumsq[id]);
                                    if (fabs(dot - (double)q_sumsq[q]) < 0.5) s = 0.0;
                                }
                            }
                        } else if (MODE == VT_SCORE_W_L2) {
                            s = sqrt(fmax(0.0, 2.0 - 2.0 * key));
                            // float32 weights leave ~1e-7 noise on the cosine; identical bags (same tf
                            // norm, cosine 1 within that noise) score exactly 0 like in tf mode
                            if (q_sumsq && img_sumsq && img_sumsq[id] == q_sumsq[q] && key >= 1.0 - 4e-6) s = 0.0;
                        } else {
                            s = fmax(0.0, 2.0 - key);
                            if (q_sumsq && img_sumsq && img_sumsq[id] == q_sumsq[q] && key >= 2.0 - 8e-6) s = 0.0;
                        }
                    }
                    out_score[q * topk + r] = s;
                }
            }
            __syncthreads();
        }
    }
}

// merge candidate lists that already carry their final score (the per-rank top-k lists of a sharded
// database, all-gathered): same comparator, the winner's score is copied instead of recomputed
__global__ void __launch_bounds__(SC_THREADS)
topk_merge_scored_kernel(const double *__restrict__ in_key, const int32_t *__restrict__ in_id,
                         const double *__restrict__ in_score, int64_t n_query, int n_cand, int topk,
                         double *__restrict__ out_key, int32_t *__restrict__ out_id, double *__restrict__ out_score)
{
    extern __shared__ unsigned char s_raw[];
    double *s_keys = reinterpret_cast<double *>(s_raw);                      // [n_cand]
    int32_t *s_ids = reinterpret_cast<int32_t *>(s_raw + (size_t)n_cand * 8);
    __shared__ double r_key[SC_WARPS];
    __shared__ int32_t r_id[SC_WARPS];
    __shared__ int r_owner[SC_WARPS];
    for (int64_t q = blockIdx.x; q < n_query; q += gridDim.x) {
        for (int t = threadIdx.x; t < n_cand; t += SC_THREADS) {
            const int32_t id = in_id[q * n_cand + t];
            s_ids[t] = id;
            s_keys[t] = id >= 0 ? in_key[q * n_cand + t] : -2.0;
        }
        __syncthreads();
        for (int r = 0; r < topk; ++r) {
            double bkey = -2.0; int32_t bid = 0x7fffffff; int bslot = -1;
            for (int t = threadIdx.x; t < n_cand; t += SC_THREADS) {
                const double k2 = s_keys[t];
                const int32_t i2 = k2 > -2.0 ? s_ids[t] : 0x7fffffff;
                if (better(k2, i2, bkey, bid)) { bkey = k2; bid = i2; bslot = t; }
            }
            double key = bkey; int32_t id = bid; int owner;
            block_best(key, id, owner, r_key, r_id, r_owner);
            const bool ok = key > -2.0;
            if (threadIdx.x == owner) {
                if (ok) s_keys[bslot] = -2.0;
                out_key[q * topk + r] = key;
                out_id[q * topk + r] = ok ? id : -1;
                out_score[q * topk + r] = ok ? in_score[q * n_cand + bslot] : 1e6;
            }
            __syncthreads();
        }
    }
}

// ---- dense encoders (any object with embedding(image), e.g. the reference's AlexNet, encoders/alexnet.py:15-19) -------
// score(d, q) = || d/|d| - q/|q| ||_2 exactly as Database.score evaluates it (database.py:66-69), for unit rows that are
// already normalised: one warp per database row, 8 queries at a time from shared memory, binary64, fixed summation
// order (lane-strided partial sums, then an xor butterfly).  NaN (a zero embedding) -> 1e6 (database.py:70).
constexpr int DS_Q = 8, DS_TILE = 512;
__global__ void __launch_bounds__(256)
dense_score_kernel(const double *__restrict__ db, int64_t n, int64_t d, const double *__restrict__ q, int64_t nq,
                   double *__restrict__ out)
{
    __shared__ double s_q[DS_Q][DS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * DS_Q;
    const int nqh = (int)((nq - q0) < DS_Q ? (nq - q0) : DS_Q);
    for (int64_t row0 = (int64_t)blockIdx.x * 8; row0 < n; row0 += (int64_t)gridDim.x * 8) {
        const int64_t row = row0 + warp;
        double acc[DS_Q];
#pragma unroll
        for (int j = 0; j < DS_Q; ++j) acc[j] = 0.0;
        for (int64_t t0 = 0; t0 < d; t0 += DS_TILE) {
            const int tl = (int)((d - t0) < DS_TILE ? (d - t0) : DS_TILE);
            __syncthreads();
            for (int i = threadIdx.x; i < DS_Q * DS_TILE; i += 256) {
                const int j = i / DS_TILE, e = i - j * DS_TILE;
                s_q[j][e] = (j < nqh && e < tl) ? q[(q0 + j) * d + t0 + e] : 0.0;
            }
            __syncthreads();
            if (row < n) {
                for (int e = lane; e < tl; e += 32) {
                    const double x = db[row * d + t0 + e];
#pragma unroll
                    for (int j = 0; j < DS_Q; ++j) {
                        const double df = x - s_q[j][e];
                        acc[j] += df * df;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < DS_Q; ++j) {
            double s = acc[j];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(VT_FULL, s, o);
            if (lane == 0 && row < n && j < nqh) {
                const double r = sqrt(s);
                out[(q0 + j) * n + row] = (r == r) ? r : 1e6;
            }
        }
    }
}

// per (query, shard of R rows): the topk smallest scores, ties by smaller index (the reference's stable ascending sort,
// database.py:79-80), as (key = -score, id, score) lists for vt_topk_merge_scored
template <int R>
__global__ void __launch_bounds__(SC_THREADS)
topk_rows_kernel(const double *__restrict__ score, int64_t n, int64_t nq, int n_shards, int topk, double *__restrict__ out_key,
                 int32_t *__restrict__ out_id, double *__restrict__ out_score)
{
    constexpr int PER = R / SC_THREADS;
    extern __shared__ unsigned char s_raw[];
    double *s_keys = reinterpret_cast<double *>(s_raw);
    __shared__ double r_key[SC_WARPS];
    __shared__ int32_t r_id[SC_WARPS];
    __shared__ int r_owner[SC_WARPS];
    for (int64_t pair = blockIdx.x; pair < nq * n_shards; pair += gridDim.x) {
        const int64_t q = pair / n_shards;
        const int shard = (int)(pair - q * n_shards);
        const int64_t i0 = (int64_t)shard * R;
        double bkey = -1e300; int32_t bid = 0x7fffffff;
        for (int u = 0; u < PER; ++u) {
            const int t = u * SC_THREADS + threadIdx.x;
            const double key = i0 + t < n ? -score[q * n + i0 + t] : -1e300;     // padding never wins
            s_keys[t] = key;
            if (better(key, (int32_t)(i0 + t), bkey, bid)) { bkey = key; bid = (int32_t)(i0 + t); }
        }
        __syncthreads();
        for (int r = 0; r < topk; ++r) {
            double key = bkey; int32_t id = bid; int owner;
            block_best(key, id, owner, r_key, r_id, r_owner);
            const bool ok = key > -1e300;
            if (threadIdx.x == 0) {
                const size_t o = ((size_t)q * n_shards + shard) * topk + r;
                out_key[o] = key;
                out_id[o] = ok ? id : -1;
                out_score[o] = ok ? -key : 1e6;
            }
            if (threadIdx.x == owner) {
                if (ok) s_keys[id - i0] = -1e300;
                bkey = -1e300; bid = 0x7fffffff;
                for (int u = 0; u < PER; ++u) {
                    const int t = u * SC_THREADS + threadIdx.x;
                    const double k2 = s_keys[t];
                    if (better(k2, (int32_t)(i0 + t), bkey, bid)) { bkey = k2; bid = (int32_t)(i0 + t); }
                }
            }
        }
        __syncthreads();
    }
}

__global__ void rnorm_kernel(const unsigned long long *__restrict__ sumsq, int64_t n, double *__restrict__ rnorm)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long s = sumsq[t];
        rnorm[t] = s ? 1.0 / sqrt((double)s) : -1.0;
    }
}
}  // namespace

extern "C" int vt_score_shard_size(void) { return 8192; }
extern "C" int vt_score_max_topk(void) { return SC_MAXK; }

extern "C" int vt_rnorm(const uint64_t *d_sumsq, int64_t n, double *d_rnorm, void *stream)
{
    if (n <= 0) return VT_OK;
    const int64_t cap = (int64_t)vt_sm_count() * 8, want = (n + 255) / 256;
    rnorm_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>((const unsigned long long *)d_sumsq, n, d_rnorm);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

template <int MODE>
static int launch_score(const uint32_t *post_off, const void *postings, const double *img_rnorm, int64_t n_img,
                        int64_t n_ref, const int64_t *q_off, const int32_t *q_node, const float *q_val, int64_t n_query,
                        int topk, double *out_key, int32_t *out_id, double *all_key, const uint8_t *shard_empty,
                        cudaStream_t s)
{
    constexpr int R = 8192;
    const int n_shards = (int)((n_img + R - 1) / R);
    auto kern = score_shard_kernel<MODE, R>;
    const size_t smem = (size_t)R * 12;
    VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t pairs = n_query * n_shards;
    const int64_t cap = (int64_t)vt_sm_count() * 2 * 8;
    kern<<<(int)(pairs < cap ? pairs : cap), SC_THREADS, smem, s>>>(post_off, (const uint32_t *)postings, img_rnorm, n_img, n_ref,
                                                                    n_shards, q_off, q_node, q_val, n_query, topk, out_key,
                                                                    out_id, all_key, shard_empty, nullptr, 0);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

static int launch_score_tf(const uint32_t *post_off, const void *postings, const double *img_rnorm, int64_t n_img,
                           int64_t n_ref, const int64_t *q_off, const int32_t *q_node, const float *q_val, int64_t n_query,
                           int topk, double *out_key, int32_t *out_id, const double *shard_rho, int long_lists,
                           cudaStream_t s)
{
    constexpr int R = 8192;
    const int n_shards = (int)((n_img + R - 1) / R);
    const size_t smem = (size_t)R * 4 + SC_CAND * 12;
    const int64_t pairs = n_query * n_shards;
    const int64_t cap = (int64_t)vt_sm_count() * 6 * 4;
    const int blocks = (int)(pairs < cap ? pairs : cap);
    if (long_lists) {
        auto kern = score_shard_tf_kernel<R, 6>;
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, SC_THREADS, smem, s>>>(post_off, (const uint32_t *)postings, img_rnorm, n_img, n_ref, n_shards, q_off,
                                              q_node, q_val, n_query, topk, out_key, out_id, shard_rho, nullptr, 0, nullptr);
    } else {
        auto kern = score_shard_tf_kernel<R, 4>;
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, SC_THREADS, smem, s>>>(post_off, (const uint32_t *)postings, img_rnorm, n_img, n_ref, n_shards, q_off,
                                              q_node, q_val, n_query, topk, out_key, out_id, shard_rho, nullptr, 0, nullptr);
    }
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_score_topk(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm, int64_t n_img,
                             int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node, const float *d_q_val,
                             int64_t n_query, int mode, int topk, double *d_shard_key, int32_t *d_shard_id,
                             double *d_all_key, const uint8_t *d_shard_empty, const double *d_shard_rho, int long_lists,
                             void *stream)
{
    VT_CHECK_ARG(n_img > 0 && n_ref > 0 && n_query >= 0, "bad sizes");
    VT_CHECK_ARG(topk >= 1 && topk <= SC_MAXK, "topk out of range");
    VT_CHECK_ARG(mode != VT_SCORE_TF_L2 || d_img_rnorm != nullptr, "tf mode needs image norms");
    if (n_query == 0) return VT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == VT_SCORE_TF_L2 && d_all_key == nullptr && d_shard_rho != nullptr)
        return launch_score_tf(d_post_off, d_postings, d_img_rnorm, n_img, n_ref, d_q_off, d_q_node, d_q_val, n_query, topk,
                               d_shard_key, d_shard_id, d_shard_rho, long_lists, s);
    switch (mode) {
    case VT_SCORE_TF_L2: return launch_score<VT_SCORE_TF_L2>(d_post_off, d_postings, d_img_rnorm, n_img, n_ref, d_q_off, d_q_node, d_q_val, n_query, topk, d_shard_key, d_shard_id, d_all_key, d_shard_empty, s);
    case VT_SCORE_W_L2:
    case VT_SCORE_W_L1:
        vt_set_error("vt_score_topk: the weighted modes run over the tf index in binary64: vt_score_topk_weighted");
        return VT_EINVAL;
    }
    vt_set_error("vt_score_topk: unknown mode %d", mode);
    return VT_EINVAL;
}

extern "C" int vt_topk_merge(const double *d_in_key, const int32_t *d_in_id, int64_t n_query, int n_cand, int topk,
                             int mode, const double *d_q_rnorm, const uint64_t *d_q_sumsq, const uint64_t *d_img_sumsq,
                             double *d_out_key, int32_t *d_out_id, double *d_out_score, void *stream)
{
    VT_CHECK_ARG(n_query >= 0 && n_cand >= 1 && topk >= 1 && topk <= SC_MAXK, "bad sizes");
    VT_CHECK_ARG((size_t)n_cand * 12 <= 200 * 1024, "too many candidates per query for one merge");
    VT_CHECK_ARG(mode != VT_SCORE_TF_L2 || d_out_score == nullptr || d_q_rnorm != nullptr, "tf scores need query norms");
    if (n_query == 0) return VT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)n_cand * 12;
    const int64_t cap = (int64_t)vt_sm_count() * 8;
    const int blocks = (int)(n_query < cap ? n_query : cap);
#define VT_M(MODEV)                                                                                              \
    {                                                                                                            \
        auto kern = topk_merge_kernel<MODEV>;                                                                    \
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        kern<<<blocks, SC_THREADS, smem, s>>>(d_in_key, d_in_id, n_query, n_cand, topk, d_q_rnorm,               \
                                              (const unsigned long long *)d_q_sumsq,                             \
                                              (const unsigned long long *)d_img_sumsq, d_out_key, d_out_id,      \
                                              d_out_score);                                                      \
        VT_LAUNCH_CHECK();                                                                                       \
        return VT_OK;                                                                                            \
    }
    switch (mode) {
    case VT_SCORE_TF_L2: VT_M(VT_SCORE_TF_L2)
    case VT_SCORE_W_L2: VT_M(VT_SCORE_W_L2)
    case VT_SCORE_W_L1: VT_M(VT_SCORE_W_L1)
    }
#undef VT_M
    vt_set_error("vt_topk_merge: unknown mode %d", mode);
    return VT_EINVAL;
}

// vt_topk_merge for candidates that carry their score (d_in_score [n_query, n_cand]): the multi-GPU merge of
// the all-gathered per-rank lists (ids global, -1 = empty slot).
extern "C" int vt_topk_merge_scored(const double *d_in_key, const int32_t *d_in_id, const double *d_in_score,
                                    int64_t n_query, int n_cand, int topk, double *d_out_key, int32_t *d_out_id,
                                    double *d_out_score, void *stream)
{
    VT_CHECK_ARG(n_query >= 0 && n_cand >= 1 && topk >= 1 && topk <= SC_MAXK, "bad sizes");
    VT_CHECK_ARG((size_t)n_cand * 12 <= 200 * 1024, "too many candidates per query for one merge");
    VT_CHECK_ARG(d_in_key && d_in_id && d_in_score && d_out_key && d_out_id && d_out_score, "null buffer");
    if (n_query == 0) return VT_OK;
    const size_t smem = (size_t)n_cand * 12;
    VT_CUDA(cudaFuncSetAttribute(topk_merge_scored_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)vt_sm_count() * 8;
    topk_merge_scored_kernel<<<(int)(n_query < cap ? n_query : cap), SC_THREADS, smem, (cudaStream_t)stream>>>(
        d_in_key, d_in_id, d_in_score, n_query, n_cand, topk, d_out_key, d_out_id, d_out_score);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// vt_score_topk (tf mode, exhaustive selection) whose shard accumulators start from d_seed[q * seed_stride + img]
// instead of zero: the int32 dots over the dense top levels of an all-nodes index (vt_dense_dots, csrc/vt_dense.cu);
// the postings then only hold the deep levels.
extern "C" int vt_score_topk_seeded(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                                    int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                                    const float *d_q_val, int64_t n_query, int topk, const int32_t *d_seed,
                                    int64_t seed_stride, double *d_shard_key, int32_t *d_shard_id, double *d_all_key,
                                    void *stream)
{
    VT_CHECK_ARG(n_img > 0 && n_ref > 0 && n_query >= 0 && d_img_rnorm && d_seed && seed_stride >= n_img, "bad arguments");
    VT_CHECK_ARG(topk >= 1 && topk <= SC_MAXK, "topk out of range");
    if (n_query == 0) return VT_OK;
    constexpr int R = 8192;
    const int n_shards = (int)((n_img + R - 1) / R);
    // the tf kernel in its long-list configuration (whole warps walk the runs, 6 CTAs per SM): 4 B of shared memory per
    // image, keys recomputed on demand by the exhaustive selection
    auto kern = score_shard_tf_kernel<R, 6>;
    const size_t smem = (size_t)R * 4 + SC_CAND * 12;
    VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t pairs = n_query * n_shards, cap = (int64_t)vt_sm_count() * 6 * 4;
    kern<<<(int)(pairs < cap ? pairs : cap), SC_THREADS, smem, (cudaStream_t)stream>>>(
        d_post_off, (const uint32_t *)d_postings, d_img_rnorm, n_img, n_ref, n_shards, d_q_off, d_q_node, d_q_val, n_query, topk,
        d_shard_key, d_shard_id, nullptr, d_seed, seed_stride, d_all_key);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// Dense-encoder path: scores of unit embedding rows, binary64 (d_out [n_query, n]).
extern "C" int vt_dense_scores(const double *d_db_unit, int64_t n, int64_t dim, const double *d_q_unit, int64_t n_query,
                               double *d_out, void *stream)
{
    VT_CHECK_ARG(d_db_unit && d_q_unit && d_out && n >= 0 && dim >= 1 && n_query >= 0, "bad arguments");
    if (n == 0 || n_query == 0) return VT_OK;
    const int64_t want = (n + 7) / 8, cap = (int64_t)vt_sm_count() * 8;
    dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)((n_query + DS_Q - 1) / DS_Q));
    dense_score_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_db_unit, n, dim, d_q_unit, n_query, d_out);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// top-k smallest scores per (query, shard of vt_score_shard_size() rows): lists [n_query, n_shards, topk] for
// vt_topk_merge_scored (key = -score, ties by smaller index)
extern "C" int vt_topk_rows(const double *d_score, int64_t n, int64_t n_query, int topk, double *d_shard_key,
                            int32_t *d_shard_id, double *d_shard_score, void *stream)
{
    VT_CHECK_ARG(d_score && d_shard_key && d_shard_id && d_shard_score && n > 0 && n_query >= 0, "bad arguments");
    VT_CHECK_ARG(topk >= 1 && topk <= SC_MAXK, "topk out of range");
    if (n_query == 0) return VT_OK;
    constexpr int R = 8192;
    const int n_shards = (int)((n + R - 1) / R);
    auto kern = topk_rows_kernel<R>;
    const size_t smem = (size_t)R * 8;
    VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t pairs = n_query * n_shards, cap = (int64_t)vt_sm_count() * 8;
    kern<<<(int)(pairs < cap ? pairs : cap), SC_THREADS, smem, (cudaStream_t)stream>>>(d_score, n, n_query, n_shards, topk,
                                                                                       d_shard_key, d_shard_id, d_shard_score);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
