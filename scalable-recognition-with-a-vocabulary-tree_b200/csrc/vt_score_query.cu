// vt_score_query.cu -- tf / L2 inverted-file scoring (database.py:55-87) with the top-k CARRIED from shard to shard.
//
// vt_score.cu gives every (query, shard) pair its own CTA and its own top-k: a histogram pass (or, for all-nodes indexes
// whose dots are ~1e6, ten block-wide arg-max rounds) per pair, one list per pair in HBM.  ncu (profiles/r02_*scorer*):
// that selection is half of the instructions of a pair in leaf-words mode and a third in all-nodes mode, the lane-per-word
// walk issues ~1.35 warp instructions per posting with 32 separate sectors per load instruction, and the warp-per-run walk
// of long lists has ONE load in flight per warp.  Here a CTA takes a query through a GROUP of consecutive shards:
//   * the k best (key, id) pairs seen so far stay in shared memory; their worst key T is an exact lower bound of the final
//     k-th key, so in every further shard only images that can beat T are looked at: `dot >= ceil(T / max 1/|d|)` (integer
//     test, leaf-words indexes) or `dot / |d| >= T` (all-nodes indexes: every image shares the top levels with the
//     query, the integer bound says nothing) -- typically 0-3 candidates, inserted into the list by one warp.  Only the
//     first shard of a group (no T yet), or a shard where nothing can be pruned (T <= 0), runs the exhaustive selection;
//   * the row pointers of a query word are fetched ONCE per 8 shards (the 9 pointers sit in one or two sectors) and kept
//     on chip as run lengths -- not one 32-byte sector per (word, shard) -- and the word's node is not looked up again;
//   * the walk is FLAT: the posting runs of 32 query words are concatenated (prefix sum of their lengths); lane l of the
//     warp takes position 32 c + l of chunk c, finds its word from a head-bit mask with one popc, and 4 chunks are loaded
//     before the first is accumulated -- coalesced requests with every lane busy, whatever the run length; runs longer
//     than 64 postings are walked by the whole warp with 4 x 32 postings in flight.  (Measured alternative: 2^n lanes per
//     run, 32 / 2^n runs per round, 8 rounds in flight, no shared-memory bookkeeping -- 7 % slower at the best n.)
// Shared-memory atomics are not the limit: tools/atoms_probe.cu measures 12 random 32-bit atomicAdd lanes per clock and
// SM (3.4 T/s), the scorer uses ~0.8.  The kernel is latency bound (8 warps per scheduler, dependent LDS -> LDG -> ATOMS
// chains; ncu: issue slots 35 % busy, long scoreboard 34 %, short scoreboard 19 %, barrier 18 % of the stall samples).
// Results are the same (key, id) sets as vt_score.cu's kernels: dots are exact integers, keys binary64, order =
// (key descending, id ascending) like the reference's stable ascending sort of the scores (database.py:79-80).
#include "vt_common.cuh"
#include <limits.h>

namespace {
constexpr int SQ_THREADS = 256;
constexpr int SQ_WARPS = SQ_THREADS / 32;
constexpr int SQ_MAXK = 64;
constexpr int SQ_R = 8192;                 // images per shard (vt_score_shard_size)
constexpr int SQ_PER = SQ_R / SQ_THREADS;
constexpr int SQ_CAND = 128;               // candidates per shard before the exhaustive selection takes over
constexpr int SQ_SUB = 8;                  // shards whose run lengths are fetched together (one pointer sector per word)
constexpr int SQ_WCACHE = 1024;            // query words whose runs are tracked on chip (the rest re-read their pointers)
constexpr uint32_t SQ_LONG = 64;           // longer runs leave the flat walk
constexpr int SQ_HEADS = 64;               // 32 words x 64 postings = 2048 flat positions
constexpr int SQ_U = 4;                    // chunks (posting loads per lane) in flight
constexpr int SQ_G = 3;                    // long runs walked together
constexpr uint32_t SQ_LOCAL_MASK = 8191u;  // posting = (tf << 13) | local image index (vt_posting_pack)
constexpr int SQ_LOCAL_BITS = 13;
constexpr int SQ_TAKEN = INT_MIN;

template <class W>
struct SqScratchT {                        // per warp
    uint32_t base[32];                     // compacted non-empty words: posting index of flat position 0 of the word
    W tf[32];                              //   ... and the query's count (tf modes) or binary64 weight (tf-idf) of the word
    uint32_t head[SQ_HEADS];               // bit j of head[c]: flat position 32 c + j starts a word's run
};
using SqScratch = SqScratchT<int>;

template <class W> struct SqWordStore { typedef uint16_t type; };          // cached per-word query factor
template <> struct SqWordStore<double> { typedef double type; };

template <class A, class W>
struct SqSharedT {
    A acc[SQ_R];
    double ckey[SQ_CAND];
    int cslot[SQ_CAND];
    double lkey[SQ_MAXK], nkey[SQ_MAXK];   // the carried list (best first) and the one the exhaustive selection builds
    int32_t lid[SQ_MAXK], nid[SQ_MAXK];
    SqScratchT<W> ws[SQ_WARPS];
    // thread-private slots (word i of the query belongs to thread i % 256): where the word's run starts in the current
    // shard, its query count, and the lengths of its runs in the SQ_SUB shards of the current sub-group (255 = longer)
    uint32_t w_run[SQ_WCACHE];
    typename SqWordStore<W>::type w_tf[SQ_WCACHE];
    uint8_t w_len[SQ_SUB][SQ_WCACHE];
    double r_key[SQ_WARPS];
    int32_t r_id[SQ_WARPS];
    int r_owner[SQ_WARPS];
    int hist[64];                          // dots of a group's first step (threshold bootstrap)
    int ncand, ln, boot;
};
using SqShared = SqSharedT<int, int>;

__device__ __forceinline__ bool sq_better(double ka, int32_t ia, double kb, int32_t ib)
{
    return ka > kb || (ka == kb && ia < ib);
}

template <class Shared>
__device__ __forceinline__ void sq_block_best(double &key, int32_t &id, int &owner, Shared &sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    owner = threadIdx.x;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const double k2 = __shfl_xor_sync(VT_FULL, key, o);
        const int32_t i2 = __shfl_xor_sync(VT_FULL, id, o);
        const int o2 = __shfl_xor_sync(VT_FULL, owner, o);
        if (sq_better(k2, i2, key, id)) { key = k2; id = i2; owner = o2; }
    }
    if (lane == 0) { sh.r_key[warp] = key; sh.r_id[warp] = id; sh.r_owner[warp] = owner; }
    __syncthreads();
    key = sh.r_key[0]; id = sh.r_id[0]; owner = sh.r_owner[0];
#pragma unroll
    for (int w = 1; w < SQ_WARPS; ++w)
        if (sq_better(sh.r_key[w], sh.r_id[w], key, id)) { key = sh.r_key[w]; id = sh.r_id[w]; owner = sh.r_owner[w]; }
    __syncthreads();
}

// Warp 0: insert the nc candidates (sh.cslot / sh.ckey, images img0 + slot) into the carried list (best first, sh.ln entries,
// at most topk), one at a time; the list stays sorted and (key, id) is a strict total order, so the result does not depend on
// the order in which the candidates were found.
template <class Shared>
__device__ __forceinline__ void sq_insert_candidates(Shared &sh, int nc, int64_t img0, int topk, int lane)
{
    int ln = sh.ln;                                            // (warp uniform)
    for (int i0 = 0; i0 < nc; i0 += 32) {
        const int i = i0 + lane;
        const double ck = i < nc ? sh.ckey[i] : -2.0;
        const int32_t cid = i < nc ? (int32_t)(img0 + sh.cslot[i]) : 0x7fffffff;
        unsigned int live = __ballot_sync(VT_FULL, i < nc);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            const double k = __shfl_sync(VT_FULL, ck, src);
            const int32_t id = __shfl_sync(VT_FULL, cid, src);
            const bool h0 = lane < ln, h1 = lane + 32 < ln;
            const double k0 = h0 ? sh.lkey[lane] : 0.0, k1 = h1 ? sh.lkey[lane + 32] : 0.0;
            const int32_t d0 = h0 ? sh.lid[lane] : 0, d1 = h1 ? sh.lid[lane + 32] : 0;
            const bool b0 = h0 && sq_better(k0, d0, k, id), b1 = h1 && sq_better(k1, d1, k, id);
            const int pos = __popc(__ballot_sync(VT_FULL, b0)) + __popc(__ballot_sync(VT_FULL, b1));
            if (pos < topk) {                                  // entries from pos on move down by one, a full list drops its last
                __syncwarp();
                if (h0 && !b0 && lane + 1 < topk) { sh.lkey[lane + 1] = k0; sh.lid[lane + 1] = d0; }
                if (h1 && !b1 && lane + 33 < topk) { sh.lkey[lane + 33] = k1; sh.lid[lane + 33] = d1; }
                __syncwarp();
                if (lane == 0) { sh.lkey[pos] = k; sh.lid[pos] = id; }
                __syncwarp();
                if (ln < topk) ++ln;
            }
        }
    }
    if (lane == 0) sh.ln = ln;
}

// Warp 0, after sh.hist holds the histogram of the step's dots >= 1 (clamped at 63): the largest b >= 1 such that at least
// topk images of the step have a dot >= b (0 if there is none).  Those images have keys >= b * (smallest 1/|d| of the step),
// a valid lower bound of the final k-th key: a group's first step can prune with it instead of selecting exhaustively.
template <class Shared>
__device__ __forceinline__ void sq_bootstrap_bin(Shared &sh, int topk, int lane)
{
    int hi = sh.hist[63 - lane], lo = lane < 31 ? sh.hist[31 - lane] : 0;     // (bin 0 is not a bound)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t1 = __shfl_up_sync(VT_FULL, hi, o), t2 = __shfl_up_sync(VT_FULL, lo, o);
        if (lane >= o) { hi += t1; lo += t2; }
    }
    lo += __shfl_sync(VT_FULL, hi, 31);
    const unsigned int mh = __ballot_sync(VT_FULL, hi >= topk), ml = __ballot_sync(VT_FULL, lo >= topk);
    if (lane == 0) sh.boot = mh ? 63 - (__ffs(mh) - 1) : (ml ? 31 - (__ffs(ml) - 1) : 0);
}

// acc[local image of the posting] += (tf of the posting) * (tf of the query word), predicated: one `red.shared` under a
// predicate instead of a branch around an atomic (the compiler wraps every `if (ok) atomicAdd` in BSSY / BRA / BSYNC --
// a third of the instructions of the walk).  acc_s = shared-window address of the accumulators.
__device__ __forceinline__ void sq_add(uint32_t acc_s, uint32_t pe, int tf, bool ok)
{
    const uint32_t addr = (pe & SQ_LOCAL_MASK) * 4u + acc_s;
    const int val = tf * (int)(pe >> SQ_LOCAL_BITS);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 :: "r"(addr), "r"(val), "r"((uint32_t)ok) : "memory");
}

// tf-idf: acc[local image] += (binary64 query weight of the word) * (tf of the posting)
__device__ __forceinline__ void sq_add(double *acc, uint32_t pe, double w, bool ok)
{
    if (ok) atomicAdd(acc + (pe & SQ_LOCAL_MASK), w * (double)(pe >> SQ_LOCAL_BITS));
}

// SEEDED: accumulators start from the int32 dots over the dense top levels (vt_dense_dots) and candidates are filtered by
// their exact key; otherwise they start from zero and the filter is the integer bound through shard_rmax.
// MODE: SQ_TF (integer accumulators from zero, integer pruning bound), SQ_TF_SEEDED (integer accumulators from the dense-top
// dots, exact-key filter), SQ_TFIDF_L2 (binary64 accumulators: q_val_ are the binary64 query weights of vt_query_weights,
// img_rnorm the weighted norms, q_rnorm <= 0 marks a query without weight; exact-key filter).
enum { SQ_TF = 0, SQ_TF_SEEDED = 1, SQ_TFIDF_L2 = 2 };
template <int MODE> struct SqTypes { typedef int A; typedef int W; typedef float Q; };
template <> struct SqTypes<SQ_TFIDF_L2> { typedef double A; typedef double W; typedef double Q; };

template <int MODE, int MINB, int U, int G>
__global__ void __launch_bounds__(SQ_THREADS, MINB)
score_query_kernel(const uint32_t *__restrict__ post_off, const uint32_t *__restrict__ postings,
                   const double *__restrict__ img_rnorm, int64_t n_img, int n_shards, const int64_t *__restrict__ q_off,
                   const int32_t *__restrict__ q_node, const void *__restrict__ q_val_, const double *__restrict__ q_rnorm,
                   int64_t n_query, int topk, const double *__restrict__ shard_rmax, const double *__restrict__ shard_rmin,
                   const int32_t *__restrict__ seed, int64_t seed_stride, int n_groups, int per_group,
                   double *__restrict__ out_key, int32_t *__restrict__ out_id)
{
    typedef typename SqTypes<MODE>::A A;                   // accumulator
    typedef typename SqTypes<MODE>::W W;                   // per-word query factor
    constexpr bool SEEDED = MODE == SQ_TF_SEEDED, WEIGHTED = MODE == SQ_TFIDF_L2, EXACT_FILTER = MODE != SQ_TF;
    const typename SqTypes<MODE>::Q *__restrict__ q_val = static_cast<const typename SqTypes<MODE>::Q *>(q_val_);
    extern __shared__ __align__(16) unsigned char s_raw[];
    SqSharedT<A, W> &sh = *reinterpret_cast<SqSharedT<A, W> *>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t le_mask = (2u << lane) - 1u, lt_mask = (1u << lane) - 1u;
    SqScratchT<W> &ws = sh.ws[warp];
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(sh.acc);
    auto add = [&](uint32_t pe, W w, bool ok) {
        if constexpr (WEIGHTED) sq_add(sh.acc, pe, w, ok);
        else sq_add(acc_s, pe, w, ok);
    };
    const A TAKEN = WEIGHTED ? (A)-1 : (A)SQ_TAKEN;         // (dots and weighted sums are never negative)

    for (int64_t item = blockIdx.x; item < n_query * n_groups; item += gridDim.x) {
        const int64_t q = item / n_groups;
        const int g = (int)(item - q * n_groups);
        const int s0 = g * per_group, s1 = s0 + per_group < n_shards ? s0 + per_group : n_shards;
        const int64_t qb = q_off[q], qe = q_off[q + 1];
        const bool q_ok = !WEIGHTED || __ldg(q_rnorm + q) > 0.0;

        auto fill = [&](int shard) {
            if constexpr (WEIGHTED) {
                for (int t = tid; t < SQ_R; t += SQ_THREADS) sh.acc[t] = 0.0;
            } else if constexpr (SEEDED) {
                const int64_t img0 = (int64_t)shard * SQ_R;
                const int n_here = (int)((n_img - img0) < SQ_R ? (n_img - img0) : SQ_R);
                const int32_t *row = seed + (size_t)q * seed_stride + img0;
                for (int t = tid; t < SQ_R; t += SQ_THREADS) sh.acc[t] = t < n_here ? __ldg(row + t) : 0;
            } else {
                int4 *a4 = reinterpret_cast<int4 *>(sh.acc);
                for (int t = tid; t < SQ_R / 4; t += SQ_THREADS) a4[t] = make_int4(0, 0, 0, 0);
            }
        };
        if (s0 < s1) fill(s0);
        if (tid == 0) { sh.ln = 0; sh.ncand = 0; }
        __syncthreads();

        for (int shard = s0; shard < s1; ++shard) {
            const int64_t img0 = (int64_t)shard * SQ_R;
            const int n_here = (int)((n_img - img0) < SQ_R ? (n_img - img0) : SQ_R);
            const uint32_t *off = post_off + shard;                // row of (node, shard) = node * n_shards + shard

            // ---- walk ------------------------------------------------------------------------------------------------
            // one call = the runs of 32 query words (lane l: run [b, b + len) of a word with query count tf)
            auto walk32 = [&](uint32_t b, uint32_t len, W tf) {
                const bool is_long = len > SQ_LONG;
                unsigned int m = __ballot_sync(VT_FULL, is_long);
                while (m) {
                    // long runs: the whole warp walks G runs together, the first U x 32 postings of each are loaded
                    // before any is accumulated (one DRAM round trip per G runs instead of one per run)
                    uint32_t lb[G], lend[G];
                    W ltf[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int src = m ? __ffs(m) - 1 : 0;
                        const bool have = m != 0;
                        m &= m - 1;
                        lb[g] = __shfl_sync(VT_FULL, b, src);
                        lend[g] = have ? lb[g] + __shfl_sync(VT_FULL, len, src) : lb[g];
                        ltf[g] = __shfl_sync(VT_FULL, tf, src);
                    }
                    uint32_t pe[G][U];
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const uint32_t p = lb[g] + lane + 32 * u;
                            pe[g][u] = p < lend[g] ? __ldg(postings + p) : 0u;
                        }
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            add(pe[g][u], ltf[g], lb[g] + lane + 32 * u < lend[g]);
#pragma unroll
                    for (int g = 0; g < G; ++g)                 // what is left of runs beyond U x 32 postings
                        for (uint32_t p = lb[g] + 32 * U + lane; p < lend[g]; p += 32 * U) {
                            uint32_t pt[U];
#pragma unroll
                            for (int u = 0; u < U; ++u) pt[u] = p + 32 * u < lend[g] ? __ldg(postings + p + 32 * u) : 0u;
#pragma unroll
                            for (int u = 0; u < U; ++u)
                                add(pt[u], ltf[g], p + 32 * u < lend[g]);
                        }
                }
                if (is_long) len = 0;
                // flat walk over the concatenated short runs of the 32 words
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const uint32_t total = __shfl_sync(VT_FULL, incl, 31);
                if (total == 0) return;                            // (warp uniform)
                const uint32_t excl = incl - len;
                ws.head[lane] = 0;
                ws.head[lane + 32] = 0;
                __syncwarp();
                const bool ne = len > 0;
                const unsigned int nem = __ballot_sync(VT_FULL, ne);
                if (ne) {
                    const int rank = __popc(nem & lt_mask);
                    ws.base[rank] = b - excl;
                    ws.tf[rank] = tf;
                    atomicOr(&ws.head[excl >> 5], 1u << (excl & 31));
                }
                __syncwarp();
                int before = 0;                                    // words that start before the current chunk
                for (uint32_t c0 = 0; c0 * 32 < total; c0 += U) {
                    uint32_t pe[U];
                    W tv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t c = c0 + u, j = c * 32 + lane;
                        const uint32_t h = c * 32 < total ? ws.head[c] : 0u;
                        const int idx = before + __popc(h & le_mask) - 1;
                        before += __popc(h);
                        pe[u] = 0u;
                        tv[u] = 0;                                 // (tf > 0 for every real word: 0 marks an idle slot)
                        if (j < total) {
                            tv[u] = ws.tf[idx];
                            pe[u] = __ldg(postings + (ws.base[idx] + j));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        add(pe[u], tv[u], tv[u] != 0);
                }
                __syncwarp();                                      // the scratch is rewritten by the next 32 words
            };
            const int sub = (shard - s0) % SQ_SUB;
            if (sub == 0) {
                // run lengths of the next SQ_SUB shards: the 9 row pointers of a word sit in one or two sectors and are read
                // ONCE here -- not one sector per (word, shard) -- and the word's node is not looked up again
                const int last = n_shards - shard;                 // pointer index of the row's end
                for (int i = tid; i < SQ_WCACHE; i += SQ_THREADS) {
                    const int64_t w = qb + i;
                    uint32_t p[SQ_SUB + 1];
                    W tf = 0;
                    if (w < qe) {
                        const uint32_t *r = off + (size_t)__ldg(q_node + w) * n_shards;
                        tf = (W)__ldg(q_val + w);
#pragma unroll
                        for (int j = 0; j <= SQ_SUB; ++j) p[j] = __ldg(r + (j < last ? j : last));
                    } else {
#pragma unroll
                        for (int j = 0; j <= SQ_SUB; ++j) p[j] = 0u;
                    }
                    sh.w_run[i] = p[0];
                    const bool wide = !WEIGHTED && tf >= (W)65535;      // (a count that does not fit the 16-bit slot)
                    sh.w_tf[i] = (typename SqWordStore<W>::type)(wide ? (W)65535 : tf);
#pragma unroll
                    for (int j = 0; j < SQ_SUB; ++j) {
                        const uint32_t len = p[j + 1] - p[j];
                        sh.w_len[j][i] = (uint8_t)(len < 255u && !wide ? len : 255u);
                    }
                }
                // (thread-private slots: no barrier)
            }
#pragma unroll 1
            for (int i = tid; i < SQ_WCACHE && qb + (i - lane) < qe; i += SQ_THREADS) {
                uint32_t len = sh.w_len[sub][i], b = sh.w_run[i];
                W tf = (W)sh.w_tf[i];
                if (len == 255u) {                                 // a long run (or a huge query count): read it again
                    const int64_t w = qb + i;
                    const uint32_t *r = off + (size_t)__ldg(q_node + w) * n_shards;
                    b = __ldg(r);
                    len = __ldg(r + 1) - b;
                    tf = (W)__ldg(q_val + w);
                }
                sh.w_run[i] = b + len;
                walk32(b, len, tf);
            }
            for (int64_t w0 = qb + SQ_WCACHE + (int64_t)warp * 32; w0 < qe; w0 += SQ_THREADS) {
                const int64_t w = w0 + lane;
                uint32_t b = 0, len = 0;
                W tf = 0;
                if (w < qe) {
                    const uint32_t *r = off + (size_t)__ldg(q_node + w) * n_shards;
                    b = __ldg(r);
                    len = __ldg(r + 1) - b;
                    tf = (W)__ldg(q_val + w);
                }
                walk32(b, len, tf);
            }
            __syncthreads();

            // ---- selection against the carried list ------------------------------------------------------------------
            const int ln = sh.ln;
            double T = ln == topk ? sh.lkey[topk - 1] : -1.0;
            if constexpr (MODE == SQ_TF) {
                // first step of a group (empty list): bound the k-th key from the histogram of the step's dots
                const double rmin = (ln == 0 && shard_rmin) ? __ldg(shard_rmin + shard) : 0.0;
                if (rmin > 0.0) {                                  // (block uniform)
                    if (tid < 64) sh.hist[tid] = 0;
                    __syncthreads();
                    const int4 *h4 = reinterpret_cast<const int4 *>(sh.acc);
                    for (int u = 0; u < SQ_PER / 4; ++u) {
                        const int4 a = h4[u * SQ_THREADS + tid];
                        const int av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (av[i] >= 1) atomicAdd(&sh.hist[av[i] < 63 ? av[i] : 63], 1);
                    }
                    __syncthreads();
                    if (warp == 0) sq_bootstrap_bin(sh, topk, lane);
                    __syncthreads();
                    if (sh.boot >= 1) T = (double)sh.boot * rmin;
                }
            }
            bool exhaustive = !(T > 0.0);                          // nothing can be pruned: untouched images reach T
            if (!exhaustive) {
                if constexpr (EXACT_FILTER) {
#pragma unroll 4
                    for (int u = 0; u < SQ_PER; ++u) {
                        const int t = u * SQ_THREADS + tid;
                        const A a = sh.acc[t];
                        if (a > (A)0 && t < n_here) {
                            const double key = (double)a * __ldg(img_rnorm + img0 + t);   // (empty images: a == 0)
                            if (key >= T) {
                                const int pos = atomicAdd(&sh.ncand, 1);
                                if (pos < SQ_CAND) { sh.cslot[pos] = t; sh.ckey[pos] = key; }
                            }
                        }
                    }
                } else {
                    // smallest integer dot whose key can reach T: key = dot / |d| <= dot * rmax
                    const double rmax = __ldg(shard_rmax + shard);
                    int d_lo = INT_MAX;
                    if (rmax > 0.0) {
                        const double x = ceil(T / rmax);
                        d_lo = x >= 2147483647.0 ? INT_MAX : (int)x;
                        while (d_lo > 1 && (double)(d_lo - 1) * rmax >= T) --d_lo;
                        while (d_lo < INT_MAX && (double)d_lo * rmax < T) ++d_lo;
                        if (d_lo < 1) d_lo = 1;
                    }
                    const int4 *a4 = reinterpret_cast<const int4 *>(sh.acc);
#pragma unroll 2
                    for (int u = 0; u < SQ_PER / 4; ++u) {
                        const int t4 = u * SQ_THREADS + tid;
                        const int4 a = a4[t4];
                        const int mx = max(max(a.x, a.y), max(a.z, a.w));
                        if (mx >= d_lo) {
                            const int av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (av[i] >= d_lo) {
                                    const int pos = atomicAdd(&sh.ncand, 1);
                                    if (pos < SQ_CAND) sh.cslot[pos] = t4 * 4 + i;
                                }
                        }
                    }
                }
                __syncthreads();
                const int nc = sh.ncand;
                if (nc > SQ_CAND) {
                    exhaustive = true;                             // (block uniform)
                } else if (nc > 0) {
                    if constexpr (!EXACT_FILTER) {
                        for (int i = tid; i < nc; i += SQ_THREADS) {
                            const int t = sh.cslot[i];
                            sh.ckey[i] = (double)sh.acc[t] * __ldg(img_rnorm + img0 + t);
                        }
                        __syncthreads();
                    }
                    if (warp == 0) sq_insert_candidates(sh, nc, img0, topk, lane);
                }
            }
            if (exhaustive) {
                // ---- k rounds of block arg-max over (this shard's images) U (the carried list) -------------------------
                auto key_of = [&](int t) -> double {
                    if (t >= n_here) return -2.0;
                    const A a = sh.acc[t];
                    if (a == TAKEN) return -2.0;
                    const double rn = __ldg(img_rnorm + img0 + t);
                    return (rn < 0.0 || !q_ok) ? -1.0 : (double)a * rn;
                };
                double lk = tid < ln ? sh.lkey[tid] : -2.0;        // thread t owns list entry t
                const int32_t li = tid < ln ? sh.lid[tid] : 0x7fffffff;
                double bkey = -2.0; int32_t bid = 0x7fffffff;
                for (int u = 0; u < SQ_PER; ++u) {
                    const int t = u * SQ_THREADS + tid;
                    const double key = key_of(t);
                    if (sq_better(key, (int32_t)(img0 + t), bkey, bid)) { bkey = key; bid = (int32_t)(img0 + t); }
                }
                int n_new = 0;
                for (int r = 0; r < topk; ++r) {
                    double key = bkey; int32_t id = bid; int owner;
                    if (sq_better(lk, li, key, id)) { key = lk; id = li; }
                    sq_block_best(key, id, owner, sh);
                    if (key > -2.0) n_new = r + 1;
                    if (tid == 0) { sh.nkey[r] = key; sh.nid[r] = key > -2.0 ? id : -1; }
                    if (tid == owner && key > -2.0) {
                        if ((int64_t)id >= img0) {                 // one of this shard's images (list ids are smaller)
                            sh.acc[id - img0] = TAKEN;
                            bkey = -2.0; bid = 0x7fffffff;
                            for (int u = 0; u < SQ_PER; ++u) {
                                const int t = u * SQ_THREADS + tid;
                                const double k2 = key_of(t);
                                if (sq_better(k2, (int32_t)(img0 + t), bkey, bid)) { bkey = k2; bid = (int32_t)(img0 + t); }
                            }
                        } else {
                            lk = -2.0;
                        }
                    }
                }
                if (tid < topk) { sh.lkey[tid] = sh.nkey[tid]; sh.lid[tid] = sh.nid[tid]; }
                if (tid == 0) sh.ln = n_new;
            }
            __syncthreads();
            if (shard + 1 < s1) fill(shard + 1);
            if (tid == 0) sh.ncand = 0;
            __syncthreads();
        }

        if (tid < topk) {
            const size_t o = ((size_t)q * n_groups + g) * topk + tid;
            const bool have = tid < sh.ln;
            out_key[o] = have ? sh.lkey[tid] : -2.0;
            out_id[o] = have ? sh.lid[tid] : -1;
        }
        __syncthreads();
    }
}

// ---- NS shards per step, 16-bit accumulators -------------------------------------------------------------------------
// The kernel above is latency bound: per (query, shard) every warp walks a fixed chain of dependent loads and the CTA
// passes five barriers.  When every dot of the batch is known to stay below 65535 (Cauchy-Schwarz on the stored squared
// norms: the caller checks sqrt(max |q|^2 * max |d|^2) < 65535) two accumulators share a 32-bit word, so the SAME shared
// memory holds NS = 2 (or, at half the occupancy, 4) shards at once: a word's runs in consecutive shards are contiguous
// in the posting array, so a step walks ONE run of NS times the length per word (fewer, longer requests: 3 sectors per
// 16 postings instead of 2 x 2), and the barriers, the list maintenance and the pointer bookkeeping are paid once per NS
// shards.  A posting's shard inside the step follows from its position against the word's run boundaries.
template <int NS>
struct Sq16Shared {
    uint32_t acc[NS * SQ_R / 2];           // image x of the step: half (x & 1) of word x >> 1
    double ckey[SQ_CAND];
    int cslot[SQ_CAND];
    double lkey[SQ_MAXK], nkey[SQ_MAXK];
    int32_t lid[SQ_MAXK], nid[SQ_MAXK];
    SqScratch ws[SQ_WARPS];
    uint32_t wb[SQ_WARPS][NS - 1][32];     // flat positions where the runs of shards 1 .. NS-1 of a word start
    uint32_t w_run[SQ_WCACHE];
    uint16_t w_tf[SQ_WCACHE];
    uint8_t w_len[SQ_SUB][SQ_WCACHE];
    double r_key[SQ_WARPS];
    int32_t r_id[SQ_WARPS];
    int r_owner[SQ_WARPS];
    int hist[64];                          // dots of a group's first step (threshold bootstrap)
    int ncand, ln, boot;
};

// acc16[x] += val (x = image of the step, val < 65536 and the sum stays below 65536): one 32-bit reduction on the word
__device__ __forceinline__ void sq_add16(uint32_t acc_s, uint32_t x, uint32_t val, bool ok)
{
    const uint32_t addr = (x >> 1) * 4u + acc_s;
    const uint32_t v = val << ((x & 1u) * 16u);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 :: "r"(addr), "r"(v), "r"((uint32_t)ok) : "memory");
}

template <int NS, int MINB, int U>
__global__ void __launch_bounds__(SQ_THREADS, MINB)
score_query16_kernel(const uint32_t *__restrict__ post_off, const uint32_t *__restrict__ postings,
                     const double *__restrict__ img_rnorm, int64_t n_img, int n_shards, const int64_t *__restrict__ q_off,
                     const int32_t *__restrict__ q_node, const float *__restrict__ q_val, int64_t n_query, int topk,
                     const double *__restrict__ shard_rmax, const double *__restrict__ shard_rmin, int n_groups, int per_group,
                     double *__restrict__ out_key, int32_t *__restrict__ out_id)
{
    static_assert(NS == 2 || NS == 4, "shards per step");
    static_assert(SQ_SUB % NS == 0, "a step must not straddle two pointer sub-groups");
    constexpr int IM = NS * SQ_R, PER = IM / SQ_THREADS;
    constexpr uint32_t TAKEN16 = 0xffffu;
    extern __shared__ __align__(16) unsigned char s_raw[];
    Sq16Shared<NS> &sh = *reinterpret_cast<Sq16Shared<NS> *>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t le_mask = (2u << lane) - 1u, lt_mask = (1u << lane) - 1u;
    SqScratch &ws = sh.ws[warp];
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(sh.acc);
    auto half_of = [&](int x) -> uint32_t { return (sh.acc[x >> 1] >> ((x & 1) * 16)) & 0xffffu; };

    for (int64_t item = blockIdx.x; item < n_query * n_groups; item += gridDim.x) {
        const int64_t q = item / n_groups;
        const int g = (int)(item - q * n_groups);
        const int s0 = g * per_group, s1 = s0 + per_group < n_shards ? s0 + per_group : n_shards;
        const int64_t qb = q_off[q], qe = q_off[q + 1];
        auto fill = [&]() {
            uint4 *a4 = reinterpret_cast<uint4 *>(sh.acc);
            for (int t = tid; t < IM / 8; t += SQ_THREADS) a4[t] = make_uint4(0u, 0u, 0u, 0u);
        };
        fill();
        if (tid == 0) { sh.ln = 0; sh.ncand = 0; }
        __syncthreads();

        for (int shard = s0; shard < s1; shard += NS) {
            const int ns_here = s1 - shard < NS ? s1 - shard : NS;
            const int64_t img0 = (int64_t)shard * SQ_R;
            const int n_here = (int)((n_img - img0) < (int64_t)ns_here * SQ_R ? (n_img - img0) : (int64_t)ns_here * SQ_R);
            const uint32_t *off = post_off + shard;
            const int last = n_shards - shard;                     // pointer index of the row's end

            // one call = 32 query words; lane l: runs of lengths l[0..NS) starting at b, one after the other
            auto walk32 = [&](uint32_t b, const uint32_t (&l)[NS], int tf) {
                uint32_t cum[NS];                                  // cum[j] = postings of the word before shard j + 1
                cum[0] = l[0];
#pragma unroll
                for (int j = 1; j < NS; ++j) cum[j] = cum[j - 1] + l[j];
                uint32_t len = cum[NS - 1];
                const bool is_long = len > SQ_LONG;
                unsigned int m = __ballot_sync(VT_FULL, is_long);
                while (m) {                                        // (rare for leaf-word indexes) one run, the whole warp
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t lb = __shfl_sync(VT_FULL, b, src);
                    uint32_t bnd[NS];
#pragma unroll
                    for (int j = 0; j < NS; ++j) bnd[j] = lb + __shfl_sync(VT_FULL, cum[j], src);
                    const uint32_t ltf = (uint32_t)__shfl_sync(VT_FULL, tf, src);
                    for (uint32_t p = lb + lane; p < bnd[NS - 1]; p += 32) {
                        const uint32_t pe = __ldg(postings + p);
                        uint32_t so = 0;
#pragma unroll
                        for (int j = 0; j + 1 < NS; ++j) so += p >= bnd[j];
                        sq_add16(acc_s, (so << SQ_LOCAL_BITS) | (pe & SQ_LOCAL_MASK), ltf * (pe >> SQ_LOCAL_BITS), true);
                    }
                }
                if (is_long) len = 0;
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const uint32_t total = __shfl_sync(VT_FULL, incl, 31);
                if (total == 0) return;
                const uint32_t excl = incl - len;
                ws.head[lane] = 0;
                ws.head[lane + 32] = 0;
                __syncwarp();
                const bool ne = len > 0;
                const unsigned int nem = __ballot_sync(VT_FULL, ne);
                if (ne) {
                    const int rank = __popc(nem & lt_mask);
                    ws.base[rank] = b - excl;
                    ws.tf[rank] = tf;
#pragma unroll
                    for (int j = 0; j + 1 < NS; ++j) sh.wb[warp][j][rank] = excl + cum[j];
                    atomicOr(&ws.head[excl >> 5], 1u << (excl & 31));
                }
                __syncwarp();
                int before = 0;
                for (uint32_t c0 = 0; c0 * 32 < total; c0 += U) {
                    // per slot two registers: the posting, and (shard of the step << 16 | query count) -- a posting's tf
                    // and the query count are below 65536 here, or some dot would not be
                    uint32_t pe[U], ts[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t c = c0 + u, j = c * 32 + lane;
                        const uint32_t h = c * 32 < total ? ws.head[c] : 0u;
                        const int idx = before + __popc(h & le_mask) - 1;
                        before += __popc(h);
                        pe[u] = 0u;
                        ts[u] = 0u;
                        if (j < total) {
                            uint32_t so = 0;
#pragma unroll
                            for (int jj = 0; jj + 1 < NS; ++jj) so += j >= sh.wb[warp][jj][idx];
                            ts[u] = (so << 16) | (uint32_t)ws.tf[idx];
                            pe[u] = __ldg(postings + (ws.base[idx] + j));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        sq_add16(acc_s, ((ts[u] >> 16) << SQ_LOCAL_BITS) | (pe[u] & SQ_LOCAL_MASK),
                                 (ts[u] & 0xffffu) * (pe[u] >> SQ_LOCAL_BITS), ts[u] != 0);
                }
                __syncwarp();
            };
            // run lengths straight from the row pointers (words beyond the cache, runs of 255 and more)
            auto from_pointers = [&](int64_t w, uint32_t &b, uint32_t (&l)[NS], int &tf) {
                const uint32_t *r = off + (size_t)__ldg(q_node + w) * n_shards;
                tf = (int)__ldg(q_val + w);
                uint32_t prev = __ldg(r);
                b = prev;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const uint32_t nxt = __ldg(r + (j + 1 < last ? j + 1 : last));
                    l[j] = j < ns_here ? nxt - prev : 0u;
                    prev = nxt;
                }
            };

            const int sub = (shard - s0) % SQ_SUB;
            if (sub == 0) {
                for (int i = tid; i < SQ_WCACHE; i += SQ_THREADS) {
                    const int64_t w = qb + i;
                    uint32_t p[SQ_SUB + 1];
                    int tf = 0;
                    if (w < qe) {
                        const uint32_t *r = off + (size_t)__ldg(q_node + w) * n_shards;
                        tf = (int)__ldg(q_val + w);
#pragma unroll
                        for (int j = 0; j <= SQ_SUB; ++j) p[j] = __ldg(r + (j < last ? j : last));
                    } else {
#pragma unroll
                        for (int j = 0; j <= SQ_SUB; ++j) p[j] = 0u;
                    }
                    sh.w_run[i] = p[0];
                    sh.w_tf[i] = (uint16_t)(tf < 65535 ? tf : 65535);
#pragma unroll
                    for (int j = 0; j < SQ_SUB; ++j) {
                        const uint32_t len = p[j + 1] - p[j];
                        sh.w_len[j][i] = (uint8_t)(len < 255u && tf < 65535 ? len : 255u);
                    }
                }
            }
#pragma unroll 1
            for (int i = tid; i < SQ_WCACHE && qb + (i - lane) < qe; i += SQ_THREADS) {
                uint32_t l[NS], b = sh.w_run[i];
                int tf = sh.w_tf[i];
                bool esc = false;
                uint32_t adv = 0;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const uint32_t lj = sh.w_len[sub + j][i];
                    esc |= lj == 255u;
                    adv += lj;
                    l[j] = j < ns_here ? lj : 0u;
                }
                if (esc) {
                    from_pointers(qb + i, b, l, tf);
                    adv = 0;
#pragma unroll
                    for (int j = 0; j < NS; ++j) adv += l[j];       // (a group's last step may stop early: nothing follows it)
                }
                sh.w_run[i] = b + adv;
                walk32(b, l, tf);
            }
            for (int64_t w0 = qb + SQ_WCACHE + (int64_t)warp * 32; w0 < qe; w0 += SQ_THREADS) {
                const int64_t w = w0 + lane;
                uint32_t l[NS], b = 0;
                int tf = 0;
#pragma unroll
                for (int j = 0; j < NS; ++j) l[j] = 0u;
                if (w < qe) from_pointers(w, b, l, tf);
                walk32(b, l, tf);
            }
            __syncthreads();

            // ---- selection against the carried list ------------------------------------------------------------------
            const int ln = sh.ln;
            double T = ln == topk ? sh.lkey[topk - 1] : -1.0;
            {
                // first step of a group (empty list): bound the k-th key from the histogram of the step's dots
                double rmin = 0.0;
                if (ln == 0 && shard_rmin) {
                    rmin = 1e300;
                    for (int j = 0; j < ns_here; ++j) {
                        const double r = __ldg(shard_rmin + shard + j);
                        if (r > 0.0) rmin = fmin(rmin, r);         // (a shard without any described image bounds nothing)
                    }
                    if (rmin == 1e300) rmin = 0.0;
                }
                if (rmin > 0.0) {                                  // (block uniform)
                    if (tid < 64) sh.hist[tid] = 0;
                    __syncthreads();
                    const uint4 *h4 = reinterpret_cast<const uint4 *>(sh.acc);
                    for (int u = 0; u < PER / 8; ++u) {
                        const uint4 a = h4[u * SQ_THREADS + tid];
                        const uint32_t av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t d = (av[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
                            if (d >= 1u) atomicAdd(&sh.hist[d < 63u ? d : 63u], 1);
                        }
                    }
                    __syncthreads();
                    if (warp == 0) sq_bootstrap_bin(sh, topk, lane);
                    __syncthreads();
                    if (sh.boot >= 1) T = (double)sh.boot * rmin;
                }
            }
            bool exhaustive = !(T > 0.0);
            if (!exhaustive) {
                double rmax = 0.0;
                for (int j = 0; j < ns_here; ++j) rmax = fmax(rmax, __ldg(shard_rmax + shard + j));
                uint32_t d_lo = 0xffffffffu;                        // (no 16-bit dot reaches it)
                if (rmax > 0.0) {
                    const double x = ceil(T / rmax);
                    if (x < 65536.0) {
                        int d = (int)x;
                        while (d > 1 && (double)(d - 1) * rmax >= T) --d;
                        while ((double)d * rmax < T) ++d;
                        d_lo = (uint32_t)(d < 1 ? 1 : d);
                    }
                }
                const uint4 *a4 = reinterpret_cast<const uint4 *>(sh.acc);
#pragma unroll 2
                for (int u = 0; u < PER / 8; ++u) {
                    const int t8 = u * SQ_THREADS + tid;
                    const uint4 a = a4[t8];
                    const uint32_t mx = __vmaxu2(__vmaxu2(a.x, a.y), __vmaxu2(a.z, a.w));
                    if (max(mx & 0xffffu, mx >> 16) >= d_lo) {
                        const uint32_t av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (((av[i >> 1] >> ((i & 1) * 16)) & 0xffffu) >= d_lo) {
                                const int pos = atomicAdd(&sh.ncand, 1);
                                if (pos < SQ_CAND) sh.cslot[pos] = t8 * 8 + i;
                            }
                    }
                }
                __syncthreads();
                const int nc = sh.ncand;
                if (nc > SQ_CAND) {
                    exhaustive = true;
                } else if (nc > 0) {
                    for (int i = tid; i < nc; i += SQ_THREADS) {
                        const int x = sh.cslot[i];
                        sh.ckey[i] = (double)half_of(x) * __ldg(img_rnorm + img0 + x);
                    }
                    __syncthreads();
                    if (warp == 0) sq_insert_candidates(sh, nc, img0, topk, lane);
                }
            }
            if (exhaustive) {
                auto key_of = [&](int x) -> double {
                    if (x >= n_here) return -2.0;
                    const uint32_t a = half_of(x);
                    if (a == TAKEN16) return -2.0;
                    const double rn = __ldg(img_rnorm + img0 + x);
                    return rn < 0.0 ? -1.0 : (double)a * rn;
                };
                double lk = tid < ln ? sh.lkey[tid] : -2.0;
                const int32_t li = tid < ln ? sh.lid[tid] : 0x7fffffff;
                double bkey = -2.0; int32_t bid = 0x7fffffff;
                for (int u = 0; u < PER; ++u) {
                    const int x = u * SQ_THREADS + tid;
                    const double key = key_of(x);
                    if (sq_better(key, (int32_t)(img0 + x), bkey, bid)) { bkey = key; bid = (int32_t)(img0 + x); }
                }
                int n_new = 0;
                for (int r = 0; r < topk; ++r) {
                    double key = bkey; int32_t id = bid; int owner;
                    if (sq_better(lk, li, key, id)) { key = lk; id = li; }
                    sq_block_best(key, id, owner, sh);
                    if (key > -2.0) n_new = r + 1;
                    if (tid == 0) { sh.nkey[r] = key; sh.nid[r] = key > -2.0 ? id : -1; }
                    if (tid == owner && key > -2.0) {
                        if ((int64_t)id >= img0) {
                            const int x = (int)(id - img0);
                            sh.acc[x >> 1] |= TAKEN16 << ((x & 1) * 16);   // (only the owner writes between two barriers)
                            bkey = -2.0; bid = 0x7fffffff;
                            for (int u = 0; u < PER; ++u) {
                                const int x2 = u * SQ_THREADS + tid;
                                const double k2 = key_of(x2);
                                if (sq_better(k2, (int32_t)(img0 + x2), bkey, bid)) { bkey = k2; bid = (int32_t)(img0 + x2); }
                            }
                        } else {
                            lk = -2.0;
                        }
                    }
                }
                if (tid < topk) { sh.lkey[tid] = sh.nkey[tid]; sh.lid[tid] = sh.nid[tid]; }
                if (tid == 0) sh.ln = n_new;
            }
            __syncthreads();
            if (shard + NS < s1) fill();
            if (tid == 0) sh.ncand = 0;
            __syncthreads();
        }

        if (tid < topk) {
            const size_t o = ((size_t)q * n_groups + g) * topk + tid;
            const bool have = tid < sh.ln;
            out_key[o] = have ? sh.lkey[tid] : -2.0;
            out_id[o] = have ? sh.lid[tid] : -1;
        }
        __syncthreads();
    }
}
}  // namespace

// lists per query vt_score_topk_carried should write: enough (query, group) items to fill the device about four times over,
// groups as long as possible (the exhaustive first shard of a group is paid once per group)
extern "C" int vt_score_group_count(int64_t n_query, int64_t n_img)
{
    if (n_query <= 0 || n_img <= 0) return 1;
    const int64_t n_shards = (n_img + SQ_R - 1) / SQ_R;
    const int64_t want = (int64_t)vt_sm_count() * 4 * 4;
    int64_t n_groups = (want + n_query - 1) / n_query;
    if (n_groups < 1) n_groups = 1;
    if (n_groups > n_shards) n_groups = n_shards;
    const int64_t per = (n_shards + n_groups - 1) / n_groups;
    return (int)((n_shards + per - 1) / per);
}

template <int MODE, int MINB = 4, int U = SQ_U, int G = SQ_G>
static int launch_query(const uint32_t *post_off, const uint32_t *postings, const double *img_rnorm, int64_t n_img, int n_shards,
                        const int64_t *q_off, const int32_t *q_node, const void *q_val, const double *q_rnorm, int64_t n_query,
                        int topk, const double *shard_rmax, const double *shard_rmin, const int32_t *seed, int64_t seed_stride,
                        int n_groups, double *out_key, int32_t *out_id, cudaStream_t s)
{
    auto kern = score_query_kernel<MODE, MINB, U, G>;
    const size_t smem = sizeof(SqSharedT<typename SqTypes<MODE>::A, typename SqTypes<MODE>::W>);
    VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_group = (n_shards + n_groups - 1) / n_groups;
    const int64_t items = n_query * n_groups, cap = (int64_t)vt_sm_count() * MINB;
    kern<<<(int)(items < cap ? items : cap), SQ_THREADS, smem, s>>>(post_off, postings, img_rnorm, n_img, n_shards, q_off, q_node,
                                                                    q_val, q_rnorm, n_query, topk, shard_rmax, shard_rmin, seed,
                                                                    seed_stride, n_groups, per_group, out_key, out_id);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

template <int NS, int MINB, int U = SQ_U>
static int launch_query16(const uint32_t *post_off, const uint32_t *postings, const double *img_rnorm, int64_t n_img,
                          int n_shards, const int64_t *q_off, const int32_t *q_node, const float *q_val, int64_t n_query,
                          int topk, const double *shard_rmax, const double *shard_rmin, int n_groups, double *out_key,
                          int32_t *out_id, cudaStream_t s)
{
    auto kern = score_query16_kernel<NS, MINB, U>;
    const size_t smem = sizeof(Sq16Shared<NS>);
    VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_group = (n_shards + n_groups - 1) / n_groups;
    const int64_t items = n_query * n_groups, cap = (int64_t)vt_sm_count() * MINB;
    kern<<<(int)(items < cap ? items : cap), SQ_THREADS, smem, s>>>(post_off, postings, img_rnorm, n_img, n_shards, q_off, q_node,
                                                                    q_val, n_query, topk, shard_rmax, shard_rmin, n_groups,
                                                                    per_group, out_key, out_id);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_score_topk_carried(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                                     int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                                     const float *d_q_val, int64_t n_query, int topk, const double *d_shard_rmax,
                                     const double *d_shard_rmin, const int32_t *d_seed, int64_t seed_stride, int n_groups,
                                     int shards_per_step, double *d_group_key, int32_t *d_group_id, void *stream)
{
    VT_CHECK_ARG(shards_per_step == 1 || ((shards_per_step == 2 || shards_per_step == 4) && d_seed == nullptr),
                 "shards_per_step is 1, or 2 / 4 (16-bit accumulators: leaf-word indexes whose dots stay below 65535)");
    VT_CHECK_ARG(n_img > 0 && n_ref > 0 && n_query >= 0 && d_img_rnorm, "bad arguments");
    VT_CHECK_ARG(topk >= 1 && topk <= SQ_MAXK, "topk out of range");
    VT_CHECK_ARG((d_seed != nullptr) != (d_shard_rmax != nullptr), "pass either the seed rows (all-nodes) or the shard bounds (leaf words)");
    VT_CHECK_ARG(d_seed == nullptr || seed_stride >= n_img, "seed rows shorter than the image count");
    const int n_shards = (int)((n_img + SQ_R - 1) / SQ_R);
    VT_CHECK_ARG(n_groups >= 1 && n_groups <= n_shards, "n_groups out of range");
    if (n_query == 0) return VT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t *po = (const uint32_t *)d_postings;
    // chunks in flight and CTAs per SM as measured on the 1 M-image leaf-word bench: <NS 2, 4 CTAs, U 4> 240 k q/s, <2, 3, 8>
    // 268 k, <2, 3, 12> 300-310 k, <2, 3, 16> 270 k, <2, 2, 16> 206 k, <4, 2, 8> 201 k: instruction-level parallelism beats warps
    if (shards_per_step == 2)
        return launch_query16<2, 3, 12>(d_post_off, po, d_img_rnorm, n_img, n_shards, d_q_off, d_q_node, d_q_val, n_query, topk,
                                        d_shard_rmax, d_shard_rmin, n_groups, d_group_key, d_group_id, s);
    if (shards_per_step == 4)
        return launch_query16<4, 2>(d_post_off, po, d_img_rnorm, n_img, n_shards, d_q_off, d_q_node, d_q_val, n_query, topk,
                                    d_shard_rmax, d_shard_rmin, n_groups, d_group_key, d_group_id, s);
    // (all-nodes indexes: 3 CTAs with 8 or 12 chunks in flight, or 6 long runs at a time, measured within 2 % or slower)
    if (d_seed)
        return launch_query<SQ_TF_SEEDED>(d_post_off, po, d_img_rnorm, n_img, n_shards, d_q_off, d_q_node, d_q_val, nullptr, n_query,
                                          topk, nullptr, nullptr, d_seed, seed_stride, n_groups, d_group_key, d_group_id, s);
    return launch_query<SQ_TF>(d_post_off, po, d_img_rnorm, n_img, n_shards, d_q_off, d_q_node, d_q_val, nullptr, n_query, topk,
                               d_shard_rmax, d_shard_rmin, nullptr, 0, n_groups, d_group_key, d_group_id, s);
}

// tf-idf / L2 (VT_SCORE_W_L2) with the carried top-k: binary64 accumulators (64 KB per shard: 2 CTAs per SM, 8 chunks in
// flight per lane), candidates filtered by their exact key.  Same keys as vt_score_topk_weighted up to the order of the
// binary64 additions (shared-memory atomics), lists per group like vt_score_topk_carried.
extern "C" int vt_score_topk_weighted_carried(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                                              int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                                              const double *d_qv, const double *d_q_rnorm, int64_t n_query, int topk,
                                              int n_groups, double *d_group_key, int32_t *d_group_id, void *stream)
{
    VT_CHECK_ARG(n_img > 0 && n_ref > 0 && n_query >= 0 && d_img_rnorm && d_q_rnorm, "bad arguments");
    VT_CHECK_ARG(topk >= 1 && topk <= SQ_MAXK, "topk out of range");
    const int n_shards = (int)((n_img + SQ_R - 1) / SQ_R);
    VT_CHECK_ARG(n_groups >= 1 && n_groups <= n_shards, "n_groups out of range");
    if (n_query == 0) return VT_OK;
    // (4, 8, 12 or 16 chunks in flight per lane: 86 k / 91 k / 92 k / 90 k queries/s on the 1 M-image bench)
    return launch_query<SQ_TFIDF_L2, 2, 8, 3>(d_post_off, (const uint32_t *)d_postings, d_img_rnorm, n_img, n_shards, d_q_off,
                                              d_q_node, d_qv, d_q_rnorm, n_query, topk, nullptr, nullptr, nullptr, 0, n_groups,
                                              d_group_key, d_group_id, (cudaStream_t)stream);
}
