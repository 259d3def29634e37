/*
 * vtree_b200.h -- C ABI of the B200-native vocabulary-tree hot path.
 *
 * The reference (epignatelli/scalable-recognition-with-a-vocabulary-tree) is pure Python and
 * has no FFI of its own; its "plugin API" is two duck-typed protocols (encoder:
 * cbir/encoders/new_encoder_example.py:1-4, descriptor: cbir/descriptors/descriptor_base.py:14-21).
 * The entry points below are what a binding for that path binds: one per hot function of
 * cbir/encoders/vocabulary.py and cbir/database.py (cited on each declaration).  The Python
 * package in this repo calls them through ctypes; INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (nothing is allocated or
 *     freed behind the ABI except by vt_quantize_host, which stages through its own buffers);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success, a negative VT_E* code otherwise, and
 *     vt_last_error() then returns a human-readable message (thread-local);
 *   - descriptors are row-major [n, dp] with dp in {32, 64, 128} (rows zero padded up to dp),
 *     dtype VT_U8 (bytes, what SIFT/ORB emit) or VT_F32;
 *   - trees are implicit complete k-ary heaps: node h has children h*k+1 .. h*k+k, level l
 *     starts at (k^l - 1)/(k - 1); d_centroids is float32 [n_heap, dp]; d_internal[h] != 0
 *     iff node h has children; d_ref_id[h] is the id the reference gives that node (DFS
 *     pre-order, vocabulary.py:70-75) or -1 when the node does not exist.
 *   - distances follow the canonical contract in DESIGN.md section 3: the result always equals
 *     the first-wins argmin of the binary64 evaluation of sum (c - x)^2 (fast float32 pass,
 *     provable-margin check, binary64 re-evaluation of contested rows).
 */
#ifndef VTREE_B200_H
#define VTREE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_ABI_VERSION 1

enum { VT_U8 = 0, VT_F32 = 1 };
enum { VT_OK = 0, VT_EINVAL = -1, VT_ECUDA = -2, VT_EUNSUPPORTED = -3, VT_ENOMEM = -4 };
/* scoring modes: integer tf dot product (reference semantics), float dot product (tf-idf,
 * L2-normalised weights), float L1 form of Nister & Stewenius eq. (5) */
enum { VT_SCORE_TF_L2 = 0, VT_SCORE_W_L2 = 1, VT_SCORE_W_L1 = 2 };

int         vt_abi_version(void);
const char *vt_last_error(void);
/* sm_count, cc_major, cc_minor of the current device; fails when no sm_100 device is present */
int vt_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- host staging (Database.index / retrieve_batch: the descriptor arrays of a batch of images, database.py:36-44,
 * are separate host arrays; the upload wants them back to back in pinned memory).  Copies h_src[i] (h_bytes[i] bytes)
 * one after another into h_dst with n_threads memcpy threads.  Host pointers only; no device is touched. */
int vt_host_pack(const void *const *h_src, const int64_t *h_bytes, int64_t n, void *h_dst, int n_threads);

/* ---- descriptor staging (input side of vocabulary.py:29-34 extract_features) ------------- */
/* float32 [n, d] -> bytes [n, dp] zero padded.  *d_flag (device int, caller zeroes it) is set
 * to 1 when some value is not an integer in 0..255 (the conversion would be lossy). */
int vt_stage_f32_to_u8(const float *d_src, int64_t n, int d, uint8_t *d_dst, int dp,
                       int *d_flag, void *stream);
/* generic row padder: [n, d] elements of elem_size bytes -> [n, dp], zero filled */
int vt_stage_pad(const void *d_src, int64_t n, int d, void *d_dst, int dp, int elem_size,
                 void *stream);

/* ---- greedy descent: VocabularyTree.propagate_feature, vocabulary.py:104-124 ------------- */
/* For each descriptor, walks from heap node `start` while the node is internal, picking the
 * child with the smallest L2 distance (first child wins exact ties, vocabulary.py:119).
 *   d_leaf_heap [n]  (nullable) heap index of the stop node
 *   d_word      [n]  (nullable) d_ref_id[stop node] -- the reference's node id
 *   d_stats     [4]  (nullable, uint64, caller zeroes) [0] += rows re-evaluated in binary64 */
int vt_quantize(const void *d_desc, int dtype, int64_t n, int dp,
                const float *d_centroids, const uint8_t *d_internal, const int32_t *d_ref_id,
                int k, int depth, int32_t start,
                int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats, void *stream);

/* Same contract and bit-identical results, organised for LARGE batches: one pass per tree level
 * over the batch sorted by current node (warp-uniform centroid loads served by L1 instead of
 * per-descriptor gathers that miss L2 on deep levels).  Byte descriptors only; d_work is scratch
 * of vt_quantize_sorted_work_bytes(n) bytes. */
int64_t vt_quantize_sorted_work_bytes(int64_t n);
int vt_quantize_sorted(const void *d_desc, int dtype, int64_t n, int dp,
                       const float *d_centroids, const uint8_t *d_internal, const int32_t *d_ref_id,
                       int k, int depth, int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats,
                       void *d_work, int64_t work_bytes, void *stream);

/* Measurement variant of vt_quantize_sorted (bench.py): CUDA events around every stage;
 * h_level_ms[depth] / h_sort_ms[depth] (host) receive the per-level kernel and re-sort durations.
 * Synchronises the stream. */
int vt_quantize_sorted_profile(const void *d_desc, int dtype, int64_t n, int dp,
                               const float *d_centroids, const uint8_t *d_internal,
                               const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap,
                               int32_t *d_word, uint64_t *d_stats, void *d_work, int64_t work_bytes,
                               float *h_level_ms, float *h_sort_ms, void *stream);

/* Tensor-core variant of vt_quantize_sorted (tcgen05.mma kind::i8, accumulators in TMEM): after
 * sorting, the 128 descriptors of a tile meet the same <= 16 child centroids, so the distance step
 * is a dense byte contraction done in integers (centroids rounded to 8.16 fixed point and split
 * into three byte planes, exact int32 accumulation, scores compared in units of 2^-7; a provable
 * margin for both roundings + binary64 re-evaluation of the contested rows => ids bit-identical
 * to vt_quantize).  Needs dp == 128, k <= 16 and child centroids in
 * [0, 255.99] (d_stats[1] counts nodes that violate this).  vt_quantize_mma_profile is the
 * measurement variant (see vt_quantize_sorted_profile). */
int64_t vt_quantize_mma_work_bytes(int64_t n, int k, int depth);
int vt_quantize_mma(const void *d_desc, int dtype, int64_t n, int dp,
                    const float *d_centroids, const uint8_t *d_internal, const int32_t *d_ref_id,
                    int k, int depth, int32_t *d_leaf_heap, int32_t *d_word, uint64_t *d_stats,
                    void *d_work, int64_t work_bytes, void *stream);
/* planes depend only on the tree: prepare them once (vt_quantize_mma_planes_bytes(k, depth) bytes) and
 * reuse them for every batch; d_work then only needs 16 n + vt_sort_scratch_bytes(n) bytes */
int64_t vt_quantize_mma_planes_bytes(int k, int depth);
int vt_quantize_mma_prepare(const float *d_centroids, int k, int depth, void *d_planes,
                            uint64_t *d_stats, void *stream);
int vt_quantize_mma_prepared(const void *d_desc, int dtype, int64_t n, int dp,
                             const float *d_centroids, const uint8_t *d_internal,
                             const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap,
                             int32_t *d_word, uint64_t *d_stats, const void *d_planes, void *d_work,
                             int64_t work_bytes, void *stream);
int vt_quantize_mma_profile(const void *d_desc, int dtype, int64_t n, int dp,
                            const float *d_centroids, const uint8_t *d_internal,
                            const int32_t *d_ref_id, int k, int depth, int32_t *d_leaf_heap,
                            int32_t *d_word, uint64_t *d_stats, void *d_work, int64_t work_bytes,
                            float *h_level_ms, float *h_sort_ms, void *stream);

/* Host-buffer convenience used for end-to-end timing: descriptors in (preferably pinned) host
 * memory are streamed to the device in chunks on internal streams (copy/compute overlap), word
 * ids are copied back to h_word.  The tree stays device resident.  `chunk_rows` <= 0 picks a
 * default.  Blocks until h_word is complete.  Not re-entrant (one staging state per process). */
int vt_quantize_host(const void *h_desc, int dtype, int64_t n, int dp,
                     const float *d_centroids, const uint8_t *d_internal, const int32_t *d_ref_id,
                     int k, int depth, int32_t *h_word, int64_t chunk_rows);
/* vt_quantize_host keeps its device staging buffers and streams between calls; this frees them. */
int vt_host_release(void);

/* ---- tree build: VocabularyTree.fit, vocabulary.py:36-76, one LEVEL per call -------------
 * Members of all nodes of a level are processed in one launch.  d_perm[p] is the descriptor
 * row of member p, d_parent[p] the index (within its level) of the node it belongs to; members
 * are grouped by node in stable (original) order.  Child c of level-node j is level-node
 * j*k + c of the next level; d_child_centroids points at the next level's first row. */
/* initial centres: child c of node j <- member floor(c*m_j/k) of node j, for nodes with
 * d_seg_count[j] >= k (others are leaves, vocabulary.py:55).  d_seg_begin/d_seg_count: [n_nodes].
 * Multi-GPU: pass the GLOBAL member count in d_seg_count_global and this rank's global starting
 * rank inside each node in d_seg_rank0 (both nullable => single GPU); rows this rank does not
 * own are left untouched (caller zeroes and all-reduces). */
int vt_kmeans_init(const void *d_desc, int dtype, int dp, const int32_t *d_perm,
                   const int64_t *d_seg_begin, const int64_t *d_seg_count,
                   const int64_t *d_seg_count_global, const int64_t *d_seg_rank0,
                   int64_t n_nodes, int k, float *d_child_centroids, void *stream);
/* assignment + exact accumulation (replaces sklearn's E and M step sums, _kmeans.py:1566-1684):
 *   d_label [n_members] in/out uint8 -- previous labels in, new labels out
 *   d_sums  [n_nodes*k, dp] uint64, d_counts [n_nodes*k] uint64 -- caller zeroes; integer sums
 *            are exact, hence identical on any number of GPUs after an all-reduce
 *   d_changed (uint64) += number of members whose label changed
 *   d_stats as in vt_quantize.  Members whose node is a leaf (d_node_internal[j]==0) are skipped. */
int vt_kmeans_assign(const void *d_desc, int dtype, int dp, const int32_t *d_perm,
                     const int32_t *d_parent, int64_t n_members, const uint8_t *d_node_internal,
                     int64_t n_nodes, int k, const float *d_child_centroids, uint8_t *d_label,
                     uint64_t *d_sums, uint64_t *d_counts, uint64_t *d_changed,
                     uint64_t *d_stats, void *stream);
/* Tensor-core variant (dp == 128, k <= 16): distance step AND the per-cluster sums are tcgen05.mma
 * kind::i8 products on the same shared-memory tile (sums = X^T . one-hot(labels)); results are
 * bit-identical to vt_kmeans_assign.  d_planes = vt_kmeans_planes(...) of the CURRENT centres
 * (vt_kmeans_planes_bytes(n_nodes, k) bytes), refreshed after every vt_kmeans_update. */
int64_t vt_kmeans_planes_bytes(int64_t n_nodes, int k);
int vt_kmeans_planes(const float *d_child_centroids, int64_t n_nodes, int k, int dp, void *d_planes,
                     uint64_t *d_stats, void *stream);
int vt_kmeans_assign_mma(const void *d_desc, int dtype, int dp, const int32_t *d_perm,
                         const int32_t *d_parent, int64_t n_members, const uint8_t *d_node_internal,
                         int64_t n_nodes, int k, const float *d_child_centroids, const void *d_planes,
                         uint8_t *d_label, uint64_t *d_sums, uint64_t *d_counts, uint64_t *d_changed,
                         uint64_t *d_stats, void *stream);
/* centre = (float)((double)sum / (double)count) where count > 0; empty clusters keep their centre */
int vt_kmeans_update(const uint64_t *d_sums, const uint64_t *d_counts, int64_t n_rows, int dp,
                     float *d_child_centroids, void *stream);
/* exact column sums of all rows (root value = mean of all features, vocabulary.py:48-49) */
int vt_column_sums(const void *d_desc, int dtype, int64_t n, int dp, uint64_t *d_sums, void *stream);
/* keys for the stable partition at the end of a level (vocabulary.py:64-66):
 * key[p] = d_parent[p]*k + d_label[p] for members of internal nodes, `sentinel` otherwise */
int vt_partition_keys(const int32_t *d_parent, const uint8_t *d_label, int64_t n_members,
                      const uint8_t *d_node_internal, int k, uint32_t sentinel, uint32_t *d_keys,
                      void *stream);

/* ---- stable LSD radix sort of (uint32 key, uint32 value) pairs on bits [0, n_bits) --------
 * d_tmp_* are scratch of the same size; d_hist is scratch of vt_sort_scratch_bytes(n) bytes.
 * *result_in_tmp tells which buffer pair holds the sorted output. */
int64_t vt_sort_scratch_bytes(int64_t n);
int vt_sort_pairs(uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_tmp_keys, uint32_t *d_tmp_vals,
                  int64_t n, int n_bits, void *d_hist, int *result_in_tmp, void *stream);

/* ---- bag-of-words encoding: VocabularyTree.propagate + embedding, vocabulary.py:78-137 ----
 * d_word [n_desc] reference node ids of the stop nodes, grouped by image (d_img_off [n_img+1]);
 * max_words = largest number of descriptors in one image (<= vt_encode_max_words()).
 * all_nodes == 1 counts every node on the path (reference semantics, vocabulary.py:92-100) using
 * d_parent_ref/d_level_ref ([n_ref] parent id / depth of each reference node); all_nodes = 1 + m
 * keeps only the nodes at level >= m (Nister & Stewenius' level mask: weight 0 above level m; the
 * weights vocabulary.py:131,137 leave commented out); 0 counts stop nodes only ("leaf words").  Two-pass CSR: call with d_out_node == NULL to get d_nnz [n_img]
 * (entries per image), prefix-sum it into d_row_off [n_img+1], call again to fill.
 *   d_out_node/d_out_count : CSR columns (reference node id) and term frequencies
 *   d_sumsq [n_img] uint64 : sum of squared counts (exact squared L2 norm), written in pass 2
 *   d_df    [n_ref] uint32 : (nullable) document frequency, += 1 per (image, node), pass 2 */
int vt_encode(const int32_t *d_word, const int64_t *d_img_off, int64_t n_img, int32_t max_words,
              int all_nodes, const int32_t *d_parent_ref, const uint8_t *d_level_ref, int depth,
              int32_t *d_nnz, const int64_t *d_row_off, int32_t *d_out_node, int32_t *d_out_count,
              uint64_t *d_sumsq, uint32_t *d_df, void *stream);
int vt_encode_max_words(void);
/* Single-pass alternative for all_nodes == 0 (an image then has at most as many entries as descriptors):
 * call vt_encode once with d_row_off = d_img_off AND d_nnz set (the fill pass reports the row lengths too),
 * prefix-sum d_nnz into the packed d_row_off and let vt_csr_compact move the rows there. */
int vt_csr_compact(const int64_t *d_in_off, const int32_t *d_nnz, const int64_t *d_row_off,
                   int64_t n_img, const int32_t *d_node_in, const int32_t *d_count_in,
                   int32_t *d_node_out, int32_t *d_count_out, void *stream);

/* The tcgen05 descent has a second level loop that takes TWO tree levels per launch (csrc/vt_quantize_pair.cu:
 * persistent CTAs, TMA gathers, the second level re-reads the rows from L2: 3 DRAM reads of every row per descent
 * at L=6 instead of 6).  vt_quantize_mma_pair_min(rows) makes batches of at least `rows` rows use it (rows < 0:
 * never, the default; rows == 0: query only) and returns the previous threshold.  Ids are identical either way. */
int64_t vt_quantize_mma_pair_min(int64_t min_rows);

/* ---- inverted file: the dict-per-node postings of vocabulary.py:96-100 as CSR, sharded by image
 * range (vt_score_shard_size() = 8192 images per shard).  Row key = node * n_shards + img / shard_size, so a
 * node's posting list is one contiguous image-ordered run and shards are consecutive ranges of it.
 * A posting is ONE uint32: (tf << 13) | local image id inside its shard (an image has at most
 * vt_encode_max_words() = 8192 descriptors, so tf fits the upper 19 bits).  vt_posting_pack writes, for every
 * forward-CSR entry, its key and that posting as the value to sort with it; vt_sort_pairs (stable => postings
 * stay in image order) leaves the sorted values = the posting array; vt_posting_offsets derives the row
 * pointers d_post_off [n_keys + 1] (uint32) from the sorted keys.  tf-idf / L1 scoring uses the same tf
 * postings (weights enter on the query side, vt_score_topk_weighted). */
int vt_posting_pack(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count,
                    int64_t n_img, int32_t shard_size, int64_t n_ref, uint32_t *d_keys,
                    uint32_t *d_packed, void *stream);
int vt_posting_offsets(const uint32_t *d_sorted_keys, int64_t nnz, int64_t n_keys, uint32_t *d_post_off,
                       void *stream);

/* ---- scoring + top-k: Database.score / Database.retrieve, database.py:59-81 -----------------
 * Queries are a CSR batch (d_q_off [n_query+1], d_q_node, d_q_val = tf as float); mode VT_SCORE_TF_L2 only
 * (the weighted modes: vt_score_topk_weighted).
 * vt_score_topk writes, per (query, shard), the topk best (rank key, global image id) pairs:
 * d_shard_key / d_shard_id [n_query, n_shards, topk] (id -1 = no candidate); optionally every
 * image's rank key in d_all_key [n_query, n_img].  vt_topk_merge merges n_cand candidates per
 * query (shards of one GPU, or the all-gathered lists of several GPUs) into the final topk and
 * converts keys to the reference's score.  Order: key descending == score ascending, ties by
 * ascending image index (the reference's stable sort, database.py:79-80).
 *   d_img_rnorm [n_img]: 1/||tf_d||_2 as double, -1 for images without descriptors (vt_rnorm)
 *   d_shard_empty [n_shards] (nullable): != 0 when the shard holds an image without descriptors;
 *        lets the kernel skip the norm of images a query does not touch
 *   d_shard_rho [n_shards] (nullable): (smallest 1/||tf_d||) / (largest 1/||tf_d||) of the shard, scaled
 *        by (1 - 1e-9), or -1 when the shard holds an image without descriptors; enables the tf-mode
 *        fast selection (integer bound on the dot products, binary64 keys for the survivors only)
 *   long_lists: hint for the tf-mode selection kernel.  0: the posting lists of a query's words are short
 *        inside a shard (leaf-word indexes: a handful of postings per word and shard); != 0: long lists
 *        dominate (all-nodes indexes, where the top tree levels list every image) and the kernel is
 *        launched with a register budget for 6 instead of 4 CTAs per SM.  Results do not depend on it.
 *   d_q_rnorm / d_q_sumsq [n_query], d_img_sumsq [n_img]: norms for the final score (tf mode) */
int vt_score_shard_size(void);
int vt_score_max_topk(void);
int vt_rnorm(const uint64_t *d_sumsq, int64_t n, double *d_rnorm, void *stream);
int vt_score_topk(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                  int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                  const float *d_q_val, int64_t n_query, int mode, int topk, double *d_shard_key,
                  int32_t *d_shard_id, double *d_all_key, const uint8_t *d_shard_empty,
                  const double *d_shard_rho, int long_lists, void *stream);
int vt_topk_merge(const double *d_in_key, const int32_t *d_in_id, int64_t n_query, int n_cand,
                  int topk, int mode, const double *d_q_rnorm, const uint64_t *d_q_sumsq,
                  const uint64_t *d_img_sumsq, double *d_out_key, int32_t *d_out_id,
                  double *d_out_score, void *stream);

/* vt_topk_merge for candidates that already carry their final score (d_in_score [n_query, n_cand]): the merge
 * of the all-gathered per-rank top-k lists of a database sharded by image id (section 8e row 4). */
int vt_topk_merge_scored(const double *d_in_key, const int32_t *d_in_id, const double *d_in_score,
                         int64_t n_query, int n_cand, int topk, double *d_out_key, int32_t *d_out_id,
                         double *d_out_score, void *stream);

/* ---- dense top of an ALL-NODES index (vocabulary.py:92-100,132: every node on the path counts, so the top tree
 * levels list every image).  Levels 0..1 ("top", counts up to the descriptors per image) are int32 rows
 * d_top [n_img, n_top]; levels 2..Ld a uint8 block d_block [n_img, kd] (kd multiple of 128); deeper levels stay
 * postings.  d_col_of_ref [n_ref]: >= 0 dense column, -1 posting node, -2 - t top column t.
 * vt_dense_scatter fills the rows of a batch of images from their all-nodes CSR (raises *d_overflow when a dense
 * count exceeds 255).  vt_dense_dots is the tcgen05 INT8 GEMM D[q][img] = A[q].B[img] + QT[q].T[img] (int32, exact);
 * row counts padded to 256.  vt_score_topk_seeded = vt_score_topk (tf) with the accumulators started from D. */
int vt_dense_scatter(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_img,
                     int64_t img_base, const int32_t *d_col_of_ref, uint8_t *d_block, int64_t kd, int32_t *d_top,
                     int n_top, int *d_overflow, void *stream);
int vt_dense_dots(const uint8_t *d_a, const int32_t *d_qt, const uint8_t *d_b, const int32_t *d_t, int64_t qp,
                  int64_t np, int64_t kd, int n_top, int32_t *d_out, int64_t ldd, void *stream);
int vt_score_topk_seeded(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                         int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                         const float *d_q_val, int64_t n_query, int topk, const int32_t *d_seed,
                         int64_t seed_stride, double *d_shard_key, int32_t *d_shard_id, double *d_all_key,
                         void *stream);

/* ---- tf / L2 scoring with the top-k carried from shard to shard (csrc/vt_score_query.cu; database.py:55-87) ----
 * Same result as vt_score_topk (tf mode) / vt_score_topk_seeded, fewer lists: a CTA takes a query through a group of
 * consecutive shards and keeps the k best (key, id) pairs, whose worst key prunes every later shard.  Output: n_groups
 * lists per query, [n_query, n_groups, topk] (id -1 = empty slot), to be merged by vt_topk_merge with n_cand =
 * n_groups * topk.  vt_score_group_count proposes n_groups for a batch (any value in 1..number of shards is valid).
 * Leaf-word indexes pass d_shard_rmax [n_shards] = max over the shard's images of 1/|d| (0 when all are empty), optionally
 * d_shard_rmin [n_shards] = the min over its described images (0 when there is none; NULL: a group's first step selects
 * exhaustively instead of bounding the k-th key from the histogram of its dots) and
 * d_seed = NULL; all-nodes indexes pass d_seed (vt_dense_dots rows, stride seed_stride) and d_shard_rmax = NULL.
 * shards_per_step: 1 = 32-bit accumulators, one shard at a time (always valid); 2 or 4 = that many consecutive shards
 * per step with 16-bit accumulators -- only without seeds and only when the caller knows every dot of the batch stays
 * below 65535 (Cauchy-Schwarz: sqrt(max sum tf_q^2 * max sum tf_d^2) < 65535).  Same results either way. */
int vt_score_group_count(int64_t n_query, int64_t n_img);
int vt_score_topk_carried(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm, int64_t n_img,
                          int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node, const float *d_q_val,
                          int64_t n_query, int topk, const double *d_shard_rmax, const double *d_shard_rmin,
                          const int32_t *d_seed, int64_t seed_stride, int n_groups, int shards_per_step,
                          double *d_group_key, int32_t *d_group_id, void *stream);

/* ---- dense encoders: any object with embedding(image) (the reference's plugin protocol; its AlexNet encoder,
 * encoders/alexnet.py:15-19) goes through Database.retrieve with dense embeddings.  vt_dense_scores evaluates
 * || d/|d| - q/|q| ||_2 (database.py:66-69) for already normalised binary64 rows, NaN -> 1e6; vt_topk_rows selects the
 * topk smallest per (query, shard) as lists for vt_topk_merge_scored. */
int vt_dense_scores(const double *d_db_unit, int64_t n, int64_t dim, const double *d_q_unit, int64_t n_query,
                    double *d_out, void *stream);
int vt_topk_rows(const double *d_score, int64_t n, int64_t n_query, int topk, double *d_shard_key,
                 int32_t *d_shard_id, double *d_shard_score, void *stream);

/* ---- TF-IDF weighting + L1/L2 normalisation (what vocabulary.py:131,137 leaves commented out; opt-in) ----
 * Everything in binary64; the inverted file stays the exact integer tf index, the weights enter on the query
 * side (csrc/vt_weight.cu).  vt_idf: idf_n = ln(N / df_n), 0 where df_n == 0.  vt_weighted_rnorm: per CSR row
 * 1 / |idf * tf|_2 (norm_l1 == 0) or 1 / |idf * tf|_1, -1 for all-zero rows (0/0 -> NaN -> score 1e6).
 * vt_query_weights: the per-entry query factors of the keys (d_qv, d_qi [nnz_q]).  vt_score_topk_weighted: same
 * outputs as vt_score_topk, mode VT_SCORE_W_L2 (key = cosine, score = sqrt(2 - 2 key)) or VT_SCORE_W_L1
 * (key = sum over common words of q + d - |q - d|, score = 2 - key); merge with vt_topk_merge in the same mode. */
int vt_idf(const int32_t *d_df, int64_t n_ref, int64_t n_total, double *d_idf, void *stream);
int vt_weighted_rnorm(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_img,
                      const double *d_idf, int norm_l1, double *d_rnorm, void *stream);
int vt_query_weights(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_query,
                     const double *d_idf, const double *d_q_rnorm, int norm_l1, double *d_qv, double *d_qi,
                     void *stream);
int vt_score_topk_weighted(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                           int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                           const double *d_qv, const double *d_qi, const double *d_q_rnorm, int64_t n_query,
                           int mode, int topk, double *d_shard_key, int32_t *d_shard_id, double *d_all_key,
                           void *stream);
/* VT_SCORE_W_L2 with the top-k carried from shard to shard (csrc/vt_score_query.cu, binary64 accumulators): n_groups lists
 * per query like vt_score_topk_carried; d_qv from vt_query_weights, d_q_rnorm its per-query norms (<= 0: no weight). */
int vt_score_topk_weighted_carried(const uint32_t *d_post_off, const void *d_postings, const double *d_img_rnorm,
                                   int64_t n_img, int64_t n_ref, const int64_t *d_q_off, const int32_t *d_q_node,
                                   const double *d_qv, const double *d_q_rnorm, int64_t n_query, int topk, int n_groups,
                                   double *d_group_key, int32_t *d_group_id, void *stream);

/* ---- measurement helper (bench.py): best-of-reps FP32 FFMA throughput of the current device in
 * TFLOP/s; d_scratch holds at least sm_count*8*256 floats. */
int vt_measure_fp32_peak(float *d_scratch, int iters, int reps, double *tflops, double *ms_best,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VTREE_B200_H */
