/*
 * oracle/vt_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Plain-C CPU restatement of the vocabulary-tree hot path of
 * epignatelli/scalable-recognition-with-a-vocabulary-tree.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (the CUDA library behind
 * include/vtree_b200.h) never does.
 *
 * Parity status: PINNED against the reference itself -- oracle/make_golden.py
 * imports the unmodified reference package from /root/reference, runs it on
 * seeded inputs and stores its outputs under tests/golden/; tests/test_oracle_golden.py
 * checks this file against those vectors.  (The reference ships no tests or golden
 * vectors of its own)
 *
 * Arithmetic contract ("canonical" distance):
 *   dist2(c, x) = sum_d (double(c_d) - double(x_d))^2 evaluated in IEEE binary64,
 *   no FMA contraction, in this fixed order: the (zero padded) dimension axis is cut
 *   into 32 contiguous chunks, each chunk is summed left to right, and the 32 chunk
 *   sums are combined by an xor-butterfly with strides 16, 8, 4, 2, 1.
 *   argmin over children uses strict '<' so the first child wins exact ties, as
 *   cbir/encoders/vocabulary.py:119 does.
 * The reference evaluates the same quantity as sqrt(sdot(c-x, c-x)) in float32
 * (vocabulary.py:117-118, numpy linalg.norm fast path); the two agree except on
 * near-ties (relative gap below ~1e-5) which tests report explicitly.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define VT_OR_CHUNKS 32

/* canonical fp64 squared distance; dp must be a multiple of 32 */
static double canon_dist2(const float *c, const float *x, int dp)
{
    const int cs = dp / VT_OR_CHUNKS;
    double p[VT_OR_CHUNKS], q[VT_OR_CHUNKS];
    for (int l = 0; l < VT_OR_CHUNKS; ++l) {
        double s = 0.0;
        for (int e = 0; e < cs; ++e) {
            const double d = (double)c[l * cs + e] - (double)x[l * cs + e];
            const double sq = d * d;
            s = s + sq;
        }
        p[l] = s;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < VT_OR_CHUNKS; ++l) q[l] = p[l] + p[l ^ off];
        memcpy(p, q, sizeof p);
    }
    return p[0];
}

double vt_or_dist2(const float *c, const float *x, int dp) { return canon_dist2(c, x, dp); }

/* first-wins argmin over k centroid rows; also returns best and runner-up values */
static int argmin_children(const float *x, const float *C, int k, int dp,
                           double *best_out, double *second_out)
{
    double best = INFINITY, second = INFINITY;
    int arg = -1;
    for (int c = 0; c < k; ++c) {
        const double s = canon_dist2(C + (size_t)c * dp, x, dp);
        if (s < best) { second = best; best = s; arg = c; }   /* strict <: vocabulary.py:119 */
        else if (s < second) second = s;
    }
    if (best_out) *best_out = best;
    if (second_out) *second_out = second;
    return arg;
}

/*
 * Assignment step of k-means (the arithmetic the seeded Lloyd uses in place of
 * sklearn MiniBatchKMeans at cbir/encoders/vocabulary.py:62-63) and one level of
 * the greedy descent (vocabulary.py:114-122).
 * X [m,dp] float32, C [k,dp] float32, labels [m] int32,
 * best/second: optional [m] float64 (canonical distances of winner and runner-up).
 */
void vt_or_assign(const float *X, int64_t m, int dp, const float *C, int k,
                  int32_t *labels, double *best, double *second)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; ++i) {
        double b, s;
        labels[i] = argmin_children(X + (size_t)i * dp, C, k, dp, &b, &s);
        if (best) best[i] = b;
        if (second) second[i] = s;
    }
}

/*
 * Seeded exact Lloyd for one tree node (stands in for the unseeded
 * MiniBatchKMeans(n_clusters=k).fit(features) at vocabulary.py:62-63).
 *   init    : centre j = member floor(j*m/k) in the node's (stable) member order
 *   iterate : n_iter times { assign (canonical argmin); exact int64 sums;
 *             centre = (float)((double)sum / (double)count); empty cluster keeps
 *             its previous centre }  -- stops early when labels repeat (fixed point)
 *   labels_ : one more assignment against the final centres
 * X must hold integer values (SIFT/ORB descriptors are bytes) so that the sums are
 * exact and independent of summation order / sharding.
 * Returns the number of update steps performed.
 */
int vt_or_lloyd(const float *X, int64_t m, int dp, int k, int n_iter,
                float *C, int32_t *labels)
{
    int32_t *prev = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    int64_t *sums = (int64_t *)malloc(sizeof(int64_t) * (size_t)k * dp);
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    for (int j = 0; j < k; ++j) {
        const int64_t r = ((int64_t)j * m) / k;
        memcpy(C + (size_t)j * dp, X + (size_t)r * dp, sizeof(float) * dp);
    }
    int done = 0;
    for (int it = 0; it < n_iter; ++it) {
        vt_or_assign(X, m, dp, C, k, labels, NULL, NULL);
        if (it > 0 && memcmp(prev, labels, sizeof(int32_t) * (size_t)m) == 0) break;
        memcpy(prev, labels, sizeof(int32_t) * (size_t)m);
        memset(sums, 0, sizeof(int64_t) * (size_t)k * dp);
        memset(cnt, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < m; ++i) {
            const int c = labels[i];
            cnt[c] += 1;
            const float *x = X + (size_t)i * dp;
            int64_t *s = sums + (size_t)c * dp;
            for (int d = 0; d < dp; ++d) s[d] += (int64_t)x[d];
        }
        for (int j = 0; j < k; ++j) {
            if (cnt[j] == 0) continue;
            for (int d = 0; d < dp; ++d)
                C[(size_t)j * dp + d] = (float)((double)sums[(size_t)j * dp + d] / (double)cnt[j]);
        }
        ++done;
    }
    vt_or_assign(X, m, dp, C, k, labels, NULL, NULL);
    free(prev); free(sums); free(cnt);
    return done;
}

/* exact mean of integer valued rows: root value, vocabulary.py:48-49 */
void vt_or_mean(const float *X, int64_t m, int dp, float *out)
{
    int64_t *sums = (int64_t *)calloc((size_t)dp, sizeof(int64_t));
    for (int64_t i = 0; i < m; ++i)
        for (int d = 0; d < dp; ++d) sums[d] += (int64_t)X[(size_t)i * dp + d];
    for (int d = 0; d < dp; ++d) out[d] = m > 0 ? (float)((double)sums[d] / (double)m) : 0.0f;
    free(sums);
}

/*
 * Greedy descent of a flat tree (VocabularyTree.propagate_feature,
 * vocabulary.py:104-124).  The tree is stored as an implicit complete k-ary heap:
 * node n has children n*k+1 .. n*k+k, internal[n] != 0 iff the reference node has
 * out-degree > 0.  cent is [n_nodes, dp].
 * leaf[i]   : heap index of the node where descriptor i stops
 * path      : optional [n, max_depth+1] heap indices from the root, -1 padded
 * min_gap   : optional [n], smallest relative gap (second-best)/best - 1 met on the way
 */
void vt_or_descend(const float *X, int64_t n, int dp, const float *cent,
                   const uint8_t *internal, int k, int max_depth,
                   int32_t *leaf, int32_t *path, double *min_gap)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float *x = X + (size_t)i * dp;
        int64_t node = 0;
        int depth = 0;
        double g = INFINITY;
        if (path) {
            for (int l = 0; l <= max_depth; ++l) path[i * (max_depth + 1) + l] = -1;
            path[i * (max_depth + 1)] = 0;
        }
        while (depth < max_depth && internal[node]) {
            double b, s;
            const int c = argmin_children(x, cent + (size_t)(node * k + 1) * dp, k, dp, &b, &s);
            const double rel = b > 0.0 ? (s - b) / b : (s > 0.0 ? INFINITY : 0.0);
            if (rel < g) g = rel;
            node = node * k + 1 + c;
            ++depth;
            if (path) path[i * (max_depth + 1) + depth] = (int32_t)node;
        }
        leaf[i] = (int32_t)node;
        if (min_gap) min_gap[i] = g;
    }
}
