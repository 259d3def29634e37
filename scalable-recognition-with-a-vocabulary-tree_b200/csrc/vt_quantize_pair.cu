// vt_quantize_pair.cu -- TWO tree levels per launch of the level-synchronous tcgen05 descent
// (VocabularyTree.propagate_feature, cbir/encoders/vocabulary.py:104-124; same arithmetic, margins and binary64
// re-evaluation as level_mma_kernel in vt_quantize_mma.cu, hence the same ids).
//
// Why: one level per launch reads every 128-byte descriptor row from DRAM once per level (6 x 12.8 GB for 100 M rows
// at k=10, L=6 against 13.2 GB of algorithmic traffic).  Here a persistent CTA takes a CHUNK of 1024 consecutive rows of
// the node-grouped batch through two levels before anything goes back to DRAM:
//   phase 1  the chunk's rows are gathered tile by tile (DRAM -> shared memory), one MMA per tile against the
//            children of the tile's node, the epilogue leaves each row's child index in shared memory;
//   sort     a stable counting sort of the chunk by child index (shared memory only) -- rows of one child become
//            contiguous, so phase-2 tiles again meet a single node;
//   phase 2  the same rows are gathered a second time, microseconds after the first read: 148 SMs x 2 CTAs x 128 KB of
//            rows = 38 MB in flight, so this read is served by the 126 MB L2, not by DRAM;
//   output   packed paths with TWO new digits (one 8-bit radix pass regroups the batch instead of two split passes).
// DRAM row traffic: 3 reads per descent at L=6 instead of 6.
//
// Roles inside a CTA (288 threads, 2 CTAs per SM): warps 0-3 epilogue (thread = TMEM lane = tile row), warps 4-7
// producers (chunk bookkeeping, tile building, cp.async gathers into a 4-stage ring), warp 8 issues the MMAs.  Three
// mbarrier arrays connect them (full: producers -> MMA, acc: tcgen05.commit -> epilogue, empty: epilogue ->
// producers); a fourth barrier tells the producers that the epilogue has finished phase 1 of the chunk.
#include "vt_mma_common.cuh"
#include <stdlib.h>

int vt_sort_pass_impl(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, int64_t n, int shift,
                      uint32_t mask, void *d_hist, cudaStream_t s);
int vt_mma_deferred_output(const uint32_t *rows, const uint32_t *child, int64_t n, const int32_t *ref_id, int32_t *leaf,
                           int32_t *word, cudaStream_t s);

namespace {
constexpr int LP_CHUNK = 1024;                  // rows a CTA takes through both levels together
constexpr int LP_LAG = 2;                       // cp.async groups a producer keeps in flight behind the one it issues
constexpr int LP_THREADS = 288;
constexpr int LP_PRODUCERS = 128;
constexpr uint32_t LP_NONE = 0xffffffffu;
constexpr int LPF_PHASE2 = 1, LPF_LAST_P1 = 2, LPF_END = 4;
constexpr int LP_STOP = 31;                      // child-index value of a row that takes no part in phase 2 (k <= 16 < 31)

struct LpTile { int start, len; uint32_t node; int flags; int pos0; };

struct LpArgs {
    const uint8_t *desc; const uint32_t *perm; const uint32_t *parent; long long n;
    int level, phases, bits, k, mode;           // mode of the final phase: 0 = paths for the next launch, 1 = ids, 2 = (row, child) pairs
    long long n_heap;
    const float *cent; const uint8_t *internal; const int32_t *ref_id; const uint8_t *bimg; const int *cn32;
    uint32_t *keys_out, *perm_out; int32_t *leaf_out, *word_out; unsigned long long *stats;
};

template <int KP>
struct LpCfg {
    using P = PlaneCfg<KP>;
    static constexpr int STAGES = KP <= 10 ? 4 : 3;
    static constexpr int TMEM_ALLOC = KP <= 10 ? 128 : 256;     // STAGES x P::TMEM_COLS, power of two
    static constexpr int OFF_B = STAGES * MM_A_BYTES;
    static constexpr int OFF_CN = OFF_B + STAGES * P::B_BYTES;
    static constexpr int OFF_LIST = OFF_CN + STAGES * 64;
    // six u32 lists (c_row c_key c_node s_row s_key s_node), three u8 lists (c_cls s_d1, tile lengths)
    static constexpr int SMEM = OFF_LIST + 6 * LP_CHUNK * 4 + 3 * LP_CHUNK + 1024;
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// spin with a watchdog: a protocol error must end the kernel with a trap, not hang the device
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((spins & 1023u) == 1023u) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 8000000000ll) {                  // ~4 s
                printf("vt level_pair_kernel: barrier timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Tile list of a node list in shared memory (one warp): maximal runs of equal node ids, cut every 128 rows.  The tiles
// cover the list back to back, so only their lengths (minus one, a byte) are kept.
__device__ __forceinline__ void build_tiles(const uint32_t *nodes, int len, uint8_t *tl_len, int *n_tiles, int *n_real,
                                            int lane)
{
    int cnt = 0, real = 0, run_start = 0;
    uint32_t run_node = LP_NONE, prev = 0xfffffffeu;          // 0xfffffffe never occurs: the first row always starts a run
    for (int w0 = 0; w0 < len; w0 += 32) {
        const int i = w0 + lane;
        const uint32_t nd = i < len ? nodes[i] : 0xfffffffdu;
        uint32_t up = __shfl_up_sync(VT_FULL, nd, 1);
        if (lane == 0) up = prev;
        unsigned int m = __ballot_sync(VT_FULL, i < len && nd != up);
        while (m) {
            const int p = __ffs(m) - 1;
            m &= m - 1;
            const int pos = w0 + p;
            for (int s = run_start; s < pos; s += MM_TILE) {     // close the run [run_start, pos)
                if (lane == 0) tl_len[cnt] = (uint8_t)((pos - s < MM_TILE ? pos - s : MM_TILE) - 1);
                ++cnt;
                real += run_node != LP_NONE;
            }
            run_start = pos;
            run_node = __shfl_sync(VT_FULL, nd, p);
        }
        prev = __shfl_sync(VT_FULL, nd, 31);
    }
    for (int s = run_start; s < len; s += MM_TILE) {
        if (lane == 0) tl_len[cnt] = (uint8_t)((len - s < MM_TILE ? len - s : MM_TILE) - 1);
        ++cnt;
        real += run_node != LP_NONE;
    }
    if (lane == 0) { *n_tiles = cnt; *n_real = real; }
}

template <int KP>
__global__ void __launch_bounds__(LP_THREADS, 2)
level_pair_kernel(const LpArgs a)
{
    using Cfg = LpCfg<KP>;
    using P = PlaneCfg<KP>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    uint8_t *sB = smem + Cfg::OFF_B;
    int *s_cn = reinterpret_cast<int *>(smem + Cfg::OFF_CN);
    uint32_t *c_row = reinterpret_cast<uint32_t *>(smem + Cfg::OFF_LIST);
    uint32_t *c_key = c_row + LP_CHUNK, *c_node = c_key + LP_CHUNK;
    uint32_t *s_row = c_node + LP_CHUNK, *s_key = s_row + LP_CHUNK, *s_node = s_key + LP_CHUNK;
    uint8_t *c_cls = reinterpret_cast<uint8_t *>(s_node + LP_CHUNK);
    uint8_t *s_d1 = c_cls + LP_CHUNK;
    uint8_t *tl_len = s_d1 + LP_CHUNK;
    __shared__ __align__(8) unsigned long long bar_full[S], bar_acc[S], bar_empty[S], bar_p1;
    __shared__ LpTile tinfo[S];
    __shared__ uint32_t s_tmem;
    __shared__ int s_ntiles, s_nreal;
    __shared__ int s_wcnt[4][32], s_run[4][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, bits = a.bits;
    const uint32_t dmask = (1u << bits) - 1u;

    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(Cfg::TMEM_ALLOC));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], LP_PRODUCERS); mbar_init(&bar_acc[s], 1); mbar_init(&bar_empty[s], 4); }
        mbar_init(&bar_p1, 4);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const long long n_chunks = (a.n + LP_CHUNK - 1) / LP_CHUNK;

    if (warp < 4) {
        // ================================ epilogue: thread = TMEM lane = row of the tile ================================
        unsigned int rechecks = 0;
        const int none = 0x7fffffff;
        for (uint32_t g = 0;; ++g) {
            const int st = (int)(g % S);
            mbar_wait(&bar_acc[st], (g / S) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const LpTile ti = tinfo[st];
            if (ti.flags & LPF_END) break;
            uint32_t v[P::N];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * P::TMEM_COLS);
#pragma unroll
            for (int gq = 0; gq < P::N / 16; ++gq) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[16 * gq + 0]), "=r"(v[16 * gq + 1]), "=r"(v[16 * gq + 2]), "=r"(v[16 * gq + 3]),
                               "=r"(v[16 * gq + 4]), "=r"(v[16 * gq + 5]), "=r"(v[16 * gq + 6]), "=r"(v[16 * gq + 7]),
                               "=r"(v[16 * gq + 8]), "=r"(v[16 * gq + 9]), "=r"(v[16 * gq + 10]), "=r"(v[16 * gq + 11]),
                               "=r"(v[16 * gq + 12]), "=r"(v[16 * gq + 13]), "=r"(v[16 * gq + 14]), "=r"(v[16 * gq + 15])
                             : "r"(taddr + 16 * gq));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const bool p2 = (ti.flags & LPF_PHASE2) != 0;
            const uint32_t *lrow = p2 ? s_row : c_row;
            const uint32_t *lkey = p2 ? s_key : c_key;
            const bool active = tid < ti.len;
            const int idx = ti.start + tid;
            const uint32_t row = active ? lrow[idx] : 0u;
            int best = none, second = none, arg = 0;
            if (active) score_children<KP, P::N>(v, s_cn + st * 16, k, best, second, arg);
            const uint8_t *tileA = sA + (size_t)st * MM_A_BYTES;
            const bool contested = contested_row(active && k > 1, best, second, tileA, tid);
            unsigned int m = __ballot_sync(VT_FULL, contested);
            const uint8_t *xrow = a.desc + (size_t)row * 128;
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long xq = __shfl_sync(VT_FULL, (unsigned long long)xrow, src);
                const int res = canon_argmin_warp<uint8_t, 128>(reinterpret_cast<const uint8_t *>(xq),
                                                                a.cent + ((size_t)ti.node * k + 1) * 128, k, lane);
                if (lane == src) { arg = res; ++rechecks; }
            }
            if (active) {
                const long long child = (long long)ti.node * k + 1 + arg;
                const bool final_phase = p2 || a.phases == 1;
                if (!final_phase) {
                    // phase 1 of 2: the child index stays in shared memory; a row whose child is a leaf stops here
                    if ((child * k + k < a.n_heap) && a.internal[child]) {
                        c_cls[idx] = (uint8_t)arg;
                    } else {
                        if (a.leaf_out) a.leaf_out[row] = (int32_t)child;
                        if (a.word_out) a.word_out[row] = a.ref_id[child];
                    }
                } else {
                    const long long pos = (long long)ti.pos0 + tid;
                    if (a.mode == 2) {
                        a.keys_out[pos] = row;
                        a.perm_out[pos] = (uint32_t)child;
                    } else if (a.mode == 1) {
                        if (a.leaf_out) a.leaf_out[row] = (int32_t)child;
                        if (a.word_out) a.word_out[row] = a.ref_id[child];
                    } else {
                        const bool cont = (child * k + k < a.n_heap) && a.internal[child];
                        if (!cont) {
                            if (a.leaf_out) a.leaf_out[row] = (int32_t)child;
                            if (a.word_out) a.word_out[row] = a.ref_id[child];
                        }
                        const uint32_t last = cont ? (uint32_t)arg : dmask;
                        const uint32_t old = lkey[idx];
                        a.keys_out[pos] = a.phases == 2 ? (((old << bits) | (uint32_t)s_d1[idx]) << bits) | last : (old << bits) | last;
                        a.perm_out[pos] = row;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_empty[st]);
                if (ti.flags & LPF_LAST_P1) mbar_arrive(&bar_p1);
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
        if (a.stats && lane == 0 && rechecks) atomicAdd(a.stats, (unsigned long long)rechecks);
    } else if (warp < 8) {
        // ================================ producers ================================
        const int ptid = tid - 128, pw = warp - 4;
        uint32_t g = 0, p1_uses = 0;
        int pending = 0;                                        // committed gathers whose full barrier is still to be arrived on
        const int newbits = bits * a.phases;
        const uint32_t newmask = (1u << newbits) - 1u;

        auto arrive_oldest = [&]() {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> tensor core
            mbar_arrive(&bar_full[(g - (uint32_t)pending) % S]);
            --pending;
        };
        auto flush = [&]() {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            while (pending > 0) arrive_oldest();
        };
        // rows that take no part in the final phase (stopped earlier, or at phase 1): their slot of the output
        auto none_outputs = [&](const uint32_t *lrow, const uint32_t *lkey, int start, int len, long long base) {
            if (a.mode == 1) return;
            for (int i = ptid; i < len; i += LP_PRODUCERS) {
                const long long pos = base + start + i;
                if (a.mode == 2) {
                    a.keys_out[pos] = 0xffffffffu;
                } else {
                    a.keys_out[pos] = (lkey[start + i] << newbits) | newmask;
                    a.perm_out[pos] = lrow[start + i];
                }
            }
        };
        auto issue_phase = [&](const uint32_t *lrow, const uint32_t *lkey, const uint32_t *lnode, int phase_flag,
                               bool final_phase, long long base) {
            const int nt = s_ntiles, nreal = s_nreal;
            int seen = 0, start = 0, len = 0;
            for (int t = 0; t < nt; ++t) {
                start += len;
                len = (int)tl_len[t] + 1;
                const uint32_t node = lnode[start];
                if (node == LP_NONE) {
                    if (final_phase) none_outputs(lrow, lkey, start, len, base);
                    continue;
                }
                ++seen;
                const int st = (int)(g % S);
                mbar_wait(&bar_empty[st], ((g / S) & 1u) ^ 1u);
                {   // A tile: 8 lanes copy one 128 B row, 16 rows per pass, swizzled destination (as level_mma_kernel)
                    const int r0 = ptid >> 3, j = ptid & 7;
                    const uint32_t dst0 = smem_u32(sA + (size_t)st * MM_A_BYTES) +
                                          (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
                    const uint8_t *col = a.desc + j * 16;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = r0 + 16 * i;
                        const bool ok = r < len;
                        const uint32_t rr = ok ? lrow[start + r] : 0u;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 2048u * i),
                                     "l"(col + (size_t)rr * 128), "r"(ok ? 16 : 0));
                    }
                }
                {
                    const uint8_t *bsrc = a.bimg + (size_t)node * P::B_BYTES;
                    const uint32_t bdst = smem_u32(sB + (size_t)st * P::B_BYTES);
#pragma unroll
                    for (int i = 0; i < P::B_BYTES / 16 / LP_PRODUCERS; ++i) {
                        const int c = ptid + LP_PRODUCERS * i;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bdst + c * 16), "l"(bsrc + c * 16));
                    }
                    if (ptid < 4)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_cn + st * 16) + ptid * 16),
                                     "l"(a.cn32 + (size_t)node * 16 + ptid * 4));
                }
                if (ptid == 0) {
                    LpTile ti;
                    ti.start = start; ti.len = len; ti.node = node; ti.pos0 = (int)(base + start);
                    ti.flags = phase_flag | ((seen == nreal && (!final_phase || a.phases == 1)) ? LPF_LAST_P1 : 0);
                    tinfo[st] = ti;
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                ++g;
                ++pending;
                if (pending > LP_LAG) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(LP_LAG) : "memory");
                    arrive_oldest();
                }
            }
        };

        for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
            const long long base = chunk * LP_CHUNK;
            const int clen = (int)((a.n - base) < LP_CHUNK ? (a.n - base) : LP_CHUNK);
            // ---- load the chunk: row index, packed path, node (NONE: stopped earlier or a leaf right here) ----
            for (int i = ptid; i < clen; i += LP_PRODUCERS) {
                const long long pos = base + i;
                const uint32_t row = a.perm ? a.perm[pos] : (uint32_t)pos;
                const uint32_t key = a.parent ? a.parent[pos] : 0u;
                uint32_t node = LP_NONE;
                if (a.level == 0 || (key & dmask) != dmask) {
                    const uint32_t h = path_to_heap32(key, a.level, bits, (uint32_t)k);
                    if (((long long)h * k + k < a.n_heap) && a.internal[h]) {
                        node = h;
                    } else {                                     // ragged tree: this node is a leaf
                        if (a.leaf_out) a.leaf_out[row] = (int32_t)h;
                        if (a.word_out) a.word_out[row] = a.ref_id[h];
                    }
                }
                c_row[i] = row; c_key[i] = key; c_node[i] = node; c_cls[i] = LP_STOP;
            }
            producers_sync();
            if (pw == 0) build_tiles(c_node, clen, tl_len, &s_ntiles, &s_nreal, lane);
            producers_sync();
            const int nreal1 = s_nreal;
            issue_phase(c_row, c_key, c_node, 0, a.phases == 1, base);
            if (a.phases == 1 && nreal1 > 0) {
                // single-level launch (odd depth): the epilogue reads c_row / c_key for its output, so the lists must
                // not be reloaded before it is through with this chunk
                flush();
                mbar_wait(&bar_p1, p1_uses & 1u);
                ++p1_uses;
            }
            if (a.phases == 2) {
                if (nreal1 == 0) {
                    none_outputs(c_row, c_key, 0, clen, base);   // nothing alive in this chunk
                } else {
                    flush();
                    mbar_wait(&bar_p1, p1_uses & 1u);            // the epilogue has left every child index in c_cls
                    ++p1_uses;
                    // ---- stable counting sort of the chunk by child index (LP_STOP = stopped) ----
                    s_wcnt[pw][lane] = 0;
                    __syncwarp();
                    const int w_lo = pw * (LP_CHUNK / 4);
                    for (int w0 = w_lo; w0 < w_lo + LP_CHUNK / 4; w0 += 32) {
                        const int i = w0 + lane;
                        const int c = i < clen ? (int)c_cls[i] : 32 + lane;      // padding: every lane its own class
                        const unsigned int mm = __match_any_sync(VT_FULL, c);
                        if (i < clen && lane == __ffs(mm) - 1) s_wcnt[pw][c] += __popc(mm);
                        __syncwarp();
                    }
                    producers_sync();
                    {
                        int b = 0;
                        for (int c = 0; c < lane; ++c) b += s_wcnt[0][c] + s_wcnt[1][c] + s_wcnt[2][c] + s_wcnt[3][c];
                        for (int w = 0; w < pw; ++w) b += s_wcnt[w][lane];
                        s_run[pw][lane] = b;
                    }
                    __syncwarp();
                    for (int w0 = w_lo; w0 < w_lo + LP_CHUNK / 4; w0 += 32) {
                        const int i = w0 + lane;
                        const int c = i < clen ? (int)c_cls[i] : 32 + lane;
                        const unsigned int mm = __match_any_sync(VT_FULL, c);
                        const int leader = __ffs(mm) - 1;
                        int off = 0;
                        if (i < clen && lane == leader) { off = s_run[pw][c]; s_run[pw][c] = off + __popc(mm); }
                        off = __shfl_sync(VT_FULL, off, leader);
                        if (i < clen) {
                            const int dst = off + __popc(mm & ((1u << lane) - 1u));
                            s_row[dst] = c_row[i];
                            s_key[dst] = c_key[i];
                            s_node[dst] = c == LP_STOP ? LP_NONE : c_node[i] * (uint32_t)k + 1u + (uint32_t)c;
                            s_d1[dst] = c == LP_STOP ? (uint8_t)dmask : (uint8_t)c;
                        }
                        __syncwarp();
                    }
                    producers_sync();
                    if (pw == 0) build_tiles(s_node, clen, tl_len, &s_ntiles, &s_nreal, lane);
                    producers_sync();
                    issue_phase(s_row, s_key, s_node, LPF_PHASE2, true, base);
                }
            }
            producers_sync();                                    // tile list / chunk lists are rewritten by the next chunk
        }
        flush();
        {   // end marker through the same pipeline
            const int st = (int)(g % S);
            mbar_wait(&bar_empty[st], ((g / S) & 1u) ^ 1u);
            if (ptid == 0) { LpTile ti; ti.start = 0; ti.len = 0; ti.node = 0; ti.flags = LPF_END; ti.pos0 = 0; tinfo[st] = ti; }
            mbar_arrive(&bar_full[st]);
        }
    } else {
        // ================================ MMA issuer ================================
        for (uint32_t g = 0;; ++g) {
            const int st = (int)(g % S);
            mbar_wait(&bar_full[st], (g / S) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int flags = tinfo[st].flags;
            if (flags & LPF_END) {
                if (lane == 0) mbar_arrive(&bar_acc[st]);
                break;
            }
            if (lane == 0) {
                const uint64_t a_desc = make_desc(smem_u32(sA + (size_t)st * MM_A_BYTES));
                const uint64_t b_desc = make_desc(smem_u32(sB + (size_t)st * P::B_BYTES));
                const uint32_t d_tmem = tmem + (uint32_t)(st * P::TMEM_COLS);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                // K = 4 x 32 bytes
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                        ::"r"(d_tmem), "l"(a_desc + 2 * ks), "l"(b_desc + 2 * ks), "r"(P::IDESC), "r"(acc), "r"(0u));
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_acc[st])) : "memory");
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_ALLOC));
}

template <int KP>
int launch_pair(const LpArgs &a, cudaStream_t s)
{
    static bool attr_set = false;
    static int ctas_per_sm = 2;
    auto kern = level_pair_kernel<KP>;
    if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LpCfg<KP>::SMEM));
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LP_THREADS, LpCfg<KP>::SMEM) == cudaSuccess && occ >= 1)
            ctas_per_sm = occ < 2 ? occ : 2;
        const char *e = getenv("VT_PAIR_CTAS");
        if (e && atoi(e) >= 1 && atoi(e) <= 2) ctas_per_sm = atoi(e);
        attr_set = true;
    }
    const long long n_chunks = (a.n + LP_CHUNK - 1) / LP_CHUNK, cap = (long long)vt_sm_count() * ctas_per_sm;
    kern<<<(int)(n_chunks < cap ? n_chunks : cap), LP_THREADS, LpCfg<KP>::SMEM, s>>>(a);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
}  // namespace

// The level loop with two levels per launch; same contract as vt_mma_levels (vt_quantize_mma.cu).  `ev` (optional)
// receives one event before and one after every launch-pair stage, for the profile variant.
int vt_mma_levels_paired(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id,
                         int k, int depth, int32_t *leaf, int32_t *word, uint64_t *stats, const uint8_t *bimg,
                         const int *cn32, void *sortwork, bool defer, cudaStream_t s, cudaEvent_t *ev, int *n_ev_io)
{
    int n_ev = n_ev_io ? *n_ev_io : 0;
#define VT_MARK() do { if (ev) VT_CUDA(cudaEventRecord(ev[n_ev++], s)); } while (0)
    int64_t n_heap = 0, pw = 1;
    for (int l = 0; l <= depth; ++l) { n_heap += pw; pw *= k; }
    uint32_t *kA = (uint32_t *)sortwork, *kB = kA + n, *pA = kB + n, *pB = pA + n;
    void *hist = (void *)(pB + n);
    const int bits = path_bits(k);
    uint32_t *keys = kA, *perm = pA, *tkeys = kB, *tperm = pB;
    const uint32_t *cur_perm = nullptr, *cur_parent = nullptr;
    for (int level = 0; level < depth;) {
        const int phases = depth - level >= 2 ? 2 : 1;
        const bool last = level + phases == depth;
        LpArgs a;
        a.desc = (const uint8_t *)d_desc; a.perm = cur_perm; a.parent = cur_parent; a.n = n;
        a.level = level; a.phases = phases; a.bits = bits; a.k = k; a.mode = last ? (defer ? 2 : 1) : 0;
        a.n_heap = n_heap; a.cent = cent; a.internal = internal; a.ref_id = ref_id; a.bimg = bimg; a.cn32 = cn32;
        a.keys_out = last && !defer ? nullptr : tkeys; a.perm_out = last && !defer ? nullptr : tperm;
        a.leaf_out = leaf; a.word_out = word; a.stats = (unsigned long long *)stats;
        if (level > 0) VT_MARK();                                 // (the caller recorded the first event)
        int rc = k <= 10 ? launch_pair<10>(a, s) : launch_pair<16>(a, s);
        if (rc != VT_OK) return rc;
        if (last) {
            if (defer) {
                // (row, child) pairs -> 256 buckets of consecutive rows -> ids written bucket by bucket
                int nb = 0;
                while (((int64_t)1 << nb) < n) ++nb;
                rc = vt_sort_pass_impl(tkeys, tperm, keys, perm, n, nb > 8 ? nb - 8 : 0, 0xffu, hist, s);
                if (rc != VT_OK) return rc;
                rc = vt_mma_deferred_output(keys, perm, n, ref_id, leaf, word, s);
                if (rc != VT_OK) return rc;
            }
            VT_MARK();
            break;
        }
        VT_MARK();
        // one stable pass per new digit, newest first: rows of one node were contiguous before, so they end up
        // grouped by (older digit, newer digit) = by their node two levels further down.  8 bits fit one pass.
        if (bits * phases <= 8) {
            rc = vt_sort_pass_impl(tkeys, tperm, keys, perm, n, 0, (1u << (bits * phases)) - 1u, hist, s);
            if (rc != VT_OK) return rc;
        } else {
            rc = vt_sort_pass_impl(tkeys, tperm, keys, perm, n, 0, (1u << bits) - 1u, hist, s);
            if (rc != VT_OK) return rc;
            rc = vt_sort_pass_impl(keys, perm, tkeys, tperm, n, bits, (1u << bits) - 1u, hist, s);
            if (rc != VT_OK) return rc;
            uint32_t *t; t = keys; keys = tkeys; tkeys = t; t = perm; perm = tperm; tperm = t;
        }
        cur_parent = keys;
        cur_perm = perm;
        level += phases;
    }
#undef VT_MARK
    if (n_ev_io) *n_ev_io = n_ev;
    return VT_OK;
}
