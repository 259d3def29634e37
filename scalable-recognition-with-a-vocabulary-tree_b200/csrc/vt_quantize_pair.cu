// vt_quantize_pair.cu -- TWO tree levels per launch of the level-synchronous tcgen05 descent
// (VocabularyTree.propagate_feature, cbir/encoders/vocabulary.py:104-124; same arithmetic, margins and binary64
// re-evaluation as level_mma_kernel in vt_quantize_mma.cu, hence the same ids).
//
// Why: one level per launch reads every 128-byte descriptor row from DRAM once per level (6 x 12.8 GB for 100 M rows
// at k=10, L=6 against 13.2 GB of algorithmic traffic).  Here a persistent CTA (one per SM) takes a CHUNK of 1024
// consecutive rows of the node-grouped batch through two levels before anything goes back to DRAM:
//   phase 1  the chunk's rows are gathered tile by tile (DRAM -> shared memory), one MMA per tile against the
//            children of the tile's node, the epilogue leaves each row's child index in shared memory;
//   sort     a stable counting sort of the chunk by child index (shared memory only) -- rows of one child become
//            contiguous, so phase-2 tiles again meet a single node;
//   phase 2  the same rows are gathered a second time, microseconds after the first read: 148 SMs x 2 chunks x 128 KB of
//            rows = 38 MB in flight, so this read is served by the 126 MB L2, not by DRAM (ncu: DRAM reads of a launch =
//            one pass over the rows, L2 hit rate 47 %);
//   output   packed paths with TWO new digits (one 8-bit radix pass regroups the batch instead of two split passes).
// DRAM row traffic: 3 reads per descent at L=6 instead of 6.
//
// Roles inside a CTA (544 threads): warps 0-7 are two epilogue groups (group = tile parity; thread = TMEM lane = tile
// row), warps 8-15 producers (chunk bookkeeping together; tiles are gathered by ONE warp each, warp = slot mod 8, with
// cp.async), warp 16 issues the MMAs.  Two rings: 8 shared-memory stages (16 KB row tile + 4 KB planes), freed by
// tcgen05.commit as soon as the MMA has read them, and 16 accumulator slots (32 TMEM columns + the node's norms, flags
// and the tile record), freed by the epilogue.  mbarriers: full (producer warp -> MMA), sfree (commit -> producers),
// acc (commit -> epilogue), afree (epilogue -> producers); per-group completion counters tell the producers when the
// child indices of a chunk are complete.  Phase 1 of chunk i+1 is issued BEFORE phase 2 of chunk i, so the ring does
// not drain while a chunk is being sorted.
// #define LP_DEBUG 1    (bounds checks with printf + trap on the tiles the producers issue and the rows the epilogue writes)
#include "vt_mma_common.cuh"
#include <cuda.h>
#include <stdlib.h>

int vt_sort_pass_impl(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, int64_t n, int shift,
                      uint32_t mask, void *d_hist, cudaStream_t s);
int vt_mma_deferred_output(const uint32_t *rows, const uint32_t *child, int64_t n, const int32_t *ref_id, int32_t *leaf,
                           int32_t *word, cudaStream_t s);

namespace {
constexpr int LP_CHUNK = 1024;                  // rows a CTA takes through both levels together
constexpr int LP_THREADS = 544;                 // warps 0-7 epilogue (two groups), 8-15 producers, 16 MMA issuer
constexpr int LP_PRODUCERS = 256;
constexpr int LP_PER = LP_CHUNK / LP_PRODUCERS; // chunk rows per producer thread
constexpr uint32_t LP_NONE = 0xffffffffu;
constexpr int LPF_PHASE2 = 1, LPF_END = 4;
constexpr int LP_STOP = 31;                     // child-index value of a row that takes no part in phase 2 (k <= 16 < 31)

struct LpTile { int start, len; uint32_t node; int flags; int pos0; int buf; };

struct LpArgs {
    const uint8_t *desc; const uint32_t *perm; const uint32_t *parent; long long n;
    int level, phases, bits, k, mode;           // mode of the final phase: 0 = paths for the next launch, 1 = ids, 2 = (row, child) pairs
    long long n_heap;
    const float *cent; const uint8_t *internal; const int32_t *ref_id; const uint8_t *bimg; const int *cn32;
    const int *cninfo;                          // per node 20 ints (80 B, one bulk copy): the 16 squared norms of cn32, then the
                                                // child mask (bit c set <=> child c is internal), then padding
    uint32_t *keys_out, *perm_out; int32_t *leaf_out, *word_out; unsigned long long *stats;
};

template <int KP>
struct LpCfg {
    using P = PlaneCfg<KP>;
    // Two rings.  Shared-memory stages (row tile + planes) are released as soon as the MMA has read them
    // (tcgen05.commit); accumulator slots (TMEM columns + the node's norms, flags and the tile record) live until the
    // epilogue is through with the tile.  So a row tile occupies shared memory for gather + MMA only.
    static constexpr int STAGES = KP <= 10 ? 8 : 6;
    static constexpr int ACC = 512 / P::TMEM_COLS;              // 16 (k <= 10) or 8 accumulator slots; even: slot parity == epilogue group
    static constexpr int TMEM_ALLOC = 512;
    static constexpr int OFF_B = STAGES * MM_A_BYTES;
    static constexpr int OFF_CN = OFF_B + STAGES * P::B_BYTES;
    static constexpr int OFF_LIST = OFF_CN + ACC * 80;
    static constexpr uint32_t TX_BYTES = MM_A_BYTES + P::B_BYTES + 80;   // bytes one tile brings into shared memory
    // per chunk buffer: six u32 lists (c_row c_key c_node s_row s_key s_node) and two u8 lists (c_cls s_d1); two buffers
    static constexpr int LIST_BYTES = 6 * LP_CHUNK * 4 + 2 * LP_CHUNK;
    static constexpr int OFF_TL = OFF_LIST + 2 * LIST_BYTES;    // tile lengths of the phase being issued (u8)
    static constexpr int SMEM = OFF_TL + LP_CHUNK + 1024;
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lp_timeout(const char *what)
{
    printf("vt level_pair_kernel: %s timeout (block %d thread %d)\n", what, (int)blockIdx.x, (int)threadIdx.x);
    __trap();
}
// spin with a watchdog: a protocol error must end the kernel with a trap, not hang the device
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        // (suspend-time hint: a waiting warp sleeps in the barrier unit instead of spinning through the issue slots the
        // working warps need -- the first version of this kernel spent half of its instructions in this loop)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (done) return;
        __nanosleep(64);
        if ((spins & 1023u) == 1023u) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 8000000000ll) lp_timeout("barrier");   // ~4 s
        }
    }
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ bool producers_all(bool pred)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.and.pred p, 1, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred) : "memory");
    return r != 0;
}

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
__global__ void __launch_bounds__(LP_THREADS, 1)
level_pair_kernel(const LpArgs a, const __grid_constant__ CUtensorMap tmap)
{
    using Cfg = LpCfg<KP>;
    using P = PlaneCfg<KP>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    uint8_t *sB = smem + Cfg::OFF_B;
    int *s_cn = reinterpret_cast<int *>(smem + Cfg::OFF_CN);
    uint8_t *tl_len = smem + Cfg::OFF_TL;
    constexpr int AS = Cfg::ACC;
    // full: gather landed (producers -> MMA);  sfree: MMA has read the stage (commit -> producers);
    // acc: accumulator complete (commit -> epilogue);  afree: epilogue done with the slot (epilogue -> producers)
    __shared__ __align__(8) unsigned long long bar_full[S], bar_sfree[S], bar_acc[AS], bar_afree[AS];
    __shared__ LpTile tinfo[AS];
    __shared__ uint32_t s_tmem;
    __shared__ int s_ntiles, s_nreal;
    __shared__ uint16_t s_tstart[64];                           // phase-2 tile starts of a regular chunk (<= 8 + 31 tiles)
    __shared__ int s_wcnt[8][32], s_run[8][32];
    __shared__ int s_done[8];                                   // tiles each epilogue WARP has completed (its group's tiles, in order)

    struct Lists { uint32_t *c_row, *c_key, *c_node, *s_row, *s_key, *s_node; uint8_t *c_cls, *s_d1; };
    auto lists = [&](int buf) {
        Lists l;
        l.c_row = reinterpret_cast<uint32_t *>(smem + Cfg::OFF_LIST + (size_t)buf * Cfg::LIST_BYTES);
        l.c_key = l.c_row + LP_CHUNK; l.c_node = l.c_key + LP_CHUNK;
        l.s_row = l.c_node + LP_CHUNK; l.s_key = l.s_row + LP_CHUNK; l.s_node = l.s_key + LP_CHUNK;
        l.c_cls = reinterpret_cast<uint8_t *>(l.s_node + LP_CHUNK); l.s_d1 = l.c_cls + LP_CHUNK;
        return l;
    };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, bits = a.bits;
    const uint32_t dmask = (1u << bits) - 1u;

    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(Cfg::TMEM_ALLOC));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_sfree[s], 1); }
        for (int s = 0; s < AS; ++s) { mbar_init(&bar_acc[s], 1); mbar_init(&bar_afree[s], 4); }
        for (int w = 0; w < 8; ++w) s_done[w] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const long long n_chunks = (a.n + LP_CHUNK - 1) / LP_CHUNK;

    if (warp < 8) {
        // ================================ epilogue: two groups of four warps; group = tile parity ======================
        // thread = TMEM lane = row of the tile
        const int grp = warp >> 2, qw = warp & 3, etid = tid & 127;
        unsigned int rechecks = 0;
        int tiles_done = 0;
        const int none = 0x7fffffff;
        for (uint32_t g = (uint32_t)grp;; g += 2) {
            const int st = (int)(g % AS);                       // accumulator slot
            mbar_wait(&bar_acc[st], (g / AS) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const LpTile ti = tinfo[st];
            if (ti.flags & LPF_END) break;
            uint32_t v[P::N];
            const uint32_t taddr = tmem + ((uint32_t)(qw * 32) << 16) + (uint32_t)(st * P::TMEM_COLS);
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
            const Lists L = lists(ti.buf);
            const bool p2 = (ti.flags & LPF_PHASE2) != 0;
            const uint32_t *lrow = p2 ? L.s_row : L.c_row;
            const uint32_t *lkey = p2 ? L.s_key : L.c_key;
            const bool active = etid < ti.len;
            const int idx = ti.start + etid;
            const uint32_t row = active ? lrow[idx] : 0u;
#ifdef LP_DEBUG
            if (active && (long long)row >= a.n) {
                printf("LP_DEBUG bad row: block %d warp %d slot %u idx %d row %u start %d len %d flags %d buf %d\n", (int)blockIdx.x, warp, g, idx, row, ti.start, ti.len, ti.flags, ti.buf);
                __trap();
            }
#endif
            int best = none, second = none, arg = 0;
            if (active) score_children_tree<KP, P::N>(v, s_cn + st * 20, best, second, arg);
            const uint8_t *xrow = a.desc + (size_t)row * 128;
            // (second - 15: the tournament ranks scores coarsened to 16 units, see score_children_tree).  The row tile
            // has left shared memory by now: the rare warps that still hold a candidate read |x|^2 from the row itself.
            bool contested = active && k > 1 && contest_test(best, second - 15, 8323200.f);
            if (__any_sync(VT_FULL, contested)) {
                unsigned int xn = 0;
                if (contested) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 qv = __ldg(reinterpret_cast<const uint4 *>(xrow) + j);
                        xn = __dp4a(qv.x, qv.x, xn); xn = __dp4a(qv.y, qv.y, xn); xn = __dp4a(qv.z, qv.z, xn); xn = __dp4a(qv.w, qv.w, xn);
                    }
                }
                contested = contested && contest_test(best, second - 15, (float)xn * 1.0000002f);
            }
            unsigned int m = __ballot_sync(VT_FULL, contested);
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
                const bool cont = (((uint32_t)s_cn[st * 20 + 16] >> arg) & 1u) != 0;   // the child has children of its own
                const bool final_phase = p2 || a.phases == 1;
                if (!final_phase) {
                    // phase 1 of 2: the child index stays in shared memory; a row whose child is a leaf stops here
                    if (cont) {
                        L.c_cls[idx] = (uint8_t)arg;
                    } else {
                        if (a.leaf_out) a.leaf_out[row] = (int32_t)child;
                        if (a.word_out) a.word_out[row] = a.ref_id[child];
                    }
                } else {
                    const long long pos = (long long)ti.pos0 + etid;
                    if (a.mode == 2) {
                        a.keys_out[pos] = row;
                        a.perm_out[pos] = (uint32_t)child;
                    } else if (a.mode == 1) {
                        if (a.leaf_out) a.leaf_out[row] = (int32_t)child;
                        if (a.word_out) a.word_out[row] = a.ref_id[child];
                    } else {
                        if (!cont) {
                            if (a.leaf_out) a.leaf_out[row] = (int32_t)child;
                            if (a.word_out) a.word_out[row] = a.ref_id[child];
                        }
                        const uint32_t last = cont ? (uint32_t)arg : dmask;
                        const uint32_t old = lkey[idx];
                        a.keys_out[pos] = a.phases == 2 ? (((old << bits) | (uint32_t)L.s_d1[idx]) << bits) | last : (old << bits) | last;
                        a.perm_out[pos] = row;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            ++tiles_done;
            if (lane == 0) {
                mbar_arrive(&bar_afree[st]);
                __threadfence_block();                          // c_cls of this tile before the completion count
                *(volatile int *)&s_done[warp] = tiles_done;    // per warp: a warp that runs ahead must not stand in for a slow one
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) rechecks += __shfl_xor_sync(VT_FULL, rechecks, o);
        if (a.stats && lane == 0 && rechecks) atomicAdd(a.stats, (unsigned long long)rechecks);
    } else if (warp < 16) {
        // ================================ producers ================================
        // Order of work per CTA:  P1(c0)  P1(c1)  sort(c0) P2(c0)  P1(c2)  sort(c1) P2(c1) ...  -- phase 1 of the NEXT chunk is
        // in the pipeline while the child indices of the current chunk are collected, so the ring never drains.
        const int ptid = tid - 256, pw = warp - 8;
        uint32_t g = 0;                                         // pipeline slots handed out so far (the same in every producer warp)
        int issued[2] = {0, 0};                                 // tiles handed to each epilogue group so far
        int snap_a[2] = {0, 0}, snap_b[2] = {0, 0};             // per list buffer: `issued` when its phase 1 had been issued
        const int newbits = bits * a.phases;
        const uint32_t newmask = (1u << newbits) - 1u;
        const long long my_chunks = n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

        // every tile issued before the snapshot has left the epilogue: EVERY warp of both groups (each works in order)
        auto wait_done = [&](int buf) {
            const volatile int *d = s_done;
            long long t0 = 0;
            const int na = snap_a[buf], nb2 = snap_b[buf];
            for (uint32_t spins = 0; d[0] < na || d[1] < na || d[2] < na || d[3] < na || d[4] < nb2 || d[5] < nb2 || d[6] < nb2 || d[7] < nb2;
                 ++spins) {
                __nanosleep(40);
                if ((spins & 1023u) == 1023u) {
                    const long long t = clock64();
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > 8000000000ll) lp_timeout("phase");
                }
            }
            __threadfence_block();
        };
        // rows that take no part in the final phase (stopped earlier, or at phase 1): their slot of the output
        auto none_outputs = [&](const uint32_t *lrow, const uint32_t *lkey, int start, int len, long long base, int me,
                                int stride) {
            if (a.mode == 1) return;
            for (int i = me; i < len; i += stride) {
                const long long pos = base + start + i;
                if (a.mode == 2) {
                    a.keys_out[pos] = 0xffffffffu;
                } else {
                    a.keys_out[pos] = (lkey[start + i] << newbits) | newmask;
                    a.perm_out[pos] = lrow[start + i];
                }
            }
        };
        // Tiles are issued by ONE warp each (warp = slot index mod 8) with the TMA: lane g fetches rows 4g .. 4g+3 of the tile
        // with one tile::gather4 (arbitrary row indices, 128B-swizzled destination = exactly the K-major tile the MMA reads,
        // rows past the tile's length are out of range and arrive as zeros), lanes 0 and 1 bulk-copy the node's planes
        // and its 80-byte record.  Everything completes on the stage's full barrier (expect_tx); nobody waits for data.
        auto issue_tile = [&](uint32_t slot, int start, int len, uint32_t node, const uint32_t *lrow, int phase_flag,
                              long long base, int buf) {
            const int st = (int)(slot % S), as = (int)(slot % AS);
#ifdef LP_DEBUG
            if (!((long long)node * k + k < a.n_heap) || start < 0 || len < 1 || len > MM_TILE || start + len > LP_CHUNK) {
                if (lane == 0) printf("LP_DEBUG bad tile: block %d warp %d slot %u start %d len %d node %u flag %d buf %d\n", (int)blockIdx.x, pw, slot, start, len, node, phase_flag, buf);
                __trap();
            }
#endif
            mbar_wait(&bar_afree[as], ((slot / AS) & 1u) ^ 1u);   // the slot's record / tile info may be rewritten
            mbar_wait(&bar_sfree[st], ((slot / S) & 1u) ^ 1u);    // the MMA of the stage's previous tile has read it
            const uint32_t fullb = smem_u32(&bar_full[st]);
            if (lane == 0) {
                LpTile ti;
                ti.start = start; ti.len = len; ti.node = node; ti.pos0 = (int)(base + start);
                ti.flags = phase_flag; ti.buf = buf;
                tinfo[as] = ti;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fullb), "r"(Cfg::TX_BYTES) : "memory");
            }
            __syncwarp();
            {
                int r4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = 4 * lane + u;
                    r4[u] = r < len ? (int)lrow[start + r] : -1;      // -1: out of range -> zero fill
                }
                asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(smem_u32(sA + (size_t)st * MM_A_BYTES) + lane * 512), "l"(&tmap), "r"(0), "r"(r4[0]), "r"(r4[1]),
                               "r"(r4[2]), "r"(r4[3]), "r"(fullb) : "memory");
            }
            if (lane == 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sB + (size_t)st * P::B_BYTES)), "l"(a.bimg + (size_t)node * P::B_BYTES),
                               "r"((uint32_t)P::B_BYTES), "r"(fullb) : "memory");
            if (lane == 1)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(s_cn + as * 20)), "l"(a.cninfo + (size_t)node * 20), "r"(80u), "r"(fullb) : "memory");
        };
        // slots [g, g + nt) have been handed out: per-group tile counts and the slot counter
        auto take_slots = [&](int nt) {
            const int even = (nt + ((g & 1u) ? 0 : 1)) / 2;
            issued[0] += even;
            issued[1] += nt - even;
            g += (uint32_t)nt;
        };
        // irregular chunks (several nodes, stopped rows): every warp walks the tile list and issues the tiles whose slot it owns
        auto issue_phase = [&](const uint32_t *lrow, const uint32_t *lkey, const uint32_t *lnode, int phase_flag,
                               bool final_phase, long long base, int buf) {
            const int nt = s_ntiles;
            int start = 0, len = 0;
            for (int t = 0; t < nt; ++t) {
                start += len;
                len = (int)tl_len[t] + 1;
                const uint32_t node = lnode[start];
                if (node == LP_NONE) {
                    if (final_phase && (t & 7) == pw) none_outputs(lrow, lkey, start, len, base, lane, 32);
                    continue;
                }
                const uint32_t slot = g;
                take_slots(1);
                if ((int)(slot & 7u) == pw) issue_tile(slot, start, len, node, lrow, phase_flag, base, buf);
            }
        };
        // regular chunks (one node for every row -- the usual case): tile t of phase 1 is rows [128 t, 128 t + 128); the
        // tiles of phase 2 come from the table the sort leaves in s_tstart / tl_len.  A warp goes straight to its tiles.
        auto issue_regular_p1 = [&](const Lists &L, int clen, long long base, int buf) {
            const int nt = (clen + MM_TILE - 1) / MM_TILE;
            const uint32_t node0 = L.c_node[0];
            for (int t = (int)((uint32_t)(pw - (int)(g & 7u)) & 7u); t < nt; t += 8)
                issue_tile(g + (uint32_t)t, t * MM_TILE, clen - t * MM_TILE < MM_TILE ? clen - t * MM_TILE : MM_TILE, node0, L.c_row, 0,
                           base, buf);
            take_slots(nt);
        };
        auto issue_regular_p2 = [&](const Lists &L, int clen, long long base, int buf) {
            const int nt = s_ntiles, stop_off = s_nreal;          // (s_nreal carries the offset of the stopped rows here)
            for (int t = (int)((uint32_t)(pw - (int)(g & 7u)) & 7u); t < nt; t += 8) {
                const int start = s_tstart[t];
#ifdef LP_DEBUG
                if (!((long long)L.s_node[start] * k + k < a.n_heap)) {
                    if (lane == 0) printf("LP_DEBUG p2: level %d block %d warp %d t %d nt %d stop %d clen %d start %d len %d s_node %u s_row %u s_key %u c_node0 %u c_cls0 %d wcnt %d %d %d %d\n",
                                          a.level, (int)blockIdx.x, pw, t, nt, stop_off, clen, start, (int)tl_len[t] + 1, L.s_node[start], L.s_row[start], L.s_key[start], L.c_node[0], (int)L.c_cls[0],
                                          s_wcnt[0][0], s_wcnt[0][1], s_wcnt[0][2], s_wcnt[0][31]);
                }
#endif
                issue_tile(g + (uint32_t)t, start, (int)tl_len[t] + 1, L.s_node[start], L.s_row, LPF_PHASE2, base, buf);
            }
            take_slots(nt);
            none_outputs(L.s_row, L.s_key, stop_off, clen - stop_off, base, ptid, LP_PRODUCERS);
        };

        // chunk rows are fetched one chunk ahead: the global loads of chunk i+1 are in flight while chunk i is issued
        uint32_t rrow[LP_PER], rkey[LP_PER];
        auto fetch = [&](long long i) {
            const long long base = (blockIdx.x + i * (long long)gridDim.x) * LP_CHUNK;
#pragma unroll
            for (int u = 0; u < LP_PER; ++u) {
                const long long pos = base + ptid + u * LP_PRODUCERS;
                const bool in = i < my_chunks && pos < a.n;
                rrow[u] = in ? (a.perm ? a.perm[pos] : (uint32_t)pos) : 0u;
                rkey[u] = (in && a.parent) ? a.parent[pos] : 0u;
            }
        };
        // registers -> lists of buffer `buf` + the tile list of phase 1; returns with s_ntiles / s_nreal valid
        auto commit_load = [&](int clen, int buf) {
            const Lists L = lists(buf);
            // regular chunk: every row carries the same live path -> one node, no per-row tree lookups (below level 0 a
            // live path always ends at an internal node: the launch before only extends paths into internal children)
            if (ptid == 0) s_run[0][0] = (int)rkey[0];
            producers_sync();
            const uint32_t key0 = (uint32_t)s_run[0][0];
            bool same = true;
#pragma unroll
            for (int u = 0; u < LP_PER; ++u)
                if (ptid + u * LP_PRODUCERS < clen) same = same && rkey[u] == key0;
            bool regular = producers_all(same) && (a.level == 0 || (key0 & dmask) != dmask);
            uint32_t node0 = LP_NONE;
            if (regular) {
                node0 = path_to_heap32(key0, a.level, bits, (uint32_t)k);
                if (a.level == 0 && !(((long long)node0 * k + k < a.n_heap) && a.internal[node0])) regular = false;   // a root-only tree
            }
            if (regular) {
#pragma unroll
                for (int u = 0; u < LP_PER; ++u) {
                    const int i = ptid + u * LP_PRODUCERS;
                    if (i < clen) { L.c_row[i] = rrow[u]; L.c_key[i] = key0; L.c_node[i] = node0; L.c_cls[i] = LP_STOP; }
                }
                if (ptid == 0) { s_ntiles = (clen + MM_TILE - 1) / MM_TILE; s_nreal = s_ntiles; }
                producers_sync();
                return true;
            }
            uint32_t rh[LP_PER];
            uint8_t rint[LP_PER];
#pragma unroll
            for (int u = 0; u < LP_PER; ++u) {
                const int i = ptid + u * LP_PRODUCERS;
                const bool alive = i < clen && (a.level == 0 || (rkey[u] & dmask) != dmask);
                rh[u] = alive ? path_to_heap32(rkey[u], a.level, bits, (uint32_t)k) : LP_NONE;
                rint[u] = (alive && ((long long)rh[u] * k + k < a.n_heap)) ? a.internal[rh[u]] : (uint8_t)0;
            }
#pragma unroll
            for (int u = 0; u < LP_PER; ++u) {
                const int i = ptid + u * LP_PRODUCERS;
                if (i >= clen) continue;
                uint32_t node = LP_NONE;
                if (rh[u] != LP_NONE) {
                    if (rint[u]) {
                        node = rh[u];
                    } else {                                     // ragged tree: this node is a leaf
                        if (a.leaf_out) a.leaf_out[rrow[u]] = (int32_t)rh[u];
                        if (a.word_out) a.word_out[rrow[u]] = a.ref_id[rh[u]];
                    }
                }
                L.c_row[i] = rrow[u]; L.c_key[i] = rkey[u]; L.c_node[i] = node; L.c_cls[i] = LP_STOP;
            }
            producers_sync();
            if (pw == 0) build_tiles(L.c_node, clen, tl_len, &s_ntiles, &s_nreal, lane);
            producers_sync();
            return false;
        };
        auto chunk_len = [&](long long i) {
            const long long base = (blockIdx.x + i * (long long)gridDim.x) * LP_CHUNK;
            return (int)((a.n - base) < LP_CHUNK ? (a.n - base) : LP_CHUNK);
        };
        auto chunk_base = [&](long long i) { return (blockIdx.x + i * (long long)gridDim.x) * LP_CHUNK; };
        bool uni[2] = {false, false};                          // per list buffer: its chunk is regular (one node)
        int nreal1[2] = {0, 0};
        auto issue_p1 = [&](long long i) {
            const int buf = (int)(i & 1);
            const Lists L = lists(buf);
            if (uni[buf]) issue_regular_p1(L, chunk_len(i), chunk_base(i), buf);
            else issue_phase(L.c_row, L.c_key, L.c_node, 0, a.phases == 1, chunk_base(i), buf);
            snap_a[buf] = issued[0]; snap_b[buf] = issued[1];
        };

        if (my_chunks > 0) {
            fetch(0);
            uni[0] = commit_load(chunk_len(0), 0);
            nreal1[0] = s_nreal;
            fetch(1);
            issue_p1(0);
        }
        for (long long i = 0; i < my_chunks; ++i) {
            const int buf = (int)(i & 1), nb = buf ^ 1;
            if (i + 1 < my_chunks) {
                // buffer nb held chunk i-1: its phase-1 lists were consumed by sort(i-1) (two-level launches) or are
                // still being read by the epilogue (single-level launches)
                if (a.phases == 1 && i >= 1) wait_done(nb);
                producers_sync();
                uni[nb] = commit_load(chunk_len(i + 1), nb);
                nreal1[nb] = s_nreal;
                fetch(i + 2);
                issue_p1(i + 1);
            }
            if (a.phases == 2) {
                const Lists L = lists(buf);
                const int clen = chunk_len(i);
                const long long base = chunk_base(i);
                producers_sync();
                if (nreal1[buf] == 0) {
                    none_outputs(L.c_row, L.c_key, 0, clen, base, ptid, LP_PRODUCERS);    // nothing alive in this chunk
                } else {
                    wait_done(buf);                              // the epilogue has left every child index in c_cls
                    if (uni[buf]) {
                        // ---- regular chunk: counting sort by child index with shared-memory atomics (the order inside a
                        // class is free: all its rows go to the same child node), tile table from the class counts ----
                        if (ptid < 32) s_wcnt[0][ptid] = 0;
                        producers_sync();
                        int cls[LP_PER], rank[LP_PER];
#pragma unroll
                        for (int u = 0; u < LP_PER; ++u) {
                            const int e = ptid + u * LP_PRODUCERS;
                            cls[u] = e < clen ? (int)L.c_cls[e] : -1;
                            rank[u] = cls[u] >= 0 ? atomicAdd(&s_wcnt[0][cls[u]], 1) : 0;
                        }
                        producers_sync();
                        const int total_c = s_wcnt[0][lane];        // lane = class
                        int incl = total_c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(VT_FULL, incl, o);
                            if (lane >= o) incl += t;
                        }
                        const int cbase = incl - total_c;           // first sorted position of class `lane`
                        const uint32_t node1 = L.c_node[0], key1 = L.c_key[0];
#pragma unroll
                        for (int u = 0; u < LP_PER; ++u) {
                            const int e = ptid + u * LP_PRODUCERS;
                            const int off = __shfl_sync(VT_FULL, cbase, cls[u] >= 0 ? cls[u] : 0);
                            if (cls[u] >= 0) {
                                const int dst = off + rank[u];
                                L.s_row[dst] = L.c_row[e];
                                L.s_key[dst] = key1;
                                L.s_node[dst] = cls[u] == LP_STOP ? LP_NONE : node1 * (uint32_t)k + 1u + (uint32_t)cls[u];
                                L.s_d1[dst] = cls[u] == LP_STOP ? (uint8_t)dmask : (uint8_t)cls[u];
                            }
                        }
                        if (pw == 0) {
                            const int ntile = lane == LP_STOP ? 0 : (total_c + MM_TILE - 1) / MM_TILE;
                            int ti = ntile;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(VT_FULL, ti, o);
                                if (lane >= o) ti += t;
                            }
                            int at = ti - ntile, pos = cbase;
                            for (int r = lane == LP_STOP ? 0 : total_c; r > 0; r -= MM_TILE, pos += MM_TILE) {
                                s_tstart[at] = (uint16_t)pos;
                                tl_len[at++] = (uint8_t)((r < MM_TILE ? r : MM_TILE) - 1);
                            }
                            const int all = __shfl_sync(VT_FULL, ti, 31);
                            if (lane == LP_STOP) { s_ntiles = all; s_nreal = cbase; }     // s_nreal: where the stopped rows start
                        }
                        producers_sync();
                        issue_regular_p2(L, clen, base, buf);
                        continue;
                    }
                    // ---- irregular chunk (several parent nodes): the order inside a class matters ----
                    // ---- stable counting sort of the chunk by child index (LP_STOP = stopped) ----
                    s_wcnt[pw][lane] = 0;
                    __syncwarp();
                    const int w_lo = pw * (LP_CHUNK / 8);
                    for (int w0 = w_lo; w0 < w_lo + LP_CHUNK / 8; w0 += 32) {
                        const int e = w0 + lane;
                        const int c = e < clen ? (int)L.c_cls[e] : 32 + lane;      // padding: every lane its own class
                        const unsigned int mm = __match_any_sync(VT_FULL, c);
                        if (e < clen && lane == __ffs(mm) - 1) s_wcnt[pw][c] += __popc(mm);
                        __syncwarp();
                    }
                    producers_sync();
                    int total_c = 0;                             // rows of class `lane` in the whole chunk
                    {
                        int before = 0;                          // ... of them in the warps ahead of this one
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            const int t = s_wcnt[w][lane];
                            total_c += t;
                            if (w < pw) before += t;
                        }
                        int incl = total_c;                      // exclusive prefix over the classes
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(VT_FULL, incl, o);
                            if (lane >= o) incl += t;
                        }
                        s_run[pw][lane] = incl - total_c + before;
                    }
                    __syncwarp();
                    for (int w0 = w_lo; w0 < w_lo + LP_CHUNK / 8; w0 += 32) {
                        const int e = w0 + lane;
                        const int c = e < clen ? (int)L.c_cls[e] : 32 + lane;
                        const unsigned int mm = __match_any_sync(VT_FULL, c);
                        const int leader = __ffs(mm) - 1;
                        int off = 0;
                        if (e < clen && lane == leader) { off = s_run[pw][c]; s_run[pw][c] = off + __popc(mm); }
                        off = __shfl_sync(VT_FULL, off, leader);
                        if (e < clen) {
                            const int dst = off + __popc(mm & ((1u << lane) - 1u));
                            L.s_row[dst] = L.c_row[e];
                            L.s_key[dst] = L.c_key[e];
                            L.s_node[dst] = c == LP_STOP ? LP_NONE : L.c_node[e] * (uint32_t)k + 1u + (uint32_t)c;
                            L.s_d1[dst] = c == LP_STOP ? (uint8_t)dmask : (uint8_t)c;
                        }
                        __syncwarp();
                    }
                    // ---- tile list of phase 2 ----
                    if (uni[buf]) {
                        // one parent node: the runs are the classes, in class order (lane = class)
                        if (pw == 0) {
                            const int ntile = (total_c + MM_TILE - 1) / MM_TILE;
                            int incl = ntile;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(VT_FULL, incl, o);
                                if (lane >= o) incl += t;
                            }
                            int at = incl - ntile;
                            for (int r = total_c; r > 0; r -= MM_TILE) tl_len[at++] = (uint8_t)((r < MM_TILE ? r : MM_TILE) - 1);
                            const int all = __shfl_sync(VT_FULL, incl, 31), stopped = __shfl_sync(VT_FULL, ntile, LP_STOP);
                            if (lane == 0) { s_ntiles = all; s_nreal = all - stopped; }
                        }
                        producers_sync();
                    } else {
                        producers_sync();
                        if (pw == 0) build_tiles(L.s_node, clen, tl_len, &s_ntiles, &s_nreal, lane);
                        producers_sync();
                    }
                    issue_phase(L.s_row, L.s_key, L.s_node, LPF_PHASE2, true, base, buf);
                }
            }
        }
        for (int e = 0; e < 2; ++e) {   // end markers through the same pipeline, one per epilogue group (warp 8 issues them)
            const int st = (int)(g % S), as = (int)(g % AS);
            if (pw == 0) {
                mbar_wait(&bar_afree[as], ((g / AS) & 1u) ^ 1u);
                mbar_wait(&bar_sfree[st], ((g / S) & 1u) ^ 1u);
                if (lane == 0) {
                    LpTile ti; ti.start = 0; ti.len = 0; ti.node = 0; ti.flags = LPF_END; ti.pos0 = 0; ti.buf = 0; tinfo[as] = ti;
                    mbar_arrive(&bar_full[st]);
                }
            }
            ++g;
        }
    } else {
        // ================================ MMA issuer ================================
        int ends = 0;
        for (uint32_t g = 0; ends < 2; ++g) {
            const int st = (int)(g % S), as = (int)(g % AS);
            mbar_wait(&bar_full[st], (g / S) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int flags = tinfo[as].flags;
            if (flags & LPF_END) {
                if (lane == 0) { mbar_arrive(&bar_acc[as]); mbar_arrive(&bar_sfree[st]); }
                ++ends;
                continue;
            }
            if (lane == 0) {
                const uint64_t a_desc = make_desc(smem_u32(sA + (size_t)st * MM_A_BYTES));
                const uint64_t b_desc = make_desc(smem_u32(sB + (size_t)st * P::B_BYTES));
                const uint32_t d_tmem = tmem + (uint32_t)(as * P::TMEM_COLS);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                // K = 4 x 32 bytes
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                        ::"r"(d_tmem), "l"(a_desc + 2 * ks), "l"(b_desc + 2 * ks), "r"(P::IDESC), "r"(acc), "r"(0u));
                }
                // one commit per consumer: the stage goes back to the producers, the accumulator to the epilogue
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_sfree[st])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_acc[as])) : "memory");
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_ALLOC));
}

// per internal node one 80-byte record: its 16 fixed-point squared norms (cn32) and the child mask (bit c set <=> child c
// is an internal node, the descent continues below it)
__global__ void node_record_kernel(const uint8_t *__restrict__ internal, const int *__restrict__ cn32, long long n_int, int k,
                                   int *__restrict__ cninfo)
{
    for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < n_int; h += (long long)gridDim.x * blockDim.x) {
        uint32_t m = 0;
        for (int c = 0; c < k; ++c) m |= (uint32_t)(internal[h * k + 1 + c] != 0) << c;
        for (int c = 0; c < 16; ++c) cninfo[h * 20 + c] = cn32[h * 16 + c];
        cninfo[h * 20 + 16] = (int)m;
        cninfo[h * 20 + 17] = 0; cninfo[h * 20 + 18] = 0; cninfo[h * 20 + 19] = 0;
    }
}

typedef CUresult (*LpEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of the descriptor matrix [n, 128] bytes for tile::gather4: box = one 128-byte row, 128B swizzle
int make_row_map(const void *desc, long long n, CUtensorMap *out)
{
    static LpEncodeFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            vt_set_error("vt_quantize_mma: cuTensorMapEncodeTiled is not available");
            return VT_EUNSUPPORTED;
        }
        encode = (LpEncodeFn)fn;
    }
    cuuint64_t gdim[2] = {128, (cuuint64_t)n};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {128, 1};
    cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(desc), gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        vt_set_error("vt_quantize_mma: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return VT_ECUDA;
    }
    return VT_OK;
}

template <int KP>
int launch_pair(const LpArgs &a, const CUtensorMap &tmap, cudaStream_t s)
{
    static bool attr_set = false;
    auto kern = level_pair_kernel<KP>;
    if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LpCfg<KP>::SMEM));
        attr_set = true;
    }
    const long long n_chunks = (a.n + LP_CHUNK - 1) / LP_CHUNK, cap = (long long)vt_sm_count();
    kern<<<(int)(n_chunks < cap ? n_chunks : cap), LP_THREADS, LpCfg<KP>::SMEM, s>>>(a, tmap);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
}  // namespace

// The level loop with two levels per launch; same contract as vt_mma_levels (vt_quantize_mma.cu).  `ev` (optional)
// receives one event before and one after every launch-pair stage, for the profile variant.
int vt_mma_levels_paired(const void *d_desc, int64_t n, const float *cent, const uint8_t *internal, const int32_t *ref_id,
                         int k, int depth, int32_t *leaf, int32_t *word, uint64_t *stats, const uint8_t *bimg,
                         const int *cn32, int *cninfo, void *sortwork, bool defer, cudaStream_t s, cudaEvent_t *ev,
                         int *n_ev_io)
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
    CUtensorMap tmap;
    {   // per-node records (norms + which children continue the descent), next to the prepared planes; the row map
        int64_t n_int = 0, p = 1;
        for (int l = 0; l < depth; ++l) { n_int += p; p *= k; }
        const int64_t want = (n_int + 255) / 256, cap = (int64_t)vt_sm_count() * 8;
        node_record_kernel<<<(int)(want < cap ? want : cap), 256, 0, s>>>(internal, cn32, n_int, k, cninfo);
        VT_LAUNCH_CHECK();
        const int rc = make_row_map(d_desc, n, &tmap);
        if (rc != VT_OK) return rc;
    }
    for (int level = 0; level < depth;) {
        const int phases = depth - level >= 2 ? 2 : 1;
        const bool last = level + phases == depth;
        LpArgs a;
        a.desc = (const uint8_t *)d_desc; a.perm = cur_perm; a.parent = cur_parent; a.n = n;
        a.level = level; a.phases = phases; a.bits = bits; a.k = k; a.mode = last ? (defer ? 2 : 1) : 0;
        a.n_heap = n_heap; a.cent = cent; a.internal = internal; a.ref_id = ref_id; a.bimg = bimg; a.cn32 = cn32; a.cninfo = cninfo;
        a.keys_out = last && !defer ? nullptr : tkeys; a.perm_out = last && !defer ? nullptr : tperm;
        a.leaf_out = leaf; a.word_out = word; a.stats = (unsigned long long *)stats;
        if (level > 0) VT_MARK();                                 // (the caller recorded the first event)
        int rc = k <= 10 ? launch_pair<10>(a, tmap, s) : launch_pair<16>(a, tmap, s);
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
