// vt_dense.cu -- the dense top of the inverted file for the reference's ALL-NODES embedding
// (VocabularyTree.propagate / embedding, cbir/encoders/vocabulary.py:92-100,132: every node on a descriptor's path
// counts, so the top tree levels are visited by (almost) every image and their "posting lists" are the whole database).
//
// Split of a tree with L levels at level Ld (engine.build_dense_index):
//   levels 0..1      "top" nodes, counts up to the number of descriptors of an image: int32 [n_img, n_top]
//   levels 2..Ld     dense block: uint8 counts [n_img, Kd] (K-major rows, Kd a multiple of 128 bytes)
//   levels > Ld      stay postings (vt_score.cu); their lists are short again
// For a query batch the dots over levels 0..Ld are ONE integer GEMM on the tcgen05 INT8 pipe,
//   D[q][img] = sum_col A[q][col] * B[img][col]  +  sum_t QT[q][t] * T[img][t]            (int32, exact)
// and D seeds the shard accumulators of the posting scorer instead of zero (vt_score_topk_seeded), which then adds the
// deep levels and ranks.  Database.retrieve (database.py:72-81) semantics are unchanged: exact integer dots.
//
// dense_gemm_kernel: persistent, one CTA per SM, tile = 256 queries x 256 images, K blocks of 128 bytes through a 3-stage
// TMA ring (two 2-D tile loads per K block: A 32 KB + B 32 KB, 128B-swizzled K-major like the descent kernels), two
// M=128 x N=256 x K=32 kind::i8 MMAs per K step into the two halves of TMEM (512 columns), warps 0-3 epilogue
// (tcgen05.ld, top-level correction with integer FMAs, int32 stores), warp 4 issues the TMA loads, warp 8 the MMAs.
#include "vt_mma_common.cuh"
#include <cuda.h>

namespace {
constexpr int DG_TM = 256, DG_TN = 256, DG_KB = 128;
constexpr int DG_STAGES = 3;
constexpr int DG_A_BYTES = DG_TM * DG_KB, DG_B_BYTES = DG_TN * DG_KB;
constexpr int DG_STAGE_BYTES = DG_A_BYTES + DG_B_BYTES;
constexpr int DG_THREADS = 288;
constexpr int DG_MAX_TOP = 17;                  // 1 + k, k <= 16
constexpr int DG_LAG = 1;
constexpr uint32_t DG_IDESC = (2u << 4) | ((uint32_t)(DG_TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void dg_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dg_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dg_mbar_wait(unsigned long long *bar, uint32_t parity)
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
        if ((spins & 1023u) == 1023u) {                          // watchdog: trap instead of hanging the device
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 8000000000ll) {
                printf("vt dense_gemm_kernel: barrier timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
                __trap();
            }
        }
    }
}

// all-nodes CSR rows of a batch of images -> dense block rows, top rows; entries of the deep levels are left alone
__global__ void dense_scatter_kernel(const int64_t *__restrict__ row_off, const int32_t *__restrict__ node,
                                     const int32_t *__restrict__ count, int64_t n_img, int64_t img_base,
                                     const int32_t *__restrict__ col_of_ref, uint8_t *__restrict__ B, int64_t kd,
                                     int32_t *__restrict__ T, int n_top, int *__restrict__ overflow)
{
    for (int64_t img = blockIdx.x; img < n_img; img += gridDim.x) {
        const int64_t b = row_off[img], e = row_off[img + 1];
        uint8_t *brow = B + (size_t)(img_base + img) * kd;
        int32_t *trow = T + (size_t)(img_base + img) * n_top;
        for (int64_t t = b + threadIdx.x; t < e; t += blockDim.x) {
            const int32_t c = col_of_ref[node[t]];
            const int32_t v = count[t];
            if (c >= 0) {
                if (v > 255) atomicExch(overflow, 1);
                brow[c] = (uint8_t)(v > 255 ? 255 : v);
            } else if (c <= -2) {
                trow[-2 - c] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(DG_THREADS, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const int32_t *__restrict__ QT, const int32_t *__restrict__ T, int64_t qp, int64_t np, int64_t kd, int n_top,
                  int32_t *__restrict__ D, int64_t ldd)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    int32_t *sT = reinterpret_cast<int32_t *>(smem + DG_STAGES * DG_STAGE_BYTES);      // [256][n_top]
    __shared__ __align__(8) unsigned long long bar_full[DG_STAGES], bar_empty[DG_STAGES], bar_tfull, bar_tempty;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < DG_STAGES; ++s) { dg_mbar_init(&bar_full[s], 1); dg_mbar_init(&bar_empty[s], 1); }
        dg_mbar_init(&bar_tfull, 1);
        dg_mbar_init(&bar_tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const int64_t m_tiles = qp / DG_TM, n_tiles = np / DG_TN, tiles = m_tiles * n_tiles;
    const int kblocks = (int)(kd / DG_KB);

    if (warp < 4) {
        // ---------------- epilogue: thread e owns query rows e and e + 128 of the tile ----------------
        uint32_t use = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++use) {
            const int64_t n0 = (t / m_tiles) * DG_TN, m0 = (t % m_tiles) * DG_TM;
            // top rows of the tile's 256 images (levels 0..1) -> shared memory; this thread's two query rows -> registers
            asm volatile("bar.sync 2, 128;" ::: "memory");       // the previous tile's readers are done with sT
            for (int i = tid; i < DG_TN * n_top; i += 128) sT[i] = T[(size_t)n0 * n_top + i];
            int32_t q0[DG_MAX_TOP], q1[DG_MAX_TOP];
#pragma unroll
            for (int j = 0; j < DG_MAX_TOP; ++j) {
                q0[j] = j < n_top ? QT[(size_t)(m0 + tid) * n_top + j] : 0;
                q1[j] = j < n_top ? QT[(size_t)(m0 + 128 + tid) * n_top + j] : 0;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            dg_mbar_wait(&bar_tfull, use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;");
            int32_t *d0 = D + (size_t)(m0 + tid) * ldd + n0, *d1 = D + (size_t)(m0 + 128 + tid) * ldd + n0;
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
            auto drain = [&](const int32_t (&qq)[DG_MAX_TOP], int32_t *dd, uint32_t col0) {
                for (int c0 = 0; c0 < DG_TN; c0 += 16) {
                    uint32_t v[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(lane_addr + col0 + (uint32_t)c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        int32_t acc = (int32_t)v[i];
                        const int32_t *trow = sT + (c0 + i) * n_top;
#pragma unroll
                        for (int j = 0; j < DG_MAX_TOP; ++j)
                            if (j < n_top) acc += qq[j] * trow[j];
                        v[i] = (uint32_t)acc;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<uint4 *>(dd + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            };
            drain(q0, d0, 0u);
            drain(q1, d1, (uint32_t)DG_TN);
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) dg_mbar_arrive(&bar_tempty);
        }
    } else if (warp < 8) {
        // ---------------- producer: one thread, two TMA tile loads per K block ----------------
        if (warp == 4 && lane == 0) {
            uint32_t g = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int n0 = (int)((t / m_tiles) * DG_TN), m0 = (int)((t % m_tiles) * DG_TM);
                for (int kb = 0; kb < kblocks; ++kb, ++g) {
                    const int st = (int)(g % DG_STAGES);
                    dg_mbar_wait(&bar_empty[st], ((g / DG_STAGES) & 1u) ^ 1u);
                    const uint32_t sa = smem_u32(smem + (size_t)st * DG_STAGE_BYTES), sb = sa + DG_A_BYTES, fb = smem_u32(&bar_full[st]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)DG_STAGE_BYTES) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(sa), "l"(&map_a), "r"(kb * DG_KB), "r"(m0), "r"(fb) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(sb), "l"(&map_b), "r"(kb * DG_KB), "r"(n0), "r"(fb) : "memory");
                }
            }
        }
    } else {
        // ---------------- MMA issuer ----------------
        uint32_t g = 0, use = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++use) {
            dg_mbar_wait(&bar_tempty, (use & 1u) ^ 1u);           // the epilogue has drained the accumulators
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int kb = 0; kb < kblocks; ++kb, ++g) {
                const int st = (int)(g % DG_STAGES);
                dg_mbar_wait(&bar_full[st], (g / DG_STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)st * DG_STAGE_BYTES);
                    const uint64_t a0 = make_desc(sa), a1 = make_desc(sa + 128 * DG_KB), bd = make_desc(sa + DG_A_BYTES);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                            ::"r"(tmem), "l"(a0 + 2 * ks), "l"(bd + 2 * ks), "r"(DG_IDESC), "r"(acc), "r"(0u));
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                            ::"r"(tmem + (uint32_t)DG_TN), "l"(a1 + 2 * ks), "l"(bd + 2 * ks), "r"(DG_IDESC), "r"(acc), "r"(0u));
                    }
                    // the stage is free once these MMAs have read it; after the last K block the accumulators are complete
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_empty[st])) : "memory");
                    if (kb == kblocks - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_tfull)) : "memory");
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
}  // namespace

extern "C" int vt_dense_scatter(const int64_t *d_row_off, const int32_t *d_node, const int32_t *d_count, int64_t n_img,
                                int64_t img_base, const int32_t *d_col_of_ref, uint8_t *d_block, int64_t kd, int32_t *d_top,
                                int n_top, int *d_overflow, void *stream)
{
    VT_CHECK_ARG(d_row_off && d_col_of_ref && d_block && d_top && d_overflow, "null buffer");
    VT_CHECK_ARG(n_img >= 0 && kd > 0 && kd % DG_KB == 0 && n_top >= 1 && n_top <= DG_MAX_TOP, "bad shape");
    if (n_img == 0) return VT_OK;
    const int64_t cap = (int64_t)vt_sm_count() * 16;
    dense_scatter_kernel<<<(int)(n_img < cap ? n_img : cap), 128, 0, (cudaStream_t)stream>>>(
        d_row_off, d_node, d_count, n_img, img_base, d_col_of_ref, d_block, kd, d_top, n_top, d_overflow);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

// D[q][img] (int32, row stride ldd) = A[q] . B[img] over the kd dense columns + QT[q] . T[img] over the n_top top
// nodes.  qp and np (padded row counts of A / QT and B / T) are multiples of 256, kd a multiple of 128.
typedef CUresult (*DgEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D byte matrix [rows, kd], box = 256 rows x 128 bytes, 128B swizzle: the K-major tile the MMA descriptors expect
static int dg_make_map(const void *base, int64_t rows, int64_t kd, CUtensorMap *out)
{
    static DgEncodeFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            vt_set_error("vt_dense_dots: cuTensorMapEncodeTiled is not available");
            return VT_EUNSUPPORTED;
        }
        encode = (DgEncodeFn)fn;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kd, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)kd};
    cuuint32_t box[2] = {DG_KB, 256};
    cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        vt_set_error("vt_dense_dots: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return VT_ECUDA;
    }
    return VT_OK;
}

// D[q][img] (int32, row stride ldd) = A[q] . B[img] over the kd dense columns + QT[q] . T[img] over the n_top top
// nodes.  qp and np (padded row counts of A / QT and B / T) are multiples of 256, kd a multiple of 128.
extern "C" int vt_dense_dots(const uint8_t *d_a, const int32_t *d_qt, const uint8_t *d_b, const int32_t *d_t, int64_t qp,
                             int64_t np, int64_t kd, int n_top, int32_t *d_out, int64_t ldd, void *stream)
{
    VT_CHECK_ARG(d_a && d_qt && d_b && d_t && d_out, "null buffer");
    VT_CHECK_ARG(qp > 0 && np > 0 && qp % DG_TM == 0 && np % DG_TN == 0 && kd > 0 && kd % DG_KB == 0 && ldd >= np &&
                 ldd % 4 == 0, "rows must be padded to 256, columns to 128");
    VT_CHECK_ARG(qp < ((int64_t)1 << 31) && np < ((int64_t)1 << 31) && kd < ((int64_t)1 << 31), "matrix too large");
    VT_CHECK_ARG(n_top >= 1 && n_top <= DG_MAX_TOP, "n_top out of range");
    VT_CHECK_ARG((((uintptr_t)d_a | (uintptr_t)d_b) & 15) == 0, "matrices must be 16-byte aligned");
    const int smem = DG_STAGES * DG_STAGE_BYTES + DG_TN * DG_MAX_TOP * 4 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(dense_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    CUtensorMap map_a, map_b;
    int rc = dg_make_map(d_a, qp, kd, &map_a);
    if (rc != VT_OK) return rc;
    rc = dg_make_map(d_b, np, kd, &map_b);
    if (rc != VT_OK) return rc;
    const int64_t tiles = (qp / DG_TM) * (np / DG_TN), cap = vt_sm_count();
    dense_gemm_kernel<<<(int)(tiles < cap ? tiles : cap), DG_THREADS, smem, (cudaStream_t)stream>>>(
        map_a, map_b, d_qt, d_t, qp, np, kd, n_top, d_out, ldd);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
