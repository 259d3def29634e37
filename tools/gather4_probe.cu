// Developer probe (not part of the product): does TMA tile::gather4 deliver four arbitrary 128-byte rows
// into a 128B-swizzled shared-memory tile in the layout tcgen05 expects?  Prints mismatches.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, const int *rows, uint8_t *out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *tile = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) tile[i] = 0xEE;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(32 * 512) : "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x;                     // rows 4g .. 4g+3 of the tile
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(smem_u32(tile) + g * 512), "l"(&tmap), "r"(0), "r"(rows[4 * g]), "r"(rows[4 * g + 1]),
                       "r"(rows[4 * g + 2]), "r"(rows[4 * g + 3]), "r"(smem_u32(&mbar)) : "memory");
    }
    // bounded wait: give up after ~1e6 polls instead of hanging
    uint32_t done = 0;
    for (int spin = 0; spin < 1000000 && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) out[i] = tile[i];
    if (threadIdx.x == 0) out[16384] = (uint8_t)done;
}

static uint32_t sw128_offset(int row, int kbyte)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kbyte >> 4) ^ (row & 7))) << 4) + (kbyte & 15));
}

int main(int argc, char **argv)
{
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const int n_rows = 100000;
    std::vector<uint8_t> h((size_t)n_rows * 128);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    std::vector<int> rows(128);
    for (int i = 0; i < 128; ++i) rows[i] = (int)((i * 7919u + 13u) % n_rows);
    rows[5] = n_rows + 3; rows[77] = -1;             // out of bounds: expect zero fill
    int *d_rows; cudaMalloc(&d_rows, 512); cudaMemcpy(d_rows, rows.data(), 512, cudaMemcpyHostToDevice);
    uint8_t *d_out; cudaMalloc(&d_out, 16385 + 15);

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n"); return 1;
    }
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {128, (cuuint64_t)n_rows};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (box rows %d) -> %d\n", box_rows, (int)r);
    if (r != CUDA_SUCCESS) return 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024);
    probe<<<1, 128, 16384 + 1024>>>(tmap, d_rows, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 3;
    std::vector<uint8_t> out(16385);
    cudaMemcpy(out.data(), d_out, 16385, cudaMemcpyDeviceToHost);
    printf("barrier completed: %d\n", (int)out[16384]);
    long bad = 0, untouched = 0;
    for (int row = 0; row < 128; ++row)
        for (int kb = 0; kb < 128; ++kb) {
            const bool oob = rows[row] < 0 || rows[row] >= n_rows;
            const uint8_t want = oob ? 0 : h[(size_t)rows[row] * 128 + kb];
            const uint8_t got = out[sw128_offset(row, kb)];
            if (got != want) { if (bad < 8) printf("row %d byte %d: got %02x want %02x\n", row, kb, got, want); ++bad; }
            if (got == 0xEE) ++untouched;
        }
    printf("mismatches %ld (0xEE bytes %ld)\n", bad, untouched);
    return bad ? 4 : 0;
}
