// vt_mma_common.cuh -- pieces shared by the tcgen05 kernels (vt_quantize_mma.cu, vt_quantize_pair.cu): tile geometry,
// shared-memory descriptors of the 128B-swizzled K-major tiles, and the 32-bit epilogue arithmetic with its margin.
#pragma once
#include "vt_common.cuh"

namespace {
constexpr int MM_TILE = 128;                    // descriptors per tile == threads per CTA == UMMA M
constexpr int MM_N = 48;                        // 3 byte planes x 16 child slots (k-means kernel; the descent: PlaneCfg)
constexpr int MM_A_BYTES = MM_TILE * 128;       // 16 KB
constexpr int MM_B_BYTES = MM_N * 128;          // 6 KB
constexpr int MM_TMEM_COLS = 64;                // power of two >= MM_N
constexpr uint32_t MM_IDESC = (2u << 4) | ((uint32_t)(MM_N >> 3) << 17) | ((uint32_t)(MM_TILE >> 4) << 24);
//                             c_format = S32; a_format = b_format = 0 (unsigned 8 bit); K-major A and B

// The descent keeps KP child slots per byte plane: 16 in general, 10 when k <= 10 (the usual vocabulary tree) --
// then the three planes fit N = 32 accumulator columns instead of 48: a third less plane traffic, 32 instead of
// 64 TMEM columns and 16 fewer accumulator registers per thread, which is what lets 8 CTAs share an SM.
template <int KP>
struct PlaneCfg {
    static constexpr int N = (3 * KP + 15) / 16 * 16;
    static constexpr int B_BYTES = N * 128;
    static constexpr int TMEM_COLS = N <= 32 ? 32 : 64;
    static constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(MM_TILE >> 4) << 24);
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of (row, k) inside a 128B-swizzled K-major tile whose rows are 128 bytes
__device__ __forceinline__ uint32_t sw128_offset(int row, int kbyte)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kbyte >> 4) ^ (row & 7))) << 4) + (kbyte & 15));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) /* LBO (unused) */ | (64ull << 32) /* SBO = 1024 B */ |
           (1ull << 46) /* sm100 descriptor version */ | (2ull << 61) /* SWIZZLE_128B */;
}

// ---- epilogue arithmetic ------------------------------------------------------------------------------
// Scores are compared in 32 bits, units 2^-7:  s_c = floor(|c|^2 * 2^7) - (v0 << 8) - v1 - (v2 >> 8)  where
// v0, v1, v2 are the accumulators of the three centroid byte planes (2 x.c * 2^7 = v0 2^8 + v1 + v2 2^-8).
// Each s_c is within ONE unit of the exact score of the rounded centroid, so a gap of the 32-bit scores
// is within two units of the exact gap; the contest test below subtracts three.  Fits int32: v0 < 2^23.
constexpr int MM_S_SHIFT = 25;                  // 2^-32 units -> 2^-7 units

template <int KP, int N>
__device__ __forceinline__ void score_children(const uint32_t (&v)[N], const int *s_cn, int k, int &best, int &second,
                                               int &arg)
{
    int cq[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int4 t = reinterpret_cast<const int4 *>(s_cn)[g];
        cq[4 * g] = t.x; cq[4 * g + 1] = t.y; cq[4 * g + 2] = t.z; cq[4 * g + 3] = t.w;
    }
#pragma unroll
    for (int c = 0; c < KP; ++c) {
        if (c >= k) break;                                       // block-uniform: no predicated-off work
        const int sc = cq[c] - (int)((v[c] << 8) + v[KP + c] + (v[2 * KP + c] >> 8));
        if (sc < best) { second = best; best = sc; arg = c; }
        else second = sc < second ? sc : second;
    }
}

// Same scores, ranked by a tournament instead of a serial best/second chain (the persistent two-level kernel has one
// epilogue warp per scheduler, so the dependency chain of 10 compare-select steps would bound its tile rate).  Every
// score is coarsened to 16 units and carries the child index in its low 4 bits: the minimum of the keys is the best
// child (lowest index among equals), the second smallest key the runner-up.  best / second come back as the
// coarsened scores, i.e. the true scores lie in [best, best + 15] and [second, second + 15]: callers subtract 15
// from `second` before the contest test, so everything this ranking cannot separate goes to the binary64 recheck.
template <int KP, int N>
__device__ __forceinline__ void score_children_tree(const uint32_t (&v)[N], const int *s_cn, int &best, int &second, int &arg)
{
    int cq[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int4 t = reinterpret_cast<const int4 *>(s_cn)[g];
        cq[4 * g] = t.x; cq[4 * g + 1] = t.y; cq[4 * g + 2] = t.z; cq[4 * g + 3] = t.w;
    }
    constexpr int H = KP / 2;
    int lo[H], hi[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {                                // slots >= k hold |c|^2 = 0x3fffffff and zero planes: never the minimum
        const int s0 = cq[2 * i] - (int)((v[2 * i] << 8) + v[KP + 2 * i] + (v[2 * KP + 2 * i] >> 8));
        const int s1 = cq[2 * i + 1] - (int)((v[2 * i + 1] << 8) + v[KP + 2 * i + 1] + (v[2 * KP + 2 * i + 1] >> 8));
        const int k0 = (s0 & ~15) | (2 * i), k1 = (s1 & ~15) | (2 * i + 1);
        lo[i] = min(k0, k1);
        hi[i] = max(k0, k1);
    }
#pragma unroll
    for (int stride = 1; stride < H; stride <<= 1) {
#pragma unroll
        for (int i = 0; i + stride < H; i += 2 * stride) {
            const int la = lo[i], lb = lo[i + stride];
            hi[i] = min(max(la, lb), min(hi[i], hi[i + stride]));
            lo[i] = min(la, lb);
        }
    }
    best = lo[0] & ~15;
    second = hi[0] & ~15;
    arg = lo[0] & 15;
}

// provable margin for the centroid rounding (|error| <= m (da + db), m = 2^-16 sqrt(128) 1.02, da <= db) plus
// the three units of the 32-bit scores: "gap <= 2 m db", squared to avoid the square root.  xn = |x|^2.
__device__ __forceinline__ bool contest_test(int best, int second, float xn)
{
    const float gap = (float)(second - best - 3) * 0.0078125f;
    const float d2b = fmaxf(0.f, (float)(second + 1) * 0.0078125f + xn);
    const float two_m = 2.0f * 1.52587890625e-05f * 11.313708498984761f * 1.02f;
    return gap <= 0.f || gap * gap <= two_m * two_m * 1.001f * d2b;
}

// |x|^2 of the row this thread owns, from the swizzled A tile
__device__ __forceinline__ unsigned int row_norm2(const uint8_t *sA, int tid)
{
    unsigned int xn = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 q = *reinterpret_cast<const uint4 *>(sA + (tid >> 3) * 1024 + (tid & 7) * 128 + ((j ^ (tid & 7)) << 4));
        xn = __dp4a(q.x, q.x, xn); xn = __dp4a(q.y, q.y, xn); xn = __dp4a(q.z, q.z, xn); xn = __dp4a(q.w, q.w, xn);
    }
    return xn;
}

// two-stage contest test: the crude bound |x|^2 <= 128 * 255^2 settles almost every row; only warps that
// still hold a candidate read their rows' exact norms
__device__ __forceinline__ bool contested_row(bool active, int best, int second, const uint8_t *sA, int tid)
{
    bool contested = active && contest_test(best, second, 8323200.f);
    if (__any_sync(VT_FULL, contested)) {
        const unsigned int xn = row_norm2(sA, tid);
        contested = contested && contest_test(best, second, (float)xn * 1.0000002f);
    }
    return contested;
}

}  // namespace
