// vt_common.cuh -- shared device helpers for the sm_100a vocabulary-tree kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vtree_b200.h"

#define VT_WARP 32
#define VT_FULL 0xffffffffu

// ---- host-side error plumbing (vt_abi.cu owns the storage) ------------------------------------
void vt_set_error(const char *fmt, ...);
#define VT_CHECK_ARG(cond, msg)                                   \
    do { if (!(cond)) { vt_set_error("%s: %s", __func__, msg); return VT_EINVAL; } } while (0)
#define VT_CUDA(call)                                                                   \
    do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                            \
        vt_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));   \
        return VT_ECUDA; } } while (0)
#define VT_LAUNCH_CHECK()                                                               \
    do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) {                \
        vt_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
        return VT_ECUDA; } } while (0)

int vt_sm_count();   // cached multiprocessor count of the current device

// ---- descriptor element access ----------------------------------------------------------------
// Four consecutive dimensions as floats.  Bytes convert exactly.
template <typename T> struct Elem;
template <> struct Elem<uint8_t> {
    static __device__ __forceinline__ float4 load4(const uint8_t *row, int chunk) {
        const uchar4 b = __ldg(reinterpret_cast<const uchar4 *>(row) + chunk);
        return make_float4((float)b.x, (float)b.y, (float)b.z, (float)b.w);
    }
    static __device__ __forceinline__ float load1(const uint8_t *row, int d) { return (float)__ldg(row + d); }
};
template <> struct Elem<float> {
    static __device__ __forceinline__ float4 load4(const float *row, int chunk) {
        return __ldg(reinterpret_cast<const float4 *>(row) + chunk);
    }
    static __device__ __forceinline__ float load1(const float *row, int d) { return __ldg(row + d); }
};

// ---- canonical binary64 distance -----------------------------------------
// Whole warp, lane l owns the contiguous chunk [l*DP/32, (l+1)*DP/32): left-to-right sum inside the
// chunk, then xor-butterfly 16,8,4,2,1.  No FMA contraction.  Every lane returns the same value.
// Mirrors canon_dist2() in oracle/vt_oracle.c bit for bit.
template <typename T, int DP>
__device__ __forceinline__ double canon_dist2_warp(const T *__restrict__ x, const float *__restrict__ c, int lane)
{
    constexpr int CS = DP / 32;
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < CS; ++e) {
        const double d = __dsub_rn((double)__ldg(c + lane * CS + e), (double)Elem<T>::load1(x, lane * CS + e));
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(VT_FULL, s, off));
    return s;
}

// first-wins argmin over k children rows in binary64; whole warp, uniform result
template <typename T, int DP>
__device__ __forceinline__ int canon_argmin_warp(const T *__restrict__ x, const float *__restrict__ children, int k, int lane)
{
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    int arg = 0;
    for (int c = 0; c < k; ++c) {
        const double s = canon_dist2_warp<T, DP>(x, children + (size_t)c * DP, lane);
        if (s < best) { best = s; arg = c; }                     // strict <: vocabulary.py:119
    }
    return arg;
}

// ---- fast float32 pass: 8 lanes per descriptor -------------------------------------------------
// Sub-lane j of a group owns float4 chunks j, j+8, j+16, ... (coalesced 128 B per group per load).
// Returns the group-wide float32 squared distance (identical in the 8 lanes).
template <int DP>
__device__ __forceinline__ float group_dist2(const float4 (&x)[DP / 32], const float *__restrict__ crow, int j,
                                             unsigned int gmask)
{
    constexpr int VEC = DP / 32;
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float4 c = __ldg(reinterpret_cast<const float4 *>(crow) + j + 8 * v);
        float d;
        d = c.x - x[v].x; acc = fmaf(d, d, acc);
        d = c.y - x[v].y; acc = fmaf(d, d, acc);
        d = c.z - x[v].z; acc = fmaf(d, d, acc);
        d = c.w - x[v].w; acc = fmaf(d, d, acc);
    }
    acc += __shfl_xor_sync(gmask, acc, 4);     // group-local mask: groups may diverge from each other
    acc += __shfl_xor_sync(gmask, acc, 2);
    acc += __shfl_xor_sync(gmask, acc, 1);
    return acc;
}

// Relative margin below which the float32 ranking of two candidates is not trusted.  Any float32
// evaluation of a sum of DP non-negative terms (c-x)^2 has relative error <= (DP+2)*2^-24, so two
// candidates whose float32 values differ by more than 2*(DP+2)*2^-24 (relative) are ordered like
// their exact values.  DP*2^-22 = 4*DP*2^-24 leaves a 2x safety factor.
template <int DP> __device__ __forceinline__ float contest_margin() { return (float)DP * 2.384185791015625e-07f; }

// One level of the descent for the 4 descriptor groups of a warp.
//   x        : this lane's slice of its group's descriptor
//   xrow     : pointer to the group's descriptor row (for the binary64 re-evaluation)
//   children : pointer to the k child centroid rows of the group's current node
//   active   : group takes part (warp-uniform control flow is kept by the caller)
// Returns the canonical child index; *rechecks counts binary64 re-evaluations (lane 0 of group).
template <typename T, int DP>
__device__ __forceinline__ int descend_one_level(const float4 (&x)[DP / 32], const T *__restrict__ xrow,
                                                 const float *__restrict__ children, int k, bool active,
                                                 int lane, unsigned int &rechecks)
{
    const int j = lane & 7;
    float best = __int_as_float(0x7f800000), second = __int_as_float(0x7f800000);
    int arg = 0;
    if (active) {                                          // uniform inside a group
        const unsigned int gmask = 0xffu << (lane & 24);
        for (int c = 0; c < k; ++c) {
            const float s = group_dist2<DP>(x, children + (size_t)c * DP, j, gmask);
            if (s < best) { second = best; best = s; arg = c; }
            else if (s < second) second = s;
        }
    }
    const bool contested = active && (k > 1) && (second <= best + best * contest_margin<DP>());
    unsigned int m = __ballot_sync(VT_FULL, contested && j == 0);
    while (m) {
        const int src = __ffs(m) - 1;                      // lane 0 of a contested group
        m &= m - 1;
        const unsigned long long xp = __shfl_sync(VT_FULL, (unsigned long long)xrow, src);
        const unsigned long long cp = __shfl_sync(VT_FULL, (unsigned long long)children, src);
        const int r = canon_argmin_warp<T, DP>(reinterpret_cast<const T *>(xp),
                                               reinterpret_cast<const float *>(cp), k, lane);
        if ((lane >> 3) == (src >> 3)) { arg = r; if (j == 0) ++rechecks; }
    }
    return arg;
}

// ---- packed descent paths (level-synchronous kernels) ---------------------------------------------
// Between levels the batch only has to be GROUPED by node, not ordered by node index.  A row's key is
// its path from the root packed `bits` bits per level, newest child index in the low bits; one stable
// radix pass over the low 8 bits groups the rows by their new node (rows of one node were contiguous
// before the pass and share all digits, so they stay contiguous inside their new digit class).  A row
// that stopped at a leaf of a ragged tree carries the all-ones digit and sorts to the end of its class.
__host__ __device__ __forceinline__ int path_bits(int k)
{
    int b = 1;
    while ((1 << b) - 1 < k) ++b;          // labels 0..k-1 plus the all-ones "stopped" marker
    return b;
}
// same, for trees whose heap has fewer than 2^31 slots (32-bit arithmetic: 3 instructions per level)
__device__ __forceinline__ uint32_t path_to_heap32(uint32_t path, int level, int bits, uint32_t k)
{
    const uint32_t mask = (1u << bits) - 1u;
    uint32_t h = 0;
    int sh = bits * (level - 1);
    for (int i = 0; i < level; ++i, sh -= bits) h = h * k + 1u + ((path >> sh) & mask);
    return h;
}
__device__ __forceinline__ long long path_to_heap(uint32_t path, int level, int bits, int k)
{
    const uint32_t mask = (1u << bits) - 1u;
    long long h = 0;
    for (int i = level - 1; i >= 0; --i) h = h * k + 1 + (long long)((path >> (bits * i)) & mask);
    return h;
}
