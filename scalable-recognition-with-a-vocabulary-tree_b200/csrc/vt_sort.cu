// vt_sort.cu -- stable LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits per pass.
//
// Used for (a) the stable partition of node members at the end of every tree level (the Python
// append loop at cbir/encoders/vocabulary.py:64-66) and (b) turning the forward bag-of-words CSR
// into image-ordered posting lists (the per-node dicts of vocabulary.py:96-100).
// HBM-bound: per pass one read of the keys (histogram) + one read and one scattered write of the
// pairs = 20 B per element per pass.
//
// Per pass:  hist (per-tile digit counts, digit-major)  ->  scan (one block per digit, exclusive
// over tiles)  ->  scatter (per-warp stable ranking with match_any, tile = 8 warps x 512 items, written out
// through a digit-ordered shared-memory tile so that every digit's run leaves as coalesced segments).
// Digits of at most 5 bits -- the regrouping pass of the level-synchronous descent -- take the split_*
// kernels below: same scheme, warps rank with one ballot per digit bit.
#include "vt_common.cuh"

namespace {
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;                          // items per thread
constexpr int SORT_WARP_ITEMS = 32 * SORT_ROUNDS;        // 512 contiguous items per warp
constexpr int SORT_TILE = SORT_WARP_ITEMS * SORT_WARPS;  // 4096 items per block

__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, uint32_t mask, uint32_t *__restrict__ hist,
                 int nblocks)
{
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t start = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = start + r * SORT_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&cnt[(keys[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// one block per digit: exclusive scan of that digit's per-tile counts, in place; total -> digit_total.
// Four consecutive entries per thread and iteration (the block-wide step is the latency of this kernel).
__global__ void __launch_bounds__(SORT_THREADS)
sort_scan_kernel(uint32_t *__restrict__ hist, int nblocks, uint32_t *__restrict__ digit_total)
{
    __shared__ uint32_t wsum[SORT_WARPS];
    uint32_t *row = hist + (size_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int base = 0; base < nblocks; base += 4 * SORT_THREADS) {
        const int idx = base + 4 * threadIdx.x;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = idx + j < nblocks ? row[idx + j] : 0u;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = wsum[w];
            if (w < warp) woff += t;
            total += t;
        }
        uint32_t run = carry + woff + incl - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (idx + j < nblocks) row[idx + j] = run;
            run += v[j];
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                    uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift, uint32_t mask,
                    const uint32_t *__restrict__ hist, const uint32_t *__restrict__ digit_total, int nblocks)
{
    // The tile is first ordered by digit in shared memory and then written out run by run, so that the
    // global stores of a warp cover contiguous addresses (random keys would otherwise scatter 4-byte
    // stores over 256 streams: 8x write amplification at 32 B sector granularity).
    __shared__ uint32_t wcnt[SORT_WARPS][256];        // per-warp digit counts -> per-warp offsets inside the tile
    __shared__ uint32_t dbase[256];                   // global position of this tile's first element of digit d
    __shared__ uint32_t tstart[257];                  // tile-local start of digit d (exclusive scan of the tile counts)
    __shared__ uint32_t s_key[SORT_TILE];
    __shared__ uint32_t s_val[SORT_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) wcnt[w][threadIdx.x] = 0;
    // exclusive scan of the 256 digit totals (warp shuffles + one exchange) -> global digit bases
    {
        __shared__ uint32_t dtot[SORT_WARPS];
        const uint32_t mine = digit_total[threadIdx.x];
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) dtot[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) if (w < warp) woff += dtot[w];
        dbase[threadIdx.x] = woff + incl - mine + (threadIdx.x <= mask ? hist[(size_t)threadIdx.x * nblocks + blockIdx.x] : 0u);
    }
    __syncthreads();

    const int64_t tile0 = (int64_t)blockIdx.x * SORT_TILE;
    const int64_t wstart = tile0 + (int64_t)warp * SORT_WARP_ITEMS;
    uint32_t key[SORT_ROUNDS], val[SORT_ROUNDS], lrank[SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wstart + r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0u;
        val[r] = valid ? vals_in[idx] : 0u;
        const uint32_t d = (key[r] >> shift) & mask;
        const uint32_t tag = valid ? d : (0x100u | (uint32_t)lane);        // invalid lanes match nobody
        const uint32_t peers = __match_any_sync(VT_FULL, tag);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        const uint32_t base = valid ? wcnt[warp][d] : 0u;
        __syncwarp();
        if (valid && rank == 0) wcnt[warp][d] = base + __popc(peers);
        __syncwarp();
        lrank[r] = base + rank;
    }
    __syncthreads();
    {   // thread d: tile count of digit d, per-warp offsets inside the digit's run
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += t;
        }
        tstart[threadIdx.x] = run;                     // tile count for now
    }
    __syncthreads();
    {   // exclusive scan of the tile counts
        const uint32_t cnt = tstart[threadIdx.x];
        __syncthreads();
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        __shared__ uint32_t wtot[SORT_WARPS];
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) if (w < warp) woff += wtot[w];
        tstart[threadIdx.x] = woff + incl - cnt;
        if (threadIdx.x == 255) tstart[256] = woff + incl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wstart + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = (key[r] >> shift) & mask;
            const uint32_t pos = tstart[d] + wcnt[warp][d] + lrank[r];
            s_key[pos] = key[r];
            s_val[pos] = val[r];
        }
    }
    __syncthreads();
    const uint32_t n_tile = tstart[256];
    for (uint32_t i = threadIdx.x; i < n_tile; i += SORT_THREADS) {
        const uint32_t k2 = s_key[i];
        const uint32_t d = (k2 >> shift) & mask;
        const uint32_t pos = dbase[d] + (i - tstart[d]);
        keys_out[pos] = k2;
        vals_out[pos] = s_val[i];
    }
}

// ---- narrow digits (<= 32 classes): the regrouping pass of the level-synchronous descent -------------
// Same hist -> scan -> scatter scheme and the same scratch layout, but a warp ranks its items with one
// ballot per digit bit: lane j derives the member mask of class j from the ballots (its running count
// stays in a register) and every lane the mask of its own class -- no shared-memory counters, no
// intra-warp synchronisation in the ranking loop.
template <int NB>
__global__ void __launch_bounds__(SORT_THREADS)
split_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist, int nblocks)
{
    constexpr int NC = 1 << NB;
    __shared__ uint32_t cnt[NC];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < NC) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t start = (int64_t)blockIdx.x * SORT_TILE;
    uint32_t mine = 0;                                   // lane j: items of class j seen by this warp
    uint32_t kreg[SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {                 // every load of the tile in flight before the first ballot
        const int64_t idx = start + r * SORT_THREADS + threadIdx.x;
        kreg[r] = idx < n ? keys[idx] : 0u;
    }
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = start + r * SORT_THREADS + threadIdx.x;
        const bool valid = idx < n;
        const uint32_t d = (kreg[r] >> shift) & (NC - 1);
        uint32_t cm = __ballot_sync(VT_FULL, valid);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t bb = __ballot_sync(VT_FULL, (d >> b) & 1u);
            cm &= ((lane >> b) & 1) ? bb : ~bb;
        }
        mine += __popc(cm);
    }
    if (lane < NC && mine) atomicAdd(&cnt[lane], mine);
    __syncthreads();
    if (threadIdx.x < NC) hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

template <int NB>
__global__ void __launch_bounds__(SORT_THREADS, 4)
split_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift,
                     const uint32_t *__restrict__ hist, const uint32_t *__restrict__ digit_total, int nblocks)
{
    constexpr int NC = 1 << NB;
    static_assert(NC <= 32, "one lane per class");
    __shared__ uint32_t wcnt[SORT_WARPS][NC];         // per-warp class totals -> offsets inside the class run
    __shared__ uint32_t dbase[NC];                    // global position of this tile's first element of class d
    __shared__ uint32_t tstart[NC + 1];               // tile-local start of class d
    __shared__ uint32_t s_key[SORT_TILE];
    __shared__ uint32_t s_val[SORT_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    if (warp == 0) {                                  // exclusive scan of the class totals + this tile's offsets
        const uint32_t v = lane < NC ? digit_total[lane] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane < NC) dbase[lane] = incl - v + hist[(size_t)lane * nblocks + blockIdx.x];
    }
    const int64_t tile0 = (int64_t)blockIdx.x * SORT_TILE;
    const int64_t wstart = tile0 + (int64_t)warp * SORT_WARP_ITEMS;
    // The ranking runs twice -- once to count the warp's items per class, once (after the tile offsets are known)
    // to place them -- instead of keeping 16 ranks per thread alive across the barrier: ballots are cheap, the
    // registers buy a fourth resident CTA for a kernel that waits on memory most of the time.
    uint32_t key[SORT_ROUNDS];
    uint32_t wc = 0;                                  // lane j: this warp's count of class j
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wstart + r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0u;
    }
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wstart + r * 32 + lane;
        const uint32_t d = (key[r] >> shift) & (NC - 1);
        uint32_t cm = __ballot_sync(VT_FULL, idx < n);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t bb = __ballot_sync(VT_FULL, (d >> b) & 1u);
            cm &= ((lane >> b) & 1) ? bb : ~bb;
        }
        wc += __popc(cm);
    }
    if (lane < NC) wcnt[warp][lane] = wc;
    __syncthreads();
    if (warp == 0) {
        uint32_t run = 0;
        if (lane < NC) {
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) {
                const uint32_t t = wcnt[w][lane];
                wcnt[w][lane] = run;
                run += t;
            }
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(VT_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane < NC) tstart[lane] = incl - run;
        if (lane == NC - 1) tstart[NC] = incl;
    }
    uint32_t val[SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wstart + r * 32 + lane;
        val[r] = idx < n ? vals_in[idx] : 0u;
    }
    __syncthreads();
    wc = lane < NC ? tstart[lane] + wcnt[warp][lane] : 0u;      // lane j: tile position of this warp's next item of class j
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wstart + r * 32 + lane;
        const bool valid = idx < n;
        const uint32_t d = (key[r] >> shift) & (NC - 1);
        const uint32_t vm = __ballot_sync(VT_FULL, valid);
        uint32_t cm = vm, pm = vm;                    // members of class `lane` / of this lane's own class
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t bb = __ballot_sync(VT_FULL, (d >> b) & 1u);
            cm &= ((lane >> b) & 1) ? bb : ~bb;
            pm &= ((d >> b) & 1u) ? bb : ~bb;
        }
        const uint32_t pos = __shfl_sync(VT_FULL, wc, (int)d) + __popc(pm & lt);
        if (valid) { s_key[pos] = key[r]; s_val[pos] = val[r]; }
        wc += __popc(cm);
    }
    __syncthreads();
    const uint32_t n_tile = tstart[NC];
    for (uint32_t i = threadIdx.x; i < n_tile; i += SORT_THREADS) {
        const uint32_t k2 = s_key[i];
        const uint32_t d = (k2 >> shift) & (NC - 1);
        const uint32_t pos = dbase[d] + (i - tstart[d]);
        keys_out[pos] = k2;
        vals_out[pos] = s_val[i];
    }
}

template <int NB>
static int launch_split(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, int64_t n, int shift,
                        uint32_t *hist, uint32_t *digit_total, int nblocks, cudaStream_t s)
{
    split_hist_kernel<NB><<<nblocks, SORT_THREADS, 0, s>>>(kin, n, shift, hist, nblocks);
    VT_LAUNCH_CHECK();
    sort_scan_kernel<<<1 << NB, SORT_THREADS, 0, s>>>(hist, nblocks, digit_total);
    VT_LAUNCH_CHECK();
    split_scatter_kernel<NB><<<nblocks, SORT_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, hist, digit_total, nblocks);
    VT_LAUNCH_CHECK();
    return VT_OK;
}
}  // namespace

extern "C" int64_t vt_sort_scratch_bytes(int64_t n)
{
    const int64_t nblocks = (n + SORT_TILE - 1) / SORT_TILE;
    return (256 * (nblocks > 0 ? nblocks : 1) + 256) * (int64_t)sizeof(uint32_t);
}

int vt_sort_pairs_impl(uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_tmp_keys, uint32_t *d_tmp_vals, int64_t n,
                       int n_bits, void *d_hist, int *result_in_tmp, cudaStream_t s)
{
    VT_CHECK_ARG(n >= 0 && n < ((int64_t)1 << 31), "n must be below 2^31");
    VT_CHECK_ARG(n_bits >= 0 && n_bits <= 32, "n_bits out of range");
    VT_CHECK_ARG(result_in_tmp != nullptr, "result_in_tmp is null");
    *result_in_tmp = 0;
    if (n == 0 || n_bits == 0) return VT_OK;
    VT_CHECK_ARG(d_keys && d_vals && d_tmp_keys && d_tmp_vals && d_hist, "null buffer");
    const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    uint32_t *hist = (uint32_t *)d_hist;
    uint32_t *digit_total = hist + (size_t)256 * nblocks;
    uint32_t *kin = d_keys, *vin = d_vals, *kout = d_tmp_keys, *vout = d_tmp_vals;
    int flips = 0;
    const uint32_t mask = 0xffu;
    for (int shift = 0; shift < n_bits; shift += 8) {
        sort_hist_kernel<<<nblocks, SORT_THREADS, 0, s>>>(kin, n, shift, mask, hist, nblocks);
        VT_LAUNCH_CHECK();
        sort_scan_kernel<<<256, SORT_THREADS, 0, s>>>(hist, nblocks, digit_total);
        VT_LAUNCH_CHECK();
        sort_scatter_kernel<<<nblocks, SORT_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, mask, hist, digit_total, nblocks);
        VT_LAUNCH_CHECK();
        uint32_t *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        ++flips;
    }
    *result_in_tmp = flips & 1;
    return VT_OK;
}

// one stable pass over the digit (key >> shift) & mask, mask <= 0xff (kin/vin -> kout/vout)
int vt_sort_pass_impl(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout, int64_t n, int shift,
                      uint32_t mask, void *d_hist, cudaStream_t s)
{
    VT_CHECK_ARG(n > 0 && n < ((int64_t)1 << 31) && shift >= 0 && shift <= 24 && mask <= 0xffu, "bad arguments");
    VT_CHECK_ARG(kin && vin && kout && vout && d_hist, "null buffer");
    const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    uint32_t *hist = (uint32_t *)d_hist;
    uint32_t *digit_total = hist + (size_t)256 * nblocks;
    if (mask <= 15u) return launch_split<4>(kin, vin, kout, vout, n, shift, hist, digit_total, nblocks, s);
    if (mask <= 31u) return launch_split<5>(kin, vin, kout, vout, n, shift, hist, digit_total, nblocks, s);
    sort_hist_kernel<<<nblocks, SORT_THREADS, 0, s>>>(kin, n, shift, mask, hist, nblocks);
    VT_LAUNCH_CHECK();
    VT_CUDA(cudaMemsetAsync(digit_total, 0, 256 * sizeof(uint32_t), s));
    sort_scan_kernel<<<(int)mask + 1, SORT_THREADS, 0, s>>>(hist, nblocks, digit_total);      // digits above the mask do not occur
    VT_LAUNCH_CHECK();
    sort_scatter_kernel<<<nblocks, SORT_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, mask, hist, digit_total, nblocks);
    VT_LAUNCH_CHECK();
    return VT_OK;
}

extern "C" int vt_sort_pairs(uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_tmp_keys, uint32_t *d_tmp_vals, int64_t n,
                             int n_bits, void *d_hist, int *result_in_tmp, void *stream)
{
    return vt_sort_pairs_impl(d_keys, d_vals, d_tmp_keys, d_tmp_vals, n, n_bits, d_hist, result_in_tmp,
                              (cudaStream_t)stream);
}
