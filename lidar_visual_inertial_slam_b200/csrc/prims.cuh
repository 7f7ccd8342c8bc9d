// prims.cuh -- own device-wide primitives: exclusive scan and a stable LSD radix sort of
// (uint32 key, uint32 payload) pairs.  Hand-written (no CUB/Thrust) so that every timed kernel
// of the VoxelGrid replacement (pcl::VoxelGrid's std::sort / integer_sort)
// and of the search-grid build is ours.  Stability (ties keep input order) is what makes the
// within-voxel summation order, hence the fp32 centroids, reproducible.
#pragma once

#include "common.cuh"

namespace lvreg {

// ------------------------------------------------------------------------------------------
// exclusive scan over uint32 values produced by a functor, consumed by a functor
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive prefix of `v` over the block (kScanThreads threads); also returns the block total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        uint32_t winc = warp_inclusive_scan(w, lane);
        if (lane < kScanThreads / 32) warp_sums[lane] = winc - w;
        if (lane == kScanThreads / 32 - 1) block_total = winc;
    }
    __syncthreads();
    uint32_t res = inc - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return res;
}

// Reduce-then-scan in three launches (tile sums, one-block scan of the sums, apply).  A single-pass
// look-back scan was measured SLOWER here: the tiles are short, so the look-back latency of the
// ~1000 tiles in flight dominates.  Loads are 2 x 128-bit per thread where the functor allows.
template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(In in, uint32_t n,
                                                                      uint32_t* tile_sums) {
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t s = 0;
    if (base + kScanItems <= n) {
        uint32_t v[kScanItems];
        in.load_vec(base, v);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) s += v[k];
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            uint32_t i = base + k;
            if (i < n) s += in(i);
        }
    }
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of data[0..n) in place, grand total to *total_out
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t* data, uint32_t n,
                                                         uint32_t* total_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? data[i] : 0;
        uint32_t inc = warp_inclusive_scan(v, lane);
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            uint32_t winc = warp_inclusive_scan(w, lane);
            warp_sums[lane] = winc - w;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t exc = inc - v + warp_sums[warp] + carry;
        if (i < n) data[i] = exc;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = exc + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

template <class In, class Out>
__global__ void __launch_bounds__(kScanThreads) scan_tile_apply_kernel(In in, uint32_t n,
                                                                       const uint32_t* tile_offsets,
                                                                       Out out) {
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
    const bool full = base + kScanItems <= n;
    if (full) {
        in.load_vec(base, v);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) s += v[k];
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            uint32_t i = base + k;
            v[k] = i < n ? in(i) : 0;
            s += v[k];
        }
    }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(s, &total) + tile_offsets[blockIdx.x];
    if (full) {
        out.store_vec(base, v, pre);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            uint32_t i = base + k;
            if (i < n) out(i, v[k], pre);
            pre += v[k];
        }
    }
}

inline uint32_t scan_num_tiles(uint32_t n) { return (n + kScanTile - 1) / kScanTile; }

// temp: scan_num_tiles(n) uint32.  total_out: device pointer receiving the grand total.
template <class In, class Out>
inline void exclusive_scan(In in, Out out, uint32_t n, uint32_t* temp, uint32_t* total_out,
                           cudaStream_t st, int* launches) {
    if (n == 0) {
        cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st);
        return;
    }
    const uint32_t tiles = scan_num_tiles(n);
    scan_tile_sums_kernel<<<tiles, kScanThreads, 0, st>>>(in, n, temp);
    scan_sums_kernel<<<1, 1024, 0, st>>>(temp, tiles, total_out);
    scan_tile_apply_kernel<<<tiles, kScanThreads, 0, st>>>(in, n, temp, out);
    if (launches) *launches += 3;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass, one kernel per pass ("onesweep" organisation):
//   rs_global_hist_kernel   one read of the keys -> global digit histograms of ALL passes
//   rs_onesweep_kernel      per pass: tile ranking (warp match), tile offsets resolved by
//                           decoupled look-back over the preceding tiles (no separate scan
//                           kernel, no second read of the keys), scatter staged through shared
//                           memory so that every digit run leaves the SM as one coalesced burst
// HBM traffic per pass: 8 B read + 8 B written per pair (+ 4 B per pair once for the histograms).
// Tiles take their index from an atomic ticket, so a tile only ever waits on tiles that are
// already running; within a tile the element order is (warp, round, lane) == index order, which
// makes the sort stable.
// ------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;           // tile of 4096 pairs: inputs up to kSortBigN (shorter tiles, lower latency)
constexpr int kSortThreadsBig = 512;        // tile of 8192 pairs: large inputs (half as many tiles in the look-back chain)
constexpr uint32_t kSortBigN = 600000;
constexpr int kSortItems = 16;
constexpr int kSortMaxPasses = 4;
constexpr size_t sort_smem_bytes(int threads) { return (size_t)threads * kSortItems * 8 + (size_t)(threads / 32 + 2) * 256 * 4 + 128; }
constexpr uint32_t kFlagAgg = 1u << 30, kFlagIncl = 2u << 30, kFlagMask = 3u << 30;

__global__ void __launch_bounds__(256) rs_global_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                             int passes, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t hist[8][kSortMaxPasses][256];       // one copy per warp: conflicts stay intra-warp
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * kSortMaxPasses * 256; i += 256) (&hist[0][0][0])[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t k = keys[i];
#pragma unroll
        for (int p = 0; p < kSortMaxPasses; ++p)
            if (p < passes) atomicAdd(&hist[warp][p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += 256) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += (&hist[w][0][0])[i];
        if (s) atomicAdd(&ghist[i], s);
    }
}

#ifdef LVREG_SORT_PROF
// debug build only: per-phase clock cycles summed over tiles (thread 0 of every block)
__device__ unsigned long long g_sort_prof[16];
#define SORT_STAMP(k) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_sort_prof[k], (unsigned long long)(t_ - t_prev_)); t_prev_ = t_; } } while (0)
#else
#define SORT_STAMP(k) do {} while (0)
#endif

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) rs_onesweep_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
    const uint32_t* __restrict__ ghist /*256, this pass*/, volatile uint32_t* status /*[tiles][256]*/,
    uint32_t* ticket) {
    // dynamic shared memory (more than 48 KB with 512 threads): staging | per-warp counters | small arrays
    constexpr int TILE = THREADS * kSortItems, WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char sort_smem[];
    uint2* skv = reinterpret_cast<uint2*>(sort_smem);            // [TILE] (key, payload): one 64-bit access each way
    uint32_t (*wcnt)[256] = reinterpret_cast<uint32_t (*)[256]>(sort_smem + (size_t)TILE * 8);   // [WARPS][256]
    uint32_t* thist = &wcnt[WARPS][0];  // [256] the tile's digit counts, known before the ranking
    uint32_t* gofs = thist + 256;            // [256] global offset of a digit run minus its tile-local start
    uint32_t (*scan_ws)[8] = reinterpret_cast<uint32_t (*)[8]>(gofs + 256);   // [2][8]
    uint32_t& tile_s = *(gofs + 256 + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef LVREG_SORT_PROF
    long long t_prev_ = clock64();
#endif
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) (&wcnt[0][0])[i] = 0;
    if (threadIdx.x < 256) thist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = tile_s;
    SORT_STAMP(0);

    const uint32_t base = tile * TILE + warp * (32 * kSortItems);
    uint32_t k[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        k[r] = i < n ? keys_in[i] : 0xffffffffu;        // padding sorts last inside the (final) tile
    }
#ifdef LVREG_SORT_PROF
    if (k[0] == 0x12345678u && k[kSortItems - 1] == 0x9abcdef0u) t_prev_ += 1;   // wait for the loads
    SORT_STAMP(1);
#endif

    // ---- the tile's digit counts first: they are all the other tiles wait for.  Publishing them (and
    // resolving this tile's own offsets) BEFORE the ranking takes the ranking out of the chain of
    // dependent tiles: a predecessor's count is ~5k cycles away from its start instead of ~15k, and the
    // look-back below rarely meets a tile that has not published yet. ----
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) atomicAdd(&thist[(k[r] >> shift) & 255u], 1u);
    __syncthreads();
    SORT_STAMP(2);
    const int d = threadIdx.x & 255;         // the first 256 threads own one digit each
    const bool digit_thread = threadIdx.x < 256;
    const uint32_t cnt = thist[d];
    // the padding of the last tile was counted under digit 255: remove it from the published count
    uint32_t pad = 0;
    if (d == 255) {
        const uint32_t tile_end = (tile + 1) * TILE;
        pad = tile_end > n ? tile_end - n : 0;
    }
    const uint32_t real_cnt = cnt - pad;
    if (digit_thread) status[(size_t)tile * 256 + d] = (tile == 0 ? kFlagIncl : kFlagAgg) | real_cnt;
    uint32_t excl = 0;
    if (digit_thread && tile > 0) {
        // decoupled look-back, 16 predecessors per round trip (independent loads in flight).  Measured at
        // 11.4 M pairs: 2.6 round trips per tile, a third of them meet a predecessor that has not published
        // yet; wider batches (32, 48) were slower -- the cost of a round trip grows with the loads issued.
        constexpr int LB = 16;
        int pred = (int)tile - 1;
        bool done = false;
        while (!done) {
            uint32_t s[LB];
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int idx = pred - j;
                s[j] = 0x80000000u;                                            // before tile 0: inclusive prefix 0
                if (idx >= 0) s[j] = status[(size_t)idx * 256 + d];
            }
            // common case, branch-free: everything up to the first inclusive prefix is published
            uint32_t incl_mask = 0, miss_mask = 0;
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                incl_mask |= ((s[j] >> 31) & 1u) << j;
                miss_mask |= ((s[j] & kFlagMask) == 0 ? 1u : 0u) << j;
            }
            const int first = incl_mask ? __ffs(incl_mask) - 1 : LB;          // LB: no inclusive prefix in this batch
            const uint32_t upto = first < LB ? ((2u << first) - 1u) : ((LB == 32) ? 0xffffffffu : ((1u << LB) - 1u));
#ifdef LVREG_SORT_PROF
            if (threadIdx.x == 0) { atomicAdd(&g_sort_prof[10], 1ull); if (miss_mask & upto) atomicAdd(&g_sort_prof[11], 1ull); }
            if (threadIdx.x == 0 && first < LB) atomicAdd(&g_sort_prof[12], (unsigned long long)first);
#endif
            if ((miss_mask & upto) == 0) {
#pragma unroll
                for (int j = 0; j < LB; ++j)
                    if ((upto >> j) & 1u) excl += s[j] & ~kFlagMask;
                done = first < LB;
            } else {
#pragma unroll
                for (int j = 0; j < LB; ++j) {
                    if (!done) {
                        uint32_t sv = s[j];
                        while ((sv & kFlagMask) == 0) sv = status[(size_t)(pred - j) * 256 + d];
                        excl += sv & ~kFlagMask;
                        if (sv & kFlagIncl) done = true;
                    }
                }
            }
            pred -= LB;
        }
        status[(size_t)tile * 256 + d] = kFlagIncl | (excl + real_cnt);
    }
    SORT_STAMP(3);

    // ---- rank the tile's elements (stable) ----
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        // lanes with the same digit, from 8 ballots: the hardware match.any iterates over the
        // distinct values in the warp (~30 here) and measured ~2x slower under load
        const uint32_t dg = (k[r] >> shift) & 255u;
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (dg & (1u << b)) != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bal ^ (bit ? 0u : 0xffffffffu);
        }
        // The counters are private to the warp.  Every lane reads its digit's counter (peers read the same
        // word), then the lowest peer alone stores the new value: no atomic, no branch, no shuffle.  Shared
        // memory accesses of one warp execute in program order, so the next round sees the store.
        const uint32_t lower = peers & lt_mask;
        const uint32_t old = wcnt[warp][dg];
        if (lower == 0) wcnt[warp][dg] = old + (uint32_t)__popc(peers);
        __syncwarp();
        rank[r] = (uint16_t)(old + __popc(lower));
    }
    SORT_STAMP(4);

    // the payloads are not needed before the staging: load them now, off the ranking's registers
    uint32_t v[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        v[r] = i < n ? vals_in[i] : 0u;
    }

    // exclusive scan of the global histogram (digit bases) and of the tile counts (tile-local starts)
    uint32_t gb, tl;
    {
        const uint32_t gh = ghist[d];
        const uint32_t a = warp_inclusive_scan(gh, lane);
        const uint32_t b = warp_inclusive_scan(cnt, lane);
        if (lane == 31 && digit_thread) { scan_ws[0][warp] = a; scan_ws[1][warp] = b; }
        __syncthreads();                      // also: every warp's counters are final
        uint32_t wa = 0, wb = 0;
        for (int w = 0; w < (warp & 7); ++w) { wa += scan_ws[0][w]; wb += scan_ws[1][w]; }
        gb = a - gh + wa;
        tl = b - cnt + wb;
    }
    // tile-local start of each (warp, digit) run: exclusive prefix over the warps + the digit's start
    if (digit_thread) {
        gofs[d] = gb + excl - tl;
        uint32_t run = tl;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    SORT_STAMP(5);

    // ---- stage in shared memory in tile-sorted order, then write coalesced runs ----
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t dd = (k[r] >> shift) & 255u;
        const uint32_t pos = wcnt[warp][dd] + rank[r];
        skv[pos] = make_uint2(k[r], v[r]);
    }
    __syncthreads();
    SORT_STAMP(6);
    const uint32_t tile_n = min((uint32_t)TILE, n - tile * TILE);
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t e = r * THREADS + threadIdx.x;
        if (e < tile_n) {
            const uint2 kv = skv[e];
            const uint32_t dst = gofs[(kv.x >> shift) & 255u] + e;
            keys_out[dst] = kv.x;
            vals_out[dst] = kv.y;
        }
    }
    SORT_STAMP(7);
}

inline int sort_threads_for(uint32_t n) { return n > kSortBigN ? kSortThreadsBig : kSortThreads; }
inline uint32_t sort_num_blocks(uint32_t n) {
    const uint32_t tile = (uint32_t)sort_threads_for(n) * kSortItems;
    return (n + tile - 1) / tile;
}
// scratch needed by radix_sort_pairs, in uint32 words: look-back status per pass + histograms + tickets
// (sized for any input of up to n pairs: just below kSortBigN the tiles are shorter, hence more of them)
inline size_t sort_scratch_words(uint32_t n) {
    uint32_t nb = sort_num_blocks(n);
    const uint32_t small = n < kSortBigN ? n : kSortBigN;
    if (sort_num_blocks(small) > nb) nb = sort_num_blocks(small);
    return (size_t)kSortMaxPasses * 256 * nb + kSortMaxPasses * 256 + 8;
}

inline int sort_num_passes(int key_bits) {
    int passes = key_bits <= 0 ? 1 : (key_bits + 7) / 8;
    return passes > kSortMaxPasses ? kSortMaxPasses : passes;
}
inline uint32_t* sort_ghist_ptr(uint32_t* scratch, uint32_t n) {
    return scratch + (size_t)kSortMaxPasses * 256 * sort_num_blocks(n);
}

// zeroes the look-back status words, the digit histograms and the tickets; must precede the
// histogram producer (rs_global_hist_kernel, or a key-generating kernel that accumulates the
// histograms itself through HistAccumulator)
inline void radix_sort_prepare(uint32_t* scratch, uint32_t n, int key_bits, cudaStream_t st) {
    const uint32_t nblocks = sort_num_blocks(n);
    const int passes = sort_num_passes(key_bits);
    cudaMemsetAsync(scratch, 0, ((size_t)passes * 256 * nblocks) * sizeof(uint32_t), st);
    cudaMemsetAsync(sort_ghist_ptr(scratch, n), 0, (kSortMaxPasses * 256 + 8) * sizeof(uint32_t), st);
}

// per-block digit histograms of all passes, accumulated by a kernel that already has the keys in
// registers (saves the separate histogram read of the keys)
struct HistAccumulator {
    uint32_t (*hist)[kSortMaxPasses][256];     // shared memory, one copy per warp
    __device__ __forceinline__ void init(uint32_t (*smem)[kSortMaxPasses][256]) {
        hist = smem;
        for (int i = threadIdx.x; i < 8 * kSortMaxPasses * 256; i += blockDim.x) (&hist[0][0][0])[i] = 0;
        __syncthreads();
    }
    __device__ __forceinline__ void add(uint32_t key, int passes) {
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int p = 0; p < kSortMaxPasses; ++p)
            if (p < passes) atomicAdd(&hist[warp][p][(key >> (8 * p)) & 255u], 1u);
    }
    __device__ __forceinline__ void flush(uint32_t* ghist, int passes) {
        __syncthreads();
        for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
            uint32_t s = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += (&hist[w][0][0])[i];
            if (s) atomicAdd(&ghist[i], s);
        }
    }
};

// the passes proper; `hist_ready` = the histograms were accumulated by the key producer
inline int radix_sort_run(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, uint32_t n,
                          int key_bits, uint32_t* scratch, bool hist_ready, cudaStream_t st, int* launches) {
    if (n == 0) return 0;
    const uint32_t nblocks = sort_num_blocks(n);
    const int passes = sort_num_passes(key_bits);
    uint32_t* status = scratch;
    uint32_t* ghist = sort_ghist_ptr(scratch, n);
    uint32_t* tickets = ghist + kSortMaxPasses * 256;
    if (!hist_ready) {
        uint32_t hb = (n + 256 * 16 - 1) / (256 * 16);
        if (hb > (uint32_t)kNumSMs * 8) hb = kNumSMs * 8;
        rs_global_hist_kernel<<<hb, 256, 0, st>>>(keys0, n, passes, ghist);
        if (launches) *launches += 1;
    }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const uint32_t* kin = cur ? keys1 : keys0;
        const uint32_t* vin = cur ? vals1 : vals0;
        uint32_t* kout = cur ? keys0 : keys1;
        uint32_t* vout = cur ? vals0 : vals1;
        if (sort_threads_for(n) == kSortThreadsBig)
            rs_onesweep_kernel<kSortThreadsBig><<<nblocks, kSortThreadsBig, sort_smem_bytes(kSortThreadsBig), st>>>(
                kin, vin, kout, vout, n, 8 * p, ghist + 256 * p, status + (size_t)p * 256 * nblocks, tickets + p);
        else
            rs_onesweep_kernel<kSortThreads><<<nblocks, kSortThreads, sort_smem_bytes(kSortThreads), st>>>(
                kin, vin, kout, vout, n, 8 * p, ghist + 256 * p, status + (size_t)p * 256 * nblocks, tickets + p);
        cur ^= 1;
    }
    if (launches) *launches += passes;
    return cur;
}

// Sorts (keys, vals) by the low `key_bits` bits of key, stable.  Ping-pongs between the two
// buffer pairs; returns 0 when the result is in (keys0, vals0), 1 when in (keys1, vals1).
inline int radix_sort_pairs(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1,
                            uint32_t n, int key_bits, uint32_t* scratch, cudaStream_t st,
                            int* launches) {
    if (n == 0) return 0;
    radix_sort_prepare(scratch, n, key_bits, st);
    return radix_sort_run(keys0, vals0, keys1, vals1, n, key_bits, scratch, false, st, launches);
}

}  // namespace lvreg
