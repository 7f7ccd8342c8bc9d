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

template <class In>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(In in, uint32_t n,
                                                                      uint32_t* tile_sums) {
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint32_t i = base + k;
        if (i < n) s += in(i);
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
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint32_t i = base + k;
        v[k] = i < n ? in(i) : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        uint32_t i = base + k;
        if (i < n) out(i, v[k], pre);
        pre += v[k];
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
// stable LSD radix sort, 8 bits per pass
// ------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096 pairs per block
constexpr int kSortWarps = kSortThreads / 32;

// per-tile digit histogram -> table[digit * nblocks + block]; digit totals -> totals[digit]
__global__ void __launch_bounds__(kSortThreads) rs_hist_kernel(const uint32_t* __restrict__ keys,
                                                               uint32_t n, int shift,
                                                               uint32_t* __restrict__ table,
                                                               uint32_t* __restrict__ totals,
                                                               uint32_t nblocks) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll 4
    for (int k = 0; k < kSortItems; ++k) {
        uint32_t i = base + k * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    uint32_t c = hist[threadIdx.x];
    table[threadIdx.x * nblocks + blockIdx.x] = c;
    if (c) atomicAdd(&totals[threadIdx.x], c);
}

// one block per digit: table row -> exclusive global offsets
__global__ void __launch_bounds__(256) rs_scan_kernel(uint32_t* __restrict__ table,
                                                      const uint32_t* __restrict__ totals,
                                                      uint32_t nblocks) {
    __shared__ uint32_t red[256];
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t carry_s;
    const int d = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // base offset of this digit = sum of totals of all smaller digits
    red[threadIdx.x] = (int)threadIdx.x < d ? totals[threadIdx.x] : 0;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) carry_s = red[0];
    __syncthreads();
    uint32_t* row = table + (size_t)d * nblocks;
    for (uint32_t base = 0; base < nblocks; base += 256) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? row[i] : 0;
        uint32_t inc = warp_inclusive_scan(v, lane);
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < 8 ? warp_sums[lane] : 0;
            uint32_t winc = warp_inclusive_scan(w, lane);
            if (lane < 8) warp_sums[lane] = winc - w;
        }
        __syncthreads();
        uint32_t exc = inc - v + warp_sums[warp] + carry_s;
        if (i < nblocks) row[i] = exc;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = exc + v;
        __syncthreads();
    }
}

// stable scatter: element order inside a tile is (warp, round, lane) == index order
__global__ void __launch_bounds__(kSortThreads) rs_scatter_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
    const uint32_t* __restrict__ table, uint32_t n, int shift, uint32_t nblocks) {
    __shared__ uint32_t wcnt[kSortWarps][256];
    __shared__ uint32_t gbase[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
    gbase[threadIdx.x] = table[threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();

    const uint32_t base = blockIdx.x * kSortTile + warp * (32 * kSortItems);
    uint32_t k[kSortItems], v[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        const bool valid = i < n;
        k[r] = valid ? keys_in[i] : 0xffffffffu;
        v[r] = valid ? vals_in[i] : 0u;
        const uint32_t d = (k[r] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = wcnt[warp][d];
            wcnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    {   // per-digit exclusive prefix over the warps of this block
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            uint32_t c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t i = base + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t pos = gbase[d] + wcnt[warp][d] + rank[r];
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
    }
}

inline uint32_t sort_num_blocks(uint32_t n) { return (n + kSortTile - 1) / kSortTile; }
// scratch needed by radix_sort_pairs, in uint32 words
inline size_t sort_scratch_words(uint32_t n) { return (size_t)256 * sort_num_blocks(n) + 256; }

// Sorts (keys, vals) by the low `key_bits` bits of key, stable.  Ping-pongs between the two
// buffer pairs; returns 0 when the result is in (keys0, vals0), 1 when in (keys1, vals1).
inline int radix_sort_pairs(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1,
                            uint32_t n, int key_bits, uint32_t* scratch, cudaStream_t st,
                            int* launches) {
    if (n == 0) return 0;
    const uint32_t nblocks = sort_num_blocks(n);
    uint32_t* table = scratch;
    uint32_t* totals = scratch + (size_t)256 * nblocks;
    int cur = 0;
    const int passes = key_bits <= 0 ? 1 : (key_bits + 7) / 8;
    for (int p = 0; p < passes; ++p) {
        const uint32_t* kin = cur ? keys1 : keys0;
        const uint32_t* vin = cur ? vals1 : vals0;
        uint32_t* kout = cur ? keys0 : keys1;
        uint32_t* vout = cur ? vals0 : vals1;
        cudaMemsetAsync(totals, 0, 256 * sizeof(uint32_t), st);
        rs_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, n, 8 * p, table, totals, nblocks);
        rs_scan_kernel<<<256, 256, 0, st>>>(table, totals, nblocks);
        rs_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(kin, vin, kout, vout, table, n, 8 * p, nblocks);
        if (launches) *launches += 3;
        cur ^= 1;
    }
    return cur;
}

}  // namespace lvreg
