// voxelgrid_bucket.cuh -- pcl::VoxelGrid of a large local map (extractCloud, MO:931-973) without a device-wide sort.
//
// The cached world-frame keyframe clouds (laserCloudMapContainer, MO:942-954) are kept ORDERED by their voxel index
// (a stable per-keyframe sort, done once when the cloud enters the cache: points of one voxel keep their scan order).
// PCL's output order -- ascending idx = i + dx (j + dy k), relative to the bounds of the whole map -- is the
// lexicographic order of the absolute voxel coordinates (k, j, i), so every cached cloud is a sorted run under the key
// of ANY map it takes part in.  The filter over R such runs is then a sample sort whose buckets never leave the SM:
//
//   vgb_sample_kernel   every kVgbSample-th point's key (regular sampling of sorted runs)
//   (radix sort of the samples: ~n / 64 keys)
//   vgb_split_kernel    every kVgbStride-th sorted sample is a splitter; lower bound of every splitter in every run
//                       -> the (bucket, run) segment table
//   vgb_bucket_kernel   one block per bucket (key range): gathers its <= 4096 points as R short sorted runs, sorts
//                       (local key, payload) in shared memory -- stable LSD radix, as many <= 9-bit passes as the key RANGE
//                       of the bucket needs, usually 2 --, finds the voxel starts, sums the voxels sequentially
//                       in (keyframe, scan) order = the order the stable device-wide sort gave, and writes the
//                       centroids into the bucket's own slice of a scratch cloud.
//   vgb_scan_kernel     voxel counts of the buckets -> output offsets (+ the total the host waits for)
//   vgb_compact_kernel  the slices, closed up into the down-sampled map (a chained scan inside the bucket kernel was
//                       measured first: its polling warp cost 8 % of the instructions and stalled seven others)
//
// HBM traffic: the points are read once (16 B, + the second touch from L2 for the sums), the centroids written once;
// against 4 sort passes x 16 B + keys + histograms + heads + gathers before.  Results are bit-identical to the
// sort-based path (same keys, same summation order).  A bucket that does not fit (more than `cap` points: pathological
// sampling, or one voxel with thousands of points) raises a flag and the caller redoes the job on the sort-based path.
#pragma once

#include "voxelgrid.cuh"

namespace lvreg {

constexpr int kVgbSample = 64;                 // S: one sample per 64 input points
constexpr int kVgbStride = 32;                 // t: one splitter per 32 sorted samples (nominal bucket: 2048 points)
constexpr int kVgbThreads = 256;               // 4 independent blocks per SM: their barrier bubbles overlap
constexpr int kVgbItems = 16;
constexpr int kVgbCap = kVgbThreads * kVgbItems;     // 4096
constexpr int kVgbMaxSegs = 1 << (32 - kSegShift);   // 1024

__device__ __forceinline__ uint32_t voxel_key(const float4 p, const VoxelSpec& vs) {
    const int ix = (int)(floorf(p.x * vs.inv) - (float)vs.min_b[0]);
    const int iy = (int)(floorf(p.y * vs.inv) - (float)vs.min_b[1]);
    const int iz = (int)(floorf(p.z * vs.inv) - (float)vs.min_b[2]);
    return (uint32_t)(ix * vs.mul[0] + iy * vs.mul[1] + iz * vs.mul[2]);
}

// The cached clouds carry their voxel coordinates (relative to the keyframe's own bounds) in 4 bytes per point, so the
// passes that only need keys -- the samples and the splitter search, which probes every 10th point of the input --
// read a quarter of the bytes and do integer arithmetic only (split kernel: 45 -> 15 us at C3).  key under the map's bounds = lin(packed) + a constant of the run.
constexpr int kVgbPackX = 11, kVgbPackY = 11, kVgbPackZ = 10;
__device__ __forceinline__ uint32_t pack_ijk(int i, int j, int k) {
    return (uint32_t)i | ((uint32_t)j << kVgbPackX) | ((uint32_t)k << (kVgbPackX + kVgbPackY));
}
__device__ __forceinline__ uint32_t packed_lin(uint32_t w, const VoxelSpec& vs) {
    return (w & ((1u << kVgbPackX) - 1u)) * (uint32_t)vs.mul[0] + ((w >> kVgbPackX) & ((1u << kVgbPackY) - 1u)) * (uint32_t)vs.mul[1] +
           (w >> (kVgbPackX + kVgbPackY)) * (uint32_t)vs.mul[2];
}
__device__ __forceinline__ uint32_t run_key_offset(const Segment& sg, const VoxelSpec& vs) {
    return (uint32_t)((sg.kminb[0] - vs.min_b[0]) * vs.mul[0] + (sg.kminb[1] - vs.min_b[1]) * vs.mul[1] +
                      (sg.kminb[2] - vs.min_b[2]) * vs.mul[2]);
}

// ---- cache fill: a keyframe cloud under its pose, ordered by voxel index ------------------------------------
__global__ void __launch_bounds__(256) bbox_tf_kernel(const float4* __restrict__ in, uint32_t n, Affine T,
                                                      uint32_t* __restrict__ mm) {
    float mnx = 3.4e38f, mny = 3.4e38f, mnz = 3.4e38f, mxx = -3.4e38f, mxy = -3.4e38f, mxz = -3.4e38f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float4 p = in[i];
        const float3 q = apply_affine(T, p.x, p.y, p.z);
        mnx = fminf(mnx, q.x); mny = fminf(mny, q.y); mnz = fminf(mnz, q.z);
        mxx = fmaxf(mxx, q.x); mxy = fmaxf(mxy, q.y); mxz = fmaxf(mxz, q.z);
    }
    block_minmax_commit(mnx, mny, mnz, mxx, mxy, mxz, mm);
}

__global__ void __launch_bounds__(256) voxel_keys_tf_kernel(const float4* __restrict__ in, uint32_t n, Affine T,
                                                            VoxelSpec vs, uint32_t* __restrict__ keys,
                                                            uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    const float3 q = apply_affine(T, p.x, p.y, p.z);
    keys[i] = voxel_key(make_float4(q.x, q.y, q.z, p.w), vs);
    vals[i] = i;
}

__global__ void __launch_bounds__(256) gather_tf_kernel(const float4* __restrict__ in, const uint32_t* __restrict__ order,
                                                        uint32_t n, Affine T, VoxelSpec vs, float4* __restrict__ out,
                                                        uint32_t* __restrict__ wkey) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[order[i]];
    const float3 q = apply_affine(T, p.x, p.y, p.z);
    out[i] = make_float4(q.x, q.y, q.z, p.w);
    if (wkey)
        wkey[i] = pack_ijk((int)(floorf(q.x * vs.inv) - (float)vs.min_b[0]), (int)(floorf(q.y * vs.inv) - (float)vs.min_b[1]),
                       (int)(floorf(q.z * vs.inv) - (float)vs.min_b[2]));
}

// ---- splitters ---------------------------------------------------------------------------------------------
// Sample q = the point at position q * kVgbSample of the concatenation.  Its key goes to `raw` in input order (the
// split kernel searches the samples of a run before it touches the run) and to the array that is sorted; the digit
// histograms of the sort are accumulated here.  `trunc_shift` drops low key bits from the sorted copy (one sort pass
// less for keys of more than 24 bits): the host passes 0 -- with truncated keys a block of 16 consecutive voxel indices
// cannot be split, and at C3 such a block holds up to 4 k points, which overflowed the buckets.
__global__ void __launch_bounds__(256) vgb_sample_kernel(const Segment* __restrict__ segs, uint32_t nseg, uint32_t nsamp,
                                                         VoxelSpec vs, int trunc_shift, int passes,
                                                         uint32_t* __restrict__ raw, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t hist_s[8][kSortMaxPasses][256];
    HistAccumulator acc;
    acc.init(hist_s);
    for (uint32_t q = blockIdx.x * 256 + threadIdx.x; q < nsamp; q += gridDim.x * 256) {
        const uint32_t g = q * (uint32_t)kVgbSample;
        uint32_t lo = 0, hi = nseg;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (segs[mid].begin <= g) lo = mid; else hi = mid;
        }
        const uint32_t key = segs[lo].wkey ? packed_lin(__ldg(segs[lo].wkey + (g - segs[lo].begin)), vs) + run_key_offset(segs[lo], vs)
                                           : voxel_key(__ldg(segs[lo].src + (g - segs[lo].begin)), vs);
        raw[q] = key;
        keys[q] = key >> trunc_shift;
        vals[q] = q;
        acc.add(key >> trunc_shift, passes);
    }
    acc.flush(ghist, passes);
}

__device__ __forceinline__ uint32_t vgb_splitter(const uint32_t* __restrict__ sorted_samples, uint32_t b, int trunc_shift) {
    return sorted_samples[b * (uint32_t)kVgbStride] << trunc_shift;
}

// soff[b * nseg + r] = first position of run r whose key is >= splitter b (b = 0: 0, b = nbuckets: the run length).
// Two levels: the run's own samples (a compact, cache-resident array) narrow the answer down to one window of
// kVgbSample points, which is then searched in the run.
__global__ void __launch_bounds__(256) vgb_split_kernel(const Segment* __restrict__ segs, uint32_t nseg,
                                                        const uint32_t* __restrict__ sorted_samples,
                                                        const uint32_t* __restrict__ raw_samples, uint32_t nbuckets,
                                                        int trunc_shift, VoxelSpec vs, uint32_t* __restrict__ soff) {
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    const uint32_t per = nbuckets + 1;
    if (t >= per * nseg) return;
    const uint32_t r = t / per, b = t - r * per;       // a warp searches ONE run for consecutive splitters
    const uint32_t len = segs[r].n;
    uint32_t res;
    if (b == 0) res = 0;
    else if (b == nbuckets) res = len;
    else {
        const uint32_t sp = vgb_splitter(sorted_samples, b, trunc_shift);
        const uint32_t begin = segs[r].begin;
        // samples inside the run: q in [q0, q1), at run positions q * S - begin
        const uint32_t q0 = (begin + kVgbSample - 1) / kVgbSample, q1 = (begin + len + kVgbSample - 1) / kVgbSample;
        uint32_t a = q0, z = q1;                        // first sample in [q0, q1] with key >= sp
        while (a < z) {
            const uint32_t mid = (a + z) >> 1;
            if (raw_samples[mid] < sp) a = mid + 1; else z = mid;
        }
        // the answer lies in (position of sample a-1, position of sample a]
        uint32_t lo = a > q0 ? (a - 1) * kVgbSample - begin + 1 : 0;
        uint32_t hi = a < q1 ? a * kVgbSample - begin : len;
        const uint32_t* wk = segs[r].wkey;
        const uint32_t koff = run_key_offset(segs[r], vs);
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t km = wk ? packed_lin(__ldg(wk + mid), vs) + koff : voxel_key(__ldg(segs[r].src + mid), vs);
            if (km < sp) lo = mid + 1; else hi = mid;
        }
        res = lo;
    }
    soff[(size_t)b * nseg + r] = res;
}

// ---- buckets -----------------------------------------------------------------------------------------------
struct VgbArgs {
    const Segment* segs;
    uint32_t nseg;
    uint32_t nbuckets;
    uint32_t cap;                      // <= kVgbCap (smaller only to exercise the overflow path)
    uint32_t key_end;                  // one past the largest possible key (end of the last bucket's range)
    int trunc_shift;
    VoxelSpec vs;
    const uint32_t* sorted_samples;
    const uint32_t* soff;              // [(nbuckets + 1)][nseg]
    float4* tmp;                       // [n]: bucket b's centroids start at its input offset (points of the buckets before it)
    uint32_t* bucket_nvox;             // [nbuckets + 1]: voxels per bucket (the scan kernel turns it into output offsets)
    uint32_t* bucket_in;               // [nbuckets]: input offset of every bucket
    uint32_t* info;                    // [0] voxels in total, [1] largest bucket population seen (zeroed)
};

// shared memory: (key, payload) pairs of the bucket -- later the points of half a bucket -- | per-warp digit counters
// (uint16 x 512 digits; before the sort: the run id at every run start, after it: the voxel starts) | segment tables
constexpr int kVgbDigitBits = 9;                // 512 digits = one packed counter word per thread
constexpr int kVgbDigits = 1 << kVgbDigitBits;
constexpr size_t vgb_smem_bytes(uint32_t nseg) {
    return (size_t)kVgbCap * 8 + (size_t)(kVgbThreads / 32) * kVgbDigits * 2 + (size_t)(nseg + 1) * 4 + (size_t)nseg * 4 +
           (size_t)nseg * 8 + 64;
}

__global__ void __launch_bounds__(kVgbThreads, 4) vgb_bucket_kernel(VgbArgs a) {
    constexpr int WARPS = kVgbThreads / 32;
    constexpr int HALF = kVgbCap / 2;
    extern __shared__ __align__(16) unsigned char vgb_smem[];
    uint2* skv = reinterpret_cast<uint2*>(vgb_smem);                                       // [cap] (local key, payload)
    float4* spts = reinterpret_cast<float4*>(vgb_smem);                                    // after the sort: [cap / 2] points
    uint16_t (*wcnt)[kVgbDigits] = reinterpret_cast<uint16_t (*)[kVgbDigits]>(vgb_smem + (size_t)kVgbCap * 8);   // [WARPS][512]
    uint32_t (*wcnt2)[kVgbDigits / 2] = reinterpret_cast<uint32_t (*)[kVgbDigits / 2]>(&wcnt[0][0]);   // two digits per word
    uint16_t* seghead = &wcnt[0][0];                                                       // before the sort: [cap]
    uint16_t* vstart = &wcnt[0][0];                                                        // after the sort: [cap]
    const float4** segsrc = reinterpret_cast<const float4**>(&wcnt[WARPS][0]);            // [nseg]
    uint32_t* segstart = reinterpret_cast<uint32_t*>(segsrc + a.nseg);                     // [nseg + 1]
    uint32_t* segbase = segstart + a.nseg + 1;                                             // [nseg]
    __shared__ uint32_t scan_ws[WARPS];
    __shared__ uint32_t nb_s, nv_s, inoff_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // run starts: "no run starts here" everywhere first
    for (int i = tid; i < kVgbCap / 2; i += kVgbThreads) reinterpret_cast<uint32_t*>(seghead)[i] = 0xffffffffu;
    __syncthreads();
    const uint32_t b = blockIdx.x;
    const uint32_t nseg = a.nseg;

    // ---- the bucket's segment table: (run, first position, length) -> exclusive prefix of the lengths ----
    uint32_t carry = 0, before = 0;                                // before: points of this thread's runs in earlier buckets
    for (uint32_t r0 = 0; r0 < nseg; r0 += kVgbThreads) {          // nseg <= 1024: at most four rounds
        const uint32_t r = r0 + tid;
        uint32_t len = 0;
        if (r < nseg) {
            const uint32_t s0 = a.soff[(size_t)b * nseg + r], s1 = a.soff[(size_t)(b + 1) * nseg + r];
            len = s1 - s0;
            before += s0;
            segbase[r] = s0;
            segsrc[r] = a.segs[r].src;
        }
        const uint32_t inc = warp_inclusive_scan(len, lane);
        if (lane == 31) scan_ws[warp] = inc;
        __syncthreads();
        uint32_t wpre = 0, tot = 0;
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t s = scan_ws[w];
            if (w < warp) wpre += s;
            tot += s;
        }
        if (r < nseg) {
            const uint32_t st = carry + wpre + inc - len;
            segstart[r] = st;
            if (len && st < (uint32_t)kVgbCap) seghead[st] = (uint16_t)r;
        }
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) { segstart[nseg] = carry; nb_s = carry; inoff_s = 0; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0 && before) atomicAdd(&inoff_s, before);
    uint32_t nb = nb_s;
    if (nb > a.cap) {                                              // does not fit: flag it, contribute nothing
        if (tid == 0) atomicMax(a.info + 1, nb);
        nb = 0;
    }

    const uint32_t key_lo = b == 0 ? 0u : vgb_splitter(a.sorted_samples, b, a.trunc_shift);
    const uint32_t key_hi = b + 1 == a.nbuckets ? a.key_end : vgb_splitter(a.sorted_samples, b + 1, a.trunc_shift);
    // bits of the largest local key -> passes of equal digit width (<= 10 bits)
    int bits = 1;
    {
        const uint32_t range = key_hi > key_lo ? key_hi - key_lo - 1u : 0u;
        bits = 32 - __clz(range | 1u);
    }
    const int npass = (bits + kVgbDigitBits - 1) / kVgbDigitBits;
    const int dbits = (bits + npass - 1) / npass;
    const uint32_t dmask = (1u << dbits) - 1u;

    // every warp takes the same number of 32-element rounds; element order = (warp, round, lane)
    const int rounds = (int)((nb + kVgbThreads - 1) / kVgbThreads);          // <= kVgbItems
    const uint32_t wbase = (uint32_t)warp * (uint32_t)rounds * 32u;

    // ---- gather: element e of the bucket = position (e - segstart[r]) of run r's slice; its key from the point itself
    // (reading the 4-byte packed coordinates here instead was measured 5 % slower for the kernel: the second read of the
    // points, in sorted order, then misses L2) ----
    if (wbase < nb) {
        uint32_t cur;                                              // run of the element before this warp's next round
        {
            uint32_t lo = 0, hi = nseg;                            // last run with segstart <= wbase
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (segstart[mid] <= wbase) lo = mid; else hi = mid;
            }
            cur = lo;
        }
#pragma unroll 1
        for (int g = 0; g < rounds; g += 8) {
            if (wbase + g * 32 >= nb) break;
            uint32_t pay[8];
            float4 p[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t e = wbase + (g + r) * 32 + lane;
                pay[r] = 0xffffffffu;
                if (g + r < rounds && wbase + (g + r) * 32 < nb) {     // warp-uniform
                    const uint32_t hd = e < nb ? (uint32_t)seghead[e] : 0xffffu;
                    const uint32_t bal = __ballot_sync(0xffffffffu, hd != 0xffffu);
                    const uint32_t mine = bal & (lt_mask | (1u << lane));
                    const uint32_t got = __shfl_sync(0xffffffffu, hd, mine ? 31 - __clz(mine) : 0);
                    const uint32_t run = mine ? got : cur;
                    const uint32_t last = __shfl_sync(0xffffffffu, hd, bal ? 31 - __clz(bal) : 0);
                    if (bal) cur = last;
                    if (e < nb) {
                        const uint32_t off = segbase[run] + (e - segstart[run]);
                        pay[r] = (run << kSegShift) | off;
                        p[r] = __ldg(segsrc[run] + off);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t e = wbase + (g + r) * 32 + lane;
                if (g + r < rounds && wbase + (g + r) * 32 < nb) {
                    uint2 kv = make_uint2(0xffffffffu, 0u);            // padding sorts last
                    if (e < nb) kv = make_uint2(voxel_key(p[r], a.vs) - key_lo, pay[r]);
                    skv[e] = kv;
                }
            }
        }
    }
    __syncthreads();

    // ---- stable LSD radix sort of the bucket in shared memory ----
    for (int pass = 0; pass < npass; ++pass) {
        const int shift = pass * dbits;
        uint2 kv[kVgbItems];
        uint16_t rank[kVgbItems];
        for (int j = lane; j <= (int)(dmask >> 1); j += 32) wcnt2[warp][j] = 0;
        // the loads, then the matches, then the counter chain: the 16 loads and the 16 matches of a thread are
        // independent and overlap; only the chain through the warp's counters is sequential
#pragma unroll
        for (int r = 0; r < kVgbItems; ++r)
            if (r < rounds && wbase + r * 32 < nb) kv[r] = skv[wbase + r * 32 + lane];
#pragma unroll
        for (int r = 0; r < kVgbItems; ++r) {
            if (r < rounds && wbase + r * 32 < nb) {                 // warp-uniform
                const uint32_t dg = (kv[r].x >> shift) & dmask;
#ifdef LVREG_VGB_BALLOT
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (int bb = 0; bb < kVgbDigitBits; ++bb) {
                    if (bb < dbits) {
                        const bool bit = (dg & (1u << bb)) != 0;
                        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                        peers &= bal ^ (bit ? 0u : 0xffffffffu);
                    }
                }
#else
                const uint32_t peers = __match_any_sync(0xffffffffu, dg);
#endif
                // lanes before this one with the same digit (5 bits) | size of the group (6 bits)
                rank[r] = (uint16_t)((uint32_t)__popc(peers & lt_mask) | ((uint32_t)__popc(peers) << 8));
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < kVgbItems; ++r) {
            if (r < rounds && wbase + r * 32 < nb) {
                const uint32_t dg = (kv[r].x >> shift) & dmask;
                const uint32_t lower = rank[r] & 0xffu;
                const uint32_t old = wcnt[warp][dg];
                if (lower == 0) wcnt[warp][dg] = (uint16_t)(old + (rank[r] >> 8));
                __syncwarp();
                rank[r] = (uint16_t)(old + lower);
            }
        }
        __syncthreads();
        // digit starts: exclusive prefix over (digit, warp); a thread owns digits 2 tid and 2 tid + 1 (one packed word)
        {
            const bool active = (uint32_t)(2 * tid) <= dmask;
            uint32_t c = 0;
            if (active) {
#pragma unroll
                for (int w = 0; w < WARPS; ++w) c += wcnt2[w][tid];
            }
            const uint32_t c0 = c & 0xffffu, c1 = c >> 16;
            const uint32_t inc = warp_inclusive_scan(c0 + c1, lane);
            if (lane == 31) scan_ws[warp] = inc;
            __syncthreads();
            if (active) {
                uint32_t run = inc - (c0 + c1);
                for (int w = 0; w < warp; ++w) run += scan_ws[w];
                run = run | ((run + c0) << 16);
#pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    const uint32_t cw = wcnt2[w][tid];
                    wcnt2[w][tid] = run;
                    run += cw;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < kVgbItems; ++r) {
            if (r < rounds && wbase + r * 32 < nb) {
                const uint32_t dg = (kv[r].x >> shift) & dmask;
                skv[(uint32_t)wcnt[warp][dg] + rank[r]] = kv[r];
            }
        }
        __syncthreads();
    }

    // ---- voxel starts: head flags in sorted order, ordinals by ballot + running count ----
    uint32_t flags = 0, wtotal = 0;
#pragma unroll
    for (int r = 0; r < kVgbItems; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        bool head = false;
        if (r < rounds && i < nb) head = i == 0 || skv[i].x != skv[i - 1].x;
        const uint32_t bal = __ballot_sync(0xffffffffu, head);
        flags |= (head ? 1u : 0u) << r;
        wtotal += (uint32_t)__popc(bal);
    }
    if (lane == 0) scan_ws[warp] = wtotal;
    // the payloads in STRIPED order (position r * THREADS + tid): what the point loads below go through
    uint32_t pv[kVgbItems];
#pragma unroll
    for (int r = 0; r < kVgbItems; ++r) {
        const uint32_t i = r * kVgbThreads + tid;
        pv[r] = i < nb ? skv[i].y : 0u;
    }
    __syncthreads();                                               // every read of wcnt and skv is done (vstart, spts alias them)
    uint32_t ord = 0;
    {
        uint32_t tot = 0;
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t s = scan_ws[w];
            if (w < warp) ord += s;
            tot += s;
        }
        if (tid == 0) {
            nv_s = tot;
            a.bucket_nvox[b] = tot;
            a.bucket_in[b] = inoff_s;
        }
    }
#pragma unroll
    for (int r = 0; r < kVgbItems; ++r) {
        const bool head = (flags >> r) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, head);
        if (head) vstart[ord + __popc(bal & lt_mask)] = (uint16_t)(wbase + r * 32 + lane);
        ord += (uint32_t)__popc(bal);
    }

    // ---- first half of the points into shared memory, in sorted order (8 loads in flight) ----
    {
        float4 p[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if ((uint32_t)(r * kVgbThreads + tid) < nb) p[r] = __ldg(segsrc[pv[r] >> kSegShift] + (pv[r] & kSegOffMask));
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if ((uint32_t)(r * kVgbThreads + tid) < nb) spts[r * kVgbThreads + tid] = p[r];
    }
    __syncthreads();
    const uint32_t nv = nv_s;

    // ---- centroids: sequential sums in sorted order (= keyframe order, then scan order), half a bucket at a time;
    // the one voxel that straddles the halves carries its partial sum in registers ----
    float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;                  // carry of this thread's straddling voxel
    auto sum_half = [&](uint32_t lo, uint32_t hi, uint32_t base, bool final_half) {
        for (uint32_t v = tid; v < nv; v += kVgbThreads) {
            const uint32_t s0 = vstart[v], s1 = v + 1 < nv ? (uint32_t)vstart[v + 1] : nb;
            if (s0 >= hi || s1 <= lo) continue;
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            if (s0 < lo) { sx = cx; sy = cy; sz = cz; si = ci; }
            const uint32_t j0 = s0 > lo ? s0 : lo, j1 = s1 < hi ? s1 : hi;
            for (uint32_t j = j0; j < j1; ++j) {
                const float4 p = spts[j - lo];
                sx += p.x; sy += p.y; sz += p.z; si += p.w;
            }
            if (s1 > hi && !final_half) { cx = sx; cy = sy; cz = sz; ci = si; continue; }
            const float cnt = (float)(s1 - s0);
            a.tmp[base + v] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        }
    };
    const uint32_t base = inoff_s;                                 // nv <= nb: the bucket's own input range is free
    sum_half(0, HALF, base, nb <= (uint32_t)HALF);
    if (nb > (uint32_t)HALF) {
        __syncthreads();
        float4 p[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if ((uint32_t)((r + 8) * kVgbThreads + tid) < nb)
                p[r] = __ldg(segsrc[pv[r + 8] >> kSegShift] + (pv[r + 8] & kSegOffMask));
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if ((uint32_t)((r + 8) * kVgbThreads + tid) < nb) spts[r * kVgbThreads + tid] = p[r];
        __syncthreads();
        sum_half(HALF, 2 * HALF, base, true);
    }
}

// exclusive scan of the buckets' voxel counts in place (one block; a few thousand values) + the grand total
__global__ void __launch_bounds__(1024) vgb_scan_kernel(uint32_t* __restrict__ bucket_nvox, uint32_t nbuckets,
                                                        uint32_t* __restrict__ info) {
    __shared__ uint32_t ws[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (nbuckets + 1023) / 1024;                 // consecutive values per thread
    const uint32_t i0 = (uint32_t)tid * per;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; ++k)
        if (i0 + k < nbuckets) sum += bucket_nvox[i0 + k];
    const uint32_t inc = warp_inclusive_scan(sum, lane);
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t wpre = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
        const uint32_t s = ws[w];
        if (w < warp) wpre += s;
        tot += s;
    }
    uint32_t run = wpre + inc - sum;
    for (uint32_t k = 0; k < per; ++k)
        if (i0 + k < nbuckets) {
            const uint32_t v = bucket_nvox[i0 + k];
            bucket_nvox[i0 + k] = run;
            run += v;
        }
    if (tid == 0) { bucket_nvox[nbuckets] = tot; info[0] = tot; }
}

// out[offset of bucket b + v] = tmp[input offset of bucket b + v]
// (launched BEFORE the host knows the total, into whatever the output buffer holds: does nothing if that is too small,
// the host then repeats it after growing the buffer)
__global__ void __launch_bounds__(128) vgb_compact_kernel(const float4* __restrict__ tmp, const uint32_t* __restrict__ bucket_out,
                                                          const uint32_t* __restrict__ bucket_in, uint32_t nbuckets,
                                                          uint32_t out_capacity, float4* __restrict__ out) {
    if (bucket_out[nbuckets] > out_capacity) return;
    const uint32_t b = blockIdx.x;
    const uint32_t o0 = bucket_out[b], nv = bucket_out[b + 1] - o0, i0 = bucket_in[b];
    for (uint32_t v = threadIdx.x; v < nv; v += 128) out[o0 + v] = tmp[i0 + v];
}

}  // namespace lvreg
