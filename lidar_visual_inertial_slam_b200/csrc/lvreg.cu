// lvreg.cu -- handle, device buffers and the C ABI of include/lvreg.h.
// Host orchestration only: every data-parallel step is one of the sm_100a kernels in
// voxelgrid.cuh / prims.cuh / knn.cuh / fit.cuh / register.cuh.  There is no CPU fallback:
// lvreg_create fails without a CUDA device, and nothing here links or loads oracle/.
#include "../../include/lvreg.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <deque>
#include <string>
#include <vector>

#include "common.cuh"
#include "depth.cuh"
#include "feature.cuh"
#include "fit.cuh"
#include "icp.cuh"
#include "knn.cuh"
#include "prims.cuh"
#include "projection.cuh"
#include "register.cuh"
#include "register_staged.cuh"
#include "register_warm.cuh"
#include "voxelgrid.cuh"
#include "voxelgrid_mid.cuh"
#include "voxelgrid_bucket.cuh"

using namespace lvreg;

namespace {

static thread_local double g_alloc_ms = 0.0;      // LVREG_DEBUG_ALLOC diagnostics (one handle per host thread)
static thread_local size_t g_alloc_bytes = 0;
static thread_local int g_alloc_calls = 0;
// bump allocator over large slabs: keyframe clouds come from here, so adding a keyframe does not call
// cudaMalloc (one slab holds hundreds of keyframes); everything is returned when the handle is destroyed
struct Arena {
    struct Slab { void* p; size_t bytes; };
    std::vector<Slab> slabs;
    size_t next_slab = 0;            // slabs[0 .. next_slab) are in use, the rest is free for re-use after reset()
    char* cur = nullptr;
    size_t left = 0;
    size_t slab_bytes = (size_t)64 << 20;
    cudaError_t alloc(size_t bytes, void** out) {
        bytes = (bytes + 255) & ~(size_t)255;
        while (bytes > left) {
            if (next_slab < slabs.size()) {                 // a slab kept from an earlier session
                cur = (char*)slabs[next_slab].p;
                left = slabs[next_slab].bytes;
                ++next_slab;
                continue;
            }
            const size_t sz = bytes > slab_bytes ? bytes : slab_bytes;
            void* p = nullptr;
            cudaError_t e = cudaMalloc(&p, sz);
            if (e != cudaSuccess) return e;
            slabs.push_back({p, sz});
            next_slab = slabs.size();
            cur = (char*)p;
            left = sz;
        }
        *out = cur;
        cur += bytes;
        left -= bytes;
        return cudaSuccess;
    }
    // every block is handed back at once (no keyframe is live): the slabs stay allocated and are filled again from
    // the start, so repeated sessions on one handle do not grow device memory
    void reset() {
        next_slab = 0;
        cur = nullptr;
        left = 0;
    }
    void release_all() {
        for (const Slab& sl : slabs) cudaFree(sl.p);
        slabs.clear();
        reset();
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    Arena* arena = nullptr;          // when set, the memory belongs to the arena (never freed individually)
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        struct Timer {
            std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
            ~Timer() { g_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ++g_alloc_calls; }
        } timer;
        g_alloc_bytes += bytes;
        if (arena) {
            const size_t want = (bytes + 255) & ~(size_t)255;
            void* np = nullptr;
            cudaError_t e = arena->alloc(want, &np);      // a previous, smaller block is simply abandoned
            if (e == cudaSuccess) { p = np; cap = want; }
            return e;
        }
        // geometric growth: a buffer that has to grow doubles, so a session re-allocates O(log size) times
        // (cudaFree / cudaMalloc synchronise the device and were measured to stall a call for up to seconds)
        size_t want = bytes + bytes / 4 + 256;
        if (cap && want < 2 * cap) want = 2 * cap;
        want = (want + 255) & ~(size_t)255;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p && !arena) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Keyframe {
    DevBuf cloud[2];          // sensor-frame corner / surf, float4
    uint32_t n[2] = {0, 0};
    float pose[6] = {0, 0, 0, 0, 0, 0};
    // laserCloudMapContainer (MO:106, 942-954): the clouds under the keyframe's current pose, filled on first use and
    // dropped when the poses are corrected (MO:1623), + their bounding boxes (exact: union of them = bbox of any map)
    DevBuf world[2];
    bool world_ok = false;
    bool wsorted[2] = {false, false};   // world[s] is ordered by the voxel index of leaf wleaf[s] (voxelgrid_bucket.cuh)
    float wleaf[2] = {0.f, 0.f};
    DevBuf wkey[2];                     // ... and (wpacked: when they fit 11 | 11 | 10 bits) the points' voxel coordinates relative
                                        // to wminb, packed in 4 bytes
    bool wpacked[2] = {false, false};
    int wminb[2][3] = {{0, 0, 0}, {0, 0, 0}};
    float wmn[2][3] = {{0, 0, 0}, {0, 0, 0}}, wmx[2][3] = {{0, 0, 0}, {0, 0, 0}};
};

struct DepthEntry {            // one entry of cloudQueue / timeQueue (feature_tracker_node.cpp:343-344)
    DevBuf pts;
    uint32_t n = 0;
    double stamp = 0.0;
};

struct MapSide {
    DevBuf ds;                // laserCloud*FromMapDS, float4, VoxelGrid (ascending idx) order
    DevBuf cell_pts, cell_start, hkeys;
    bool hashed = false;      // cell directory: dense prefix array, or an open-addressing hash of the occupied cells
    uint32_t hmask = 0;
    uint32_t m = 0;
    uint64_t n_in = 0;
    GridSpec gs{};
    bool valid = false;
};

enum { EV_BEGIN = 0, EV_UPLOAD, EV_MAP, EV_GRID, EV_DS, EV_REG, EV_COUNT };

// Four independent pipelines ("lanes"), each with its own stream and scratch, so that the
// corner / surf local-map builds and the corner / surf scan down-sampling of one call run
// concurrently and share their host synchronisation points (2 per call instead of 2 per cloud).
constexpr int kLanes = 4;
enum { LANE_MAP_CORNER = 0, LANE_MAP_SURF = 1, LANE_SCAN_CORNER = 2, LANE_SCAN_SURF = 3 };

struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t ev = nullptr;
    DevBuf stage, raw, concat, keys[2], vals[2], sort_scratch, scan_temp, scan_in, vox_start, vox_keys, segs, small;
    DevBuf bsoff, bstatus, bout;            // bucketed VoxelGrid: segment table, chained-scan status, output (swapped with the map)
    uint32_t* pinned = nullptr;             // 64 words of page-locked host memory
    std::vector<Segment> seg_host;
    Segment* seg_pin = nullptr;             // page-locked staging of seg_host (a pageable source makes the copy synchronous)
    size_t seg_pin_cap = 0;
};

}  // namespace

struct lvreg_handle {
    lvreg_params prm;
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    cudaEvent_t ev_main = nullptr;
    Lane lane[kLanes];
    int num_sms = kNumSMs;
    int lpq = 4;
    int tile = 16;
    int imu_available = 0;        // cloudInfo.imu_available / imu_roll_init / imu_pitch_init of the current scan (MO:1347-1366)
    float imu_roll = 0.f, imu_pitch = 0.f;
    int warm_tile = 32;
    int debug_phases = 0;         // LVREG_DEBUG_PHASES=1: per-lane phase time stamps of the VoxelGrid batch on stderr
    cudaEvent_t dbg_ev[kLanes][6] = {};
    unsigned dbg_ev_used = 0;
    double dbg_host_us[8] = {};   // LVREG_DEBUG_PHASES: host clock at call begin / first chain / all enqueued / after sync / ...
    bool kf_cache_enabled = true; // LVREG_KF_CACHE=0: transform the keyframe clouds on every map build (experiments)
    DevBuf kfmm;                  // bounding-box slots of the keyframe clouds being cached
    bool vg_mid_enabled = true;   // LVREG_VG_MID=0 disables the cooperative single-launch VoxelGrid (experiments)
    bool vg_bucket_enabled = true;  // LVREG_VG_BUCKET=0: large cached maps through the device-wide sort (experiments)
    bool vg_cached_first = false;   // LVREG_VG_CACHED_FIRST=1: enqueue the whole chain of the cached (map) jobs before the scan lanes
                                    // (measured slower at C3: the scan filters then compete with the bucket kernel instead of
                                    // overlapping the latency-bound sample sort)
    bool dbg_nowkey = false, dbg_dealt = false, dbg_alloc = false;   // LVREG_DEBUG_NOWKEY / LVREG_DEALT / LVREG_DEBUG_ALLOC
    uint32_t vg_bucket_cap = 0;     // LVREG_VG_BUCKET_CAP: smaller bucket capacity, to exercise the overflow fallback
    uint32_t vg_bucket_fallbacks = 0, vg_bucket_jobs = 0;
    bool kernel_timing = false;        // lvreg_enable_kernel_timing: event pairs around the bucket kernels of a call
    cudaEvent_t kt_ev[2][2] = {};
    bool kt_set[2] = {false, false};
    uint32_t kt_n_in[2] = {0, 0};
    int debug_tiles = 0;          // LVREG_DEBUG_TILES=1: record per-tile durations of iteration 1
    uint32_t debug_ntiles = 0;
    int force_tpq = -1;           // -1 auto, 0 grouped, 1 thread-per-query (LVREG_TPQ)
    int reg_variant_env = -1;     // LVREG_REG: -1 auto, 0 grouped, 1 thread-per-query, 2 staged (shared-memory search),
                                  // 3 warm (thread per query, warm-started radius, static tiles)
    int reg_occ_variant = -1;
    std::vector<Keyframe*> kfs;
    std::vector<Keyframe*> kf_free;   // keyframes of a cleared session: their device buffers are reused
    Arena kf_arena;                   // backing store of every keyframe cloud
    MapSide map[2];
    DevBuf scan_ds[2];
    DevBuf scan_sorted[2];        // the same points in Morton order of their 2 m cell: what the registration kernels read
    bool scan_sorted_ok[2] = {false, false};
    uint32_t n_scan[2] = {0, 0};
    // loop closure: [0] source (cureKeyframeCloud), [1] target (prevKeyframeCloud) + its search grid
    MapSide icp_cloud[2];
    MapSide icp_coarse;                    // second, coarse search grid over the target's points
    // front end: deskewed + projected scan (extractedCloud and the CloudInfo side channels)
    DevBuf proj_raw, proj_imu, proj_pts, proj_rangein, proj_colin, proj_small, proj_owner, proj_cloud, proj_range, proj_col,
        proj_rings;
    uint32_t n_proj = 0;
    int proj_n_scan = 0;
    std::vector<int32_t> proj_start_host, proj_end_host;
    // visual side: stacked depth cloud (depthCloud) and the scratch of get_depth
    std::deque<DepthEntry*> depth_queue;
    std::vector<DepthEntry*> depth_free;   // expired entries, buffers kept for reuse (no cudaMalloc per scan)
    DevBuf depth_cloud, depth_concat, depth_bins, depth_local, depth_unit, depth_feat, depth_out, depth_f3d;
    uint32_t n_depth = 0, n_depth_local = 0;
    DevBuf icp_cur, icp_partials, icp_state, icp_idx, icp_d2;
    // scratch of the main stream
    DevBuf feat_pts, feat_range, feat_col, feat_rings, feat_curv, feat_picked, feat_label, feat_flag, feat_ringof,
        feat_cidx, feat_ccnt, feat_pos, feat_cand, feat_spec, feat_idx, feat_pidx, feat_corner, feat_surf;
    uint32_t n_feat[2] = {0, 0};
    DevBuf stagestats, nnprev[2];
    DevBuf vgout, partials, regout, lmstate, posebuf, tilectr, tilens, qbuf, idxbuf, d2buf, brute_partial, coeffbuf, flagbuf;
    void* pinned = nullptr;       // 64 KB page-locked scratch: lanes use [0, 1 KB), RegOut lives at +4 KB
    cudaEvent_t ev[EV_COUNT];
    bool ev_set[EV_COUNT];
    lvreg_timings last{};
    int call_launches = 0;
    uint64_t total_launches = 0;
    std::string err;
    int reg_max_blocks_per_sm = 0;
    int reg_tiles_per_block = 4;   // LVREG_REG_TPB: a scan of fewer tiles than resident warps uses tiles / 4 blocks (four warps of a
                                   // block busy): fewer blocks at the grid barrier and in the partial sums (C1: 0.149 -> 0.141 ms;
                                   // 8 per block: 0.149)
};

namespace {

#define CK(expr)                                                                         \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            char _b[512];                                                                \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                \
            h->err = _b;                                                                 \
            return LVREG_ERR_CUDA;                                                       \
        }                                                                                \
    } while (0)

#define CKS(expr)                       \
    do {                                \
        int _s = (expr);                \
        if (_s != LVREG_OK) return _s;  \
    } while (0)

inline int fail(lvreg_handle* h, int code, const char* msg) {
    h->err = msg;
    return code;
}

inline uint32_t nblk(uint32_t n, uint32_t per) { return (n + per - 1) / per; }

// small device scalars inside h->small
enum { SM_MM = 0 /*6 u32*/, SM_NVOX = 8, SM_TOTAL = 9, SM_CONV = 10, SM_SMALLVG = 32 /*VgSmallInfo, 8 words*/, SM_WORDS = 64 };

inline void launched(lvreg_handle* h, int k = 1) { h->call_launches += k; }
inline double host_now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

inline void begin_call(lvreg_handle* h) {
    h->call_launches = 0;
    for (int i = 0; i < EV_COUNT; ++i) h->ev_set[i] = false;
    h->kt_set[0] = h->kt_set[1] = false;
    memset(&h->last, 0, sizeof(h->last));
}
inline void mark(lvreg_handle* h, int which) {
    cudaEventRecord(h->ev[which], h->st);
    h->ev_set[which] = true;
}
inline float span(lvreg_handle* h, int a, int b) {
    float ms = 0.f;
    if (h->ev_set[a] && h->ev_set[b]) cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]);
    return ms;
}
inline void end_call(lvreg_handle* h) {
    h->total_launches += (uint64_t)h->call_launches;
    h->last.kernel_launches = h->call_launches;
}

// ---- lanes -------------------------------------------------------------------------------------
inline void lanes_fork(lvreg_handle* h, unsigned mask) {          // lane streams wait for the main stream
    cudaEventRecord(h->ev_main, h->st);
    for (int l = 0; l < kLanes; ++l)
        if (mask & (1u << l)) cudaStreamWaitEvent(h->lane[l].st, h->ev_main, 0);
}
inline void lanes_join(lvreg_handle* h, unsigned mask) {          // main stream waits for the lanes
    for (int l = 0; l < kLanes; ++l)
        if (mask & (1u << l)) {
            cudaEventRecord(h->lane[l].ev, h->lane[l].st);
            cudaStreamWaitEvent(h->st, h->lane[l].ev, 0);
        }
}
inline cudaError_t lanes_sync(lvreg_handle* h, unsigned mask) {   // ONE host wait for all lanes in mask
    lanes_join(h, mask);
    return cudaStreamSynchronize(h->st);
}

// A failing call must not leave half-built state behind: unless commit() is reached, the lanes are joined and
// drained and whatever the call was rebuilding is marked empty / invalid, so that a later lvreg_scan2map cannot run
// on counts and data that do not belong together.
struct CallGuard {
    lvreg_handle* h;
    bool scan, map;
    bool ok = false;
    CallGuard(lvreg_handle* h_, bool touches_scan, bool touches_map) : h(h_), scan(touches_scan), map(touches_map) {}
    void commit() { ok = true; }
    ~CallGuard() {
        if (ok) return;
        lanes_join(h, (1u << kLanes) - 1u);
        cudaStreamSynchronize(h->st);
        cudaGetLastError();
        if (scan) {
            h->n_scan[0] = h->n_scan[1] = 0;
            h->scan_sorted_ok[0] = h->scan_sorted_ok[1] = false;
        }
        if (map) h->map[0].valid = h->map[1].valid = false;
        h->total_launches += (uint64_t)h->call_launches;
    }
};

// ---- cloud transfer ----------------------------------------------------------------------------
int check_cloud(lvreg_handle* h, const lvreg_cloud* c) {
    if (!c) return fail(h, LVREG_ERR_INVALID, "null cloud");
    if (c->n && !c->data) return fail(h, LVREG_ERR_INVALID, "cloud has n > 0 but no data");
    if (c->stride < 16 || (c->stride & 3) || c->intensity_offset + 4 > c->stride || (c->intensity_offset & 3))
        return fail(h, LVREG_ERR_INVALID, "bad cloud stride / intensity_offset");
    if (c->n > 0x7fffffffull) return fail(h, LVREG_ERR_INVALID, "cloud too large");
    return LVREG_OK;
}

// any layout (host or device) -> device float4 {x,y,z,intensity}, on stream `st`
int upload_cloud(lvreg_handle* h, const lvreg_cloud* c, DevBuf& dst, DevBuf& stage, cudaStream_t st) {
    CKS(check_cloud(h, c));
    const uint32_t n = (uint32_t)c->n;
    CK(dst.reserve((size_t)(n ? n : 1) * 16));
    if (n == 0) return LVREG_OK;
    const bool packed = c->stride == 16 && c->intensity_offset == 12;
    if (packed) {
        CK(cudaMemcpyAsync(dst.p, c->data, (size_t)n * 16,
                           c->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        return LVREG_OK;
    }
    const uint8_t* src = (const uint8_t*)c->data;
    if (!c->on_device) {
        CK(stage.reserve((size_t)n * c->stride));
        CK(cudaMemcpyAsync(stage.p, c->data, (size_t)n * c->stride, cudaMemcpyHostToDevice, st));
        src = stage.as<uint8_t>();
    }
    if (((uintptr_t)src & 15) != 0 && (c->stride & 15) == 0)
        return fail(h, LVREG_ERR_INVALID, "device clouds must be 16-byte aligned");
    pack_kernel<<<nblk(n, 256), 256, 0, st>>>(src, n, c->stride, c->intensity_offset, dst.as<float4>());
    launched(h);
    CK(cudaGetLastError());
    return LVREG_OK;
}

// main-stream download (synchronous)
int download_cloud(lvreg_handle* h, const float4* src, uint32_t n, lvreg_cloud_out* out) {
    if (!out || (!out->data && n)) return fail(h, LVREG_ERR_INVALID, "null output cloud");
    if (out->capacity < n) return fail(h, LVREG_ERR_CAPACITY, "output cloud too small");
    if (out->stride < 16 || (out->stride & 3) || out->intensity_offset + 4 > out->stride)
        return fail(h, LVREG_ERR_INVALID, "bad output stride / intensity_offset");
    if (n == 0) return LVREG_OK;
    const bool packed = out->stride == 16 && out->intensity_offset == 12;
    DevBuf& stage = h->lane[LANE_SCAN_CORNER].stage;
    if (packed) {
        CK(cudaMemcpyAsync(out->data, src, (size_t)n * 16,
                           out->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->st));
    } else if (out->on_device) {
        CK(cudaMemsetAsync(out->data, 0, (size_t)n * out->stride, h->st));
        unpack_kernel<<<nblk(n, 256), 256, 0, h->st>>>(src, n, out->stride, out->intensity_offset, (uint8_t*)out->data);
        launched(h);
    } else {
        CK(stage.reserve((size_t)n * out->stride));
        CK(cudaMemsetAsync(stage.p, 0, (size_t)n * out->stride, h->st));
        unpack_kernel<<<nblk(n, 256), 256, 0, h->st>>>(src, n, out->stride, out->intensity_offset, stage.as<uint8_t>());
        launched(h);
        CK(cudaMemcpyAsync(out->data, stage.p, (size_t)n * out->stride, cudaMemcpyDeviceToHost, h->st));
    }
    CK(cudaStreamSynchronize(h->st));
    return LVREG_OK;
}

// ---- VoxelGrid driver (batched over lanes) -------------------------------------------------------
int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) ++b;
    return b;
}

int ensure_sort_buffers(lvreg_handle* h, Lane& L, uint32_t n) {
    for (int i = 0; i < 2; ++i) {
        CK(L.keys[i].reserve((size_t)n * 4));
        CK(L.vals[i].reserve((size_t)n * 4));
    }
    CK(L.sort_scratch.reserve(sort_scratch_words(n) * 4));
    CK(L.scan_temp.reserve((size_t)(scan_num_tiles(n) + 2) * 4));
    return LVREG_OK;
}

struct VgJob {
    int lane = 0;
    const float4* pts = nullptr;       // device input (ignored when from_segments)
    uint32_t n = 0;
    bool from_segments = false;        // input = Lane::seg_host, transformed + concatenated into Lane::concat
    bool cached = false;               // segments point at cached WORLD-frame clouds; mn / mx hold their exact bbox
    bool bucket = false;               // ... every one of them ordered by voxel index: sample-sort path (voxelgrid_bucket.cuh)
    DevBuf* morton_out = nullptr;      // optional (single-launch path only): the output once more in Morton order
    bool morton_done = false;
    float leaf = 0.f;
    DevBuf* out = nullptr;
    uint32_t* n_out = nullptr;
    bool want_out_keys = false;        // per-voxel idx into Lane::vox_keys
    uint32_t* d_point_keys = nullptr;  // optional device array (n): per-point idx
    bool small = false;                // internal: handled by the single-block kernel
    bool mid = false;                  // internal: handled by the cooperative single-launch kernel
    // results
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};   // bbox of the input cloud
    int passthrough = 0;
    // internal
    VoxelSpec vs{};
    int cur = 0;
    uint32_t nbuckets = 0, compact_cap = 0;
    bool bucket_enqueued = false;
    bool mark_map_stage = false;       // record EV_MAP on the main stream when the filters are done (before the grids)
};

// segment table of a lane -> device, through page-locked staging so that the copy really is asynchronous.  Every
// VoxelGrid batch ends with a synchronisation of its lanes, so the staging buffer is free again by the next upload.
int upload_segs(lvreg_handle* h, Lane& L, cudaStream_t st) {
    const size_t n = L.seg_host.size();
    CK(L.segs.reserve(n * sizeof(Segment)));
    if (n > L.seg_pin_cap) {
        if (L.seg_pin) cudaFreeHost(L.seg_pin);
        L.seg_pin = nullptr;
        L.seg_pin_cap = 0;
        const size_t cap = n < 256 ? 256 : 2 * n;
        CK(cudaMallocHost((void**)&L.seg_pin, cap * sizeof(Segment)));
        L.seg_pin_cap = cap;
    }
    memcpy(L.seg_pin, L.seg_host.data(), n * sizeof(Segment));
    CK(cudaMemcpyAsync(L.segs.p, L.seg_pin, n * sizeof(Segment), cudaMemcpyHostToDevice, st));
    return LVREG_OK;
}

// Runs up to kLanes VoxelGrid filters concurrently.  Two host synchronisation points in total
// (bounding boxes, voxel counts).  On return the centroid kernels are enqueued on the lane
// streams; the caller joins the lanes before consuming `out` on the main stream.
int voxelgrid_batch(lvreg_handle* h, VgJob* jobs, int nj) {
    unsigned mask = 0;
    for (int j = 0; j < nj; ++j) mask |= 1u << jobs[j].lane;
    // the largest cloud is the critical path: enqueue its work first in every phase
    int order[kLanes] = {0, 1, 2, 3};
    for (int a2 = 0; a2 < nj; ++a2)
        for (int b2 = a2 + 1; b2 < nj; ++b2)
            if (jobs[order[b2]].n > jobs[order[a2]].n) { int t2 = order[a2]; order[a2] = order[b2]; order[b2] = t2; }
    // Two rounds: the jobs over cached world-frame clouds first -- they depend on nothing, and the largest of them is the
    // critical path of the call, so their whole chain is enqueued before the host spends time on the other lanes --
    // then the rest, which share the bounding-box synchronisation.
    for (int round = 0; round < 2; ++round) {
    // ---- phase 1: (transform + concatenate +) bounding box ----
    unsigned sync1_mask = 0;
    for (int jo = 0; jo < nj; ++jo) {
        VgJob& J = jobs[order[jo]];
        if ((J.cached && J.n > 0 && h->vg_cached_first) != (round == 0)) continue;
        Lane& L = h->lane[J.lane];
        *J.n_out = 0;
        J.passthrough = 0;
        J.small = false;
        J.bucket_enqueued = false;
        if (J.n == 0) {
            CK(J.out->reserve(16));
            continue;
        }
        if (!(J.leaf > 0.f)) return fail(h, LVREG_ERR_INVALID, "leaf size must be positive");
        J.small = !J.from_segments && J.n <= (uint32_t)kVgSmallMax;
        J.mid = !J.small && J.n <= (uint32_t)kVgMidMax && h->vg_mid_enabled;
        if (J.mid) {
            // the whole filter in one cooperative launch (voxelgrid_mid.cuh): no host round trip until the shared
            // final synchronisation, one launch instead of ~15
            VgMidArgs a;
            a.n = J.n;
            a.leaf = J.leaf;
            a.pts = J.pts;
            a.segs = nullptr;
            a.nseg = 0;
            a.world = nullptr;
            if (J.from_segments) {
                CKS(upload_segs(h, L, L.st));
                CK(L.concat.reserve((size_t)J.n * 16));
                a.segs = L.segs.as<Segment>();
                a.nseg = (uint32_t)L.seg_host.size();
                a.world = L.concat.as<float4>();
                a.pts = nullptr;
            }
            for (int i = 0; i < 2; ++i) {
                CK(L.keys[i].reserve((size_t)J.n * 4));
                CK(L.vals[i].reserve((size_t)J.n * 4));
            }
            CK(L.sort_scratch.reserve((vg_mid_hist_words() + vg_mid_bb_floats()) * 4));
            CK(L.vox_start.reserve((size_t)J.n * 4));
            CK(J.out->reserve((size_t)J.n * 16));
            a.k0 = L.keys[0].as<uint32_t>(); a.v0 = L.vals[0].as<uint32_t>();
            a.k1 = L.keys[1].as<uint32_t>(); a.v1 = L.vals[1].as<uint32_t>();
            a.tile_hist = L.sort_scratch.as<uint32_t>();
            a.tile_bb = reinterpret_cast<float*>(L.sort_scratch.as<uint32_t>() + vg_mid_hist_words());
            a.vstart = L.vox_start.as<uint32_t>();
            a.out = J.out->as<float4>();
            a.out_keys = nullptr;
            if (J.want_out_keys) {
                CK(L.vox_keys.reserve((size_t)J.n * 4));
                a.out_keys = L.vox_keys.as<uint32_t>();
            }
            a.point_keys = J.d_point_keys;
            a.morton_out = nullptr;
            J.morton_done = false;
            if (J.morton_out) {
                CK(J.morton_out->reserve((size_t)J.n * 16));
                a.morton_out = J.morton_out->as<float4>();
                J.morton_done = true;              // unless PCL's passthrough rule fires (checked after the synchronisation)
            }
            VgSmallInfo* d_info = reinterpret_cast<VgSmallInfo*>(L.small.as<uint32_t>() + SM_SMALLVG);
            a.info = d_info;
            void* kargs[] = {&a};
            CK(cudaLaunchCooperativeKernel((void*)voxelgrid_mid_kernel, dim3(nblk(J.n, kVgMidTile)), dim3(kVgMidThreads), kargs, 0, L.st));
            launched(h);
            CK(cudaMemcpyAsync(L.pinned + 32, d_info, sizeof(VgSmallInfo), cudaMemcpyDeviceToHost, L.st));
            if (J.from_segments) J.pts = L.concat.as<float4>();
            J.small = true;                    // from here on it is handled like the single-block path
            continue;
        }
        if (J.small) {
            // the whole filter in one block, no host round trip until the shared final synchronisation
            CK(J.out->reserve((size_t)J.n * 16));
            uint32_t* okeys = nullptr;
            if (J.want_out_keys) {
                CK(L.vox_keys.reserve((size_t)J.n * 4));
                okeys = L.vox_keys.as<uint32_t>();
            }
            VgSmallInfo* d_info = reinterpret_cast<VgSmallInfo*>(L.small.as<uint32_t>() + SM_SMALLVG);
            voxelgrid_small_kernel<<<1, kVgSmallThreads, 0, L.st>>>(J.pts, J.n, J.leaf, J.out->as<float4>(), okeys, J.d_point_keys, d_info);
            launched(h);
            CK(cudaMemcpyAsync(L.pinned + 32, d_info, sizeof(VgSmallInfo), cudaMemcpyDeviceToHost, L.st));
            continue;
        }
        if (J.cached) {                                // world-frame clouds and their exact bbox are already known
            CKS(upload_segs(h, L, L.st));
            continue;
        }
        uint32_t* mm = L.small.as<uint32_t>() + SM_MM;
        CK(cudaMemsetAsync(mm, 0xff, 3 * sizeof(uint32_t), L.st));
        CK(cudaMemsetAsync(mm + 3, 0, 3 * sizeof(uint32_t), L.st));
        if (J.from_segments) {
            CKS(upload_segs(h, L, L.st));
            CK(L.concat.reserve((size_t)J.n * 16));
            transform_concat_kernel<<<min(nblk(J.n, 2048), (uint32_t)h->num_sms * 16), 256, 0, L.st>>>(
                L.segs.as<Segment>(), (uint32_t)L.seg_host.size(), J.n, L.concat.as<float4>(), mm);
            J.pts = L.concat.as<float4>();
        } else {
            minmax_kernel<<<min(nblk(J.n, 256), (uint32_t)h->num_sms * 8), 256, 0, L.st>>>(J.pts, J.n, mm);
        }
        launched(h);
        CK(cudaMemcpyAsync(L.pinned, mm, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
        sync1_mask |= 1u << J.lane;
    }
    if (sync1_mask) CK(lanes_sync(h, sync1_mask));
    // ---- phase 2: keys, stable sort, run heads ----
    for (int jo = 0; jo < nj; ++jo) {
        VgJob& J = jobs[order[jo]];
        if ((J.cached && J.n > 0 && h->vg_cached_first) != (round == 0)) continue;
        if (J.n == 0 || J.small) continue;
        Lane& L = h->lane[J.lane];
        if (!J.cached)
            for (int a = 0; a < 3; ++a) {
                J.mn[a] = ordered_to_float(L.pinned[a]);
                J.mx[a] = ordered_to_float(L.pinned[3 + a]);
            }
        // PCL voxel_grid.hpp: leaf-size overflow rule and bounds, fp32 exactly as PCL computes them
        const float inv = 1.0f / J.leaf;
        int64_t d[3];
        for (int a = 0; a < 3; ++a) d[a] = (int64_t)((J.mx[a] - J.mn[a]) * inv) + 1;
        if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) {
            CK(J.out->reserve((size_t)J.n * 16));
            CK(cudaMemcpyAsync(J.out->p, J.pts, (size_t)J.n * 16, cudaMemcpyDeviceToDevice, L.st));
            if (J.d_point_keys) CK(cudaMemsetAsync(J.d_point_keys, 0, (size_t)J.n * 4, L.st));
            *J.n_out = J.n;
            J.passthrough = 1;
            continue;
        }
        VoxelSpec& vs = J.vs;
        vs.inv = inv;
        int div_b[3];
        for (int a = 0; a < 3; ++a) {
            vs.min_b[a] = (int)floorf(J.mn[a] * inv);
            int max_b = (int)floorf(J.mx[a] * inv);
            div_b[a] = max_b - vs.min_b[a] + 1;
        }
        vs.mul[0] = 1;
        vs.mul[1] = div_b[0];
        vs.mul[2] = div_b[0] * div_b[1];
        vs.key_bits = bits_for((uint64_t)div_b[0] * div_b[1] * div_b[2] - 1);
        if (J.bucket && !J.want_out_keys && !J.d_point_keys) {
            // sorted runs -> sample sort with the buckets in shared memory (voxelgrid_bucket.cuh)
            const uint32_t nseg = (uint32_t)L.seg_host.size();
            const uint32_t nsamp = (J.n + kVgbSample - 1) / (uint32_t)kVgbSample;     // every q with q * S < n
            const uint32_t nbuckets = nsamp ? (nsamp + kVgbStride - 1) / kVgbStride : 1;
            if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][0], L.st);
            CKS(ensure_sort_buffers(h, L, nsamp ? nsamp : 1));
            CK(L.bsoff.reserve((size_t)(nbuckets + 1) * nseg * 4));
            CK(L.bstatus.reserve((size_t)(2 * nbuckets + 8) * 4));
            CK(L.bout.reserve((size_t)J.n * 16));
            int cur = 0;
            const int trunc_shift = 0;      // truncated sample keys (one sort pass less) unbalance the buckets: a block of 16 keys can hold 4 k points
            const int sample_bits = vs.key_bits - trunc_shift;
            CK(L.vox_start.reserve((size_t)(nsamp ? nsamp : 1) * 4));
            if (nsamp) {
                radix_sort_prepare(L.sort_scratch.as<uint32_t>(), nsamp, sample_bits, L.st);
                vgb_sample_kernel<<<min(nblk(nsamp, 1024), (uint32_t)h->num_sms * 2), 256, 0, L.st>>>(
                    L.segs.as<Segment>(), nseg, nsamp, vs, trunc_shift, sort_num_passes(sample_bits), L.vox_start.as<uint32_t>(),
                    L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), sort_ghist_ptr(L.sort_scratch.as<uint32_t>(), nsamp));
                launched(h);
                cur = radix_sort_run(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), L.keys[1].as<uint32_t>(),
                                     L.vals[1].as<uint32_t>(), nsamp, sample_bits, L.sort_scratch.as<uint32_t>(), true, L.st,
                                     &h->call_launches);
            }
            if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][1], L.st);
            const uint32_t* sorted_samples = L.keys[cur].as<uint32_t>();
            vgb_split_kernel<<<nblk((nbuckets + 1) * nseg, 256), 256, 0, L.st>>>(L.segs.as<Segment>(), nseg, sorted_samples,
                                                                                L.vox_start.as<uint32_t>(), nbuckets, trunc_shift,
                                                                                vs, L.bsoff.as<uint32_t>());
            launched(h);
            if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][2], L.st);
            uint32_t* bnvox = L.bstatus.as<uint32_t>();            // [nbuckets + 1]
            uint32_t* bin = bnvox + nbuckets + 1;                  // [nbuckets]
            uint32_t* binfo = bin + nbuckets;                      // [2]
            CK(cudaMemsetAsync(binfo, 0, 8, L.st));
            VgbArgs a;
            a.segs = L.segs.as<Segment>();
            a.nseg = nseg;
            a.nbuckets = nbuckets;
            a.cap = h->vg_bucket_cap && h->vg_bucket_cap < (uint32_t)kVgbCap ? h->vg_bucket_cap : (uint32_t)kVgbCap;
            a.key_end = (uint32_t)((uint64_t)div_b[0] * div_b[1] * div_b[2]);
            a.trunc_shift = trunc_shift;
            a.vs = vs;
            a.sorted_samples = sorted_samples;
            a.soff = L.bsoff.as<uint32_t>();
            a.tmp = L.bout.as<float4>();
            a.bucket_nvox = bnvox;
            a.bucket_in = bin;
            a.info = binfo;
            const bool timed = h->kernel_timing && J.lane < 2;
            if (timed) cudaEventRecord(h->kt_ev[J.lane][0], L.st);
            vgb_bucket_kernel<<<nbuckets, kVgbThreads, vgb_smem_bytes(nseg), L.st>>>(a);
            if (timed) { cudaEventRecord(h->kt_ev[J.lane][1], L.st); h->kt_set[J.lane] = true; h->kt_n_in[J.lane] = J.n; }
            vgb_scan_kernel<<<1, 1024, 0, L.st>>>(bnvox, nbuckets, binfo);
            launched(h, 2);
            CK(cudaMemcpyAsync(L.pinned + 8, binfo, 8, cudaMemcpyDeviceToHost, L.st));
            if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][3], L.st);
            J.nbuckets = nbuckets;
            // the slices are closed up right away (below, after every chain is enqueued) if the output buffer of the
            // previous build is, as usual, large enough
            J.compact_cap = (uint32_t)std::min<size_t>(J.out->cap / 16, 0xffffffffu);
            J.bucket_enqueued = true;
            continue;
        }
        J.bucket = false;
        CKS(ensure_sort_buffers(h, L, J.n));
        if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][0], L.st);
        radix_sort_prepare(L.sort_scratch.as<uint32_t>(), J.n, vs.key_bits, L.st);
        if (J.cached)
            voxel_keys_seg_kernel<<<min(nblk(J.n, 2048), (uint32_t)h->num_sms * 8), 256, 0, L.st>>>(
                L.segs.as<Segment>(), (uint32_t)L.seg_host.size(), J.n, vs, L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                sort_ghist_ptr(L.sort_scratch.as<uint32_t>(), J.n), sort_num_passes(vs.key_bits));
        else
            voxel_keys_kernel<<<min(nblk(J.n, 256 * 8), (uint32_t)h->num_sms * 8), 256, 0, L.st>>>(
                J.pts, J.n, vs, L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                sort_ghist_ptr(L.sort_scratch.as<uint32_t>(), J.n), sort_num_passes(vs.key_bits));
        launched(h);
        if (J.d_point_keys)
            CK(cudaMemcpyAsync(J.d_point_keys, L.keys[0].p, (size_t)J.n * 4, cudaMemcpyDeviceToDevice, L.st));
        if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][1], L.st);
        J.cur = radix_sort_run(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), L.keys[1].as<uint32_t>(),
                               L.vals[1].as<uint32_t>(), J.n, vs.key_bits, L.sort_scratch.as<uint32_t>(), true, L.st,
                               &h->call_launches);
        if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][2], L.st);
        CK(L.vox_start.reserve((size_t)J.n * 4));
        uint32_t* d_nvox = L.small.as<uint32_t>() + SM_NVOX;
        exclusive_scan(HeadFlagIn{L.keys[J.cur].as<uint32_t>()}, VoxelStartOut{L.vox_start.as<uint32_t>()}, J.n,
                       L.scan_temp.as<uint32_t>(), d_nvox, L.st, &h->call_launches);
        CK(cudaMemcpyAsync(L.pinned + 8, d_nvox, 4, cudaMemcpyDeviceToHost, L.st));
        if (h->debug_phases) cudaEventRecord(h->dbg_ev[J.lane][3], L.st);
    }
    }   // rounds
    // compaction of the sample-sort jobs (after every chain is enqueued: the largest job's chain goes out first)
    for (int jo = 0; jo < nj; ++jo) {
        VgJob& J = jobs[order[jo]];
        if (!J.bucket_enqueued || !J.compact_cap) continue;
        Lane& L = h->lane[J.lane];
        const uint32_t* bnvox = L.bstatus.as<uint32_t>();
        vgb_compact_kernel<<<J.nbuckets, 128, 0, L.st>>>(L.bout.as<float4>(), bnvox, bnvox + J.nbuckets + 1, J.nbuckets,
                                                        J.compact_cap, J.out->as<float4>());
        launched(h);
        if (h->debug_phases) { cudaEventRecord(h->dbg_ev[J.lane][4], L.st); h->dbg_ev_used |= 1u << J.lane; }
    }
    // local-map jobs: the filters end here for the stage timings (the host's synchronisation below and what it enqueues
    // afterwards belong to the grid stage)
    {
        bool any = false;            // only where the filter is a long chain: small maps do not pay for the extra join
        for (int j = 0; j < nj; ++j) any = any || (jobs[j].mark_map_stage && jobs[j].bucket_enqueued);
        if (any && !h->ev_set[EV_MAP]) {
            lanes_join(h, mask);
            mark(h, EV_MAP);
        }
    }
    if (h->debug_phases) h->dbg_host_us[2] = host_now_us();
    CK(lanes_sync(h, mask));
    if (h->debug_phases) h->dbg_host_us[3] = host_now_us();
    // ---- phase 3: centroids (left running on the lane streams) ----
    for (int jo = 0; jo < nj; ++jo) {
        VgJob& J = jobs[order[jo]];
        if (J.small) {                                 // everything already happened in one block
            const VgSmallInfo* info = reinterpret_cast<const VgSmallInfo*>(h->lane[J.lane].pinned + 32);
            *J.n_out = info->nvox;
            J.passthrough = info->passthrough;
            if (J.passthrough) J.morton_done = false;
            for (int a = 0; a < 3; ++a) { J.mn[a] = info->mn[a]; J.mx[a] = info->mx[a]; }
            continue;
        }
        if (J.n == 0 || J.passthrough) continue;
        Lane& L = h->lane[J.lane];
        if (J.bucket) {
            const uint32_t a_cap = h->vg_bucket_cap && h->vg_bucket_cap < (uint32_t)kVgbCap ? h->vg_bucket_cap : (uint32_t)kVgbCap;
            if (L.pinned[9] > a_cap) {                 // a bucket did not fit: redo this job with the device-wide sort
                ++h->vg_bucket_fallbacks;
                J.bucket = false;
                CKS(voxelgrid_batch(h, &J, 1));
                continue;
            }
            ++h->vg_bucket_jobs;
            const uint32_t nvox_b = L.pinned[8];
            *J.n_out = nvox_b;
            if (nvox_b > J.compact_cap) {              // first build, or the map grew past the buffer: grow it, then compact
                if (h->debug_phases) fprintf(stderr, "[lvreg] lane %d: compaction repeated after the synchronisation (%u voxels, buffer for %u)\n", J.lane, nvox_b, J.compact_cap);
                CK(J.out->reserve((size_t)nvox_b * 16 + (size_t)nvox_b * 2));
                const uint32_t* bnvox = L.bstatus.as<uint32_t>();
                vgb_compact_kernel<<<J.nbuckets, 128, 0, L.st>>>(L.bout.as<float4>(), bnvox, bnvox + J.nbuckets + 1, J.nbuckets,
                                                                nvox_b, J.out->as<float4>());
                launched(h);
            }
            if (nvox_b == 0) CK(J.out->reserve(16));
            continue;
        }
        const uint32_t nvox = L.pinned[8];
        CK(J.out->reserve((size_t)(nvox ? nvox : 1) * 16));
        uint32_t* okeys = nullptr;
        if (J.want_out_keys) {
            CK(L.vox_keys.reserve((size_t)(nvox ? nvox : 1) * 4));
            okeys = L.vox_keys.as<uint32_t>();
        }
        if (J.cached)
            centroid_seg_kernel<<<nblk(nvox, 128), 128, 0, L.st>>>(L.segs.as<Segment>(), L.keys[J.cur].as<uint32_t>(),
                                                                   L.vals[J.cur].as<uint32_t>(), L.vox_start.as<uint32_t>(),
                                                                   L.small.as<uint32_t>() + SM_NVOX, J.n, J.out->as<float4>());
        else
            centroid_kernel<<<nblk(nvox, 128), 128, 0, L.st>>>(J.pts, L.keys[J.cur].as<uint32_t>(), L.vals[J.cur].as<uint32_t>(),
                                                               L.vox_start.as<uint32_t>(), L.small.as<uint32_t>() + SM_NVOX,
                                                               J.n, J.out->as<float4>(), okeys);
        launched(h);
        if (h->debug_phases) { cudaEventRecord(h->dbg_ev[J.lane][4], L.st); h->dbg_ev_used |= 1u << J.lane; }
        *J.n_out = nvox;
    }
    CK(cudaGetLastError());
    return LVREG_OK;
}

// ---- search-grid build (on a lane stream; no host synchronisation when the bbox is known) --------
int build_grid(lvreg_handle* h, Lane& L, MapSide& ms, const float* bb_min, const float* bb_max,
               float cell_override = 0.f, const float4* pts_override = nullptr, bool allow_hash = false) {
    const uint32_t m = ms.m;
    const float4* pts = pts_override ? pts_override : ms.ds.as<float4>();
    GridSpec gs;
    const float gate_r = sqrtf(h->prm.knn_gate_sq);
    float cell = cell_override > 0.f ? cell_override : gate_r * (1.0f + 1.0f / 128.0f);
    if (m == 0) {
        gs.ox = gs.oy = gs.oz = 0.f;
        gs.cell = cell;
        gs.inv = 1.0f / cell;
        gs.dx = gs.dy = gs.dz = 1;
        ms.hashed = false;
        CK(ms.cell_start.reserve(2 * 4));
        CK(cudaMemsetAsync(ms.cell_start.p, 0, 8, L.st));
        CK(ms.cell_pts.reserve(16));
        ms.gs = gs;
        return LVREG_OK;
    }
    float mn[3], mx[3];
    if (bb_min && bb_max) {
        // bounding box of the cloud the map was down-sampled from: every centroid lies inside it
        // (cell coordinates are clamped, so a last-bit excursion is harmless)
        for (int a = 0; a < 3; ++a) { mn[a] = bb_min[a]; mx[a] = bb_max[a]; }
    } else {
        uint32_t* mm = L.small.as<uint32_t>() + SM_MM;
        CK(cudaMemsetAsync(mm, 0xff, 3 * sizeof(uint32_t), L.st));
        CK(cudaMemsetAsync(mm + 3, 0, 3 * sizeof(uint32_t), L.st));
        minmax_kernel<<<min(nblk(m, 256), (uint32_t)h->num_sms * 8), 256, 0, L.st>>>(pts, m, mm);
        launched(h);
        CK(cudaMemcpyAsync(L.pinned, mm, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
        CK(cudaStreamSynchronize(L.st));
        for (int a = 0; a < 3; ++a) {
            mn[a] = ordered_to_float(L.pinned[a]);
            mx[a] = ordered_to_float(L.pinned[3 + a]);
        }
    }
    // Directory: dense while it is small next to the map (cells <= max(2^24, 16 m): memory stays O(m) above a fixed
    // 64 MB floor), else a hash of the occupied cells.  Only a map wider than 2048 cells per axis gets coarser cells
    // (the fp32 error bound of the cell coordinate assumes dims <= 2048; coarser cells stay exact).
    ms.hashed = false;
    for (;;) {
        const float inv = 1.0f / cell;
        int64_t dims[3];
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            dims[a] = (int64_t)floorf((mx[a] - mn[a]) * inv) + 1;
            if (dims[a] > 2048 || dims[a] < 1) ok = false;
        }
        if (ok) {
            const int64_t nc = dims[0] * dims[1] * dims[2];
            int64_t dense_cap = (int64_t)16 * (int64_t)m;
            if (dense_cap < ((int64_t)1 << 24)) dense_cap = (int64_t)1 << 24;
            if (dense_cap > ((int64_t)1 << 28)) dense_cap = (int64_t)1 << 28;
            if (!allow_hash) dense_cap = (int64_t)1 << 26;
            if (nc <= dense_cap || allow_hash) {
                ms.hashed = nc > dense_cap;
                gs.ox = mn[0]; gs.oy = mn[1]; gs.oz = mn[2];
                gs.cell = cell;
                gs.inv = inv;
                gs.dx = (int)dims[0]; gs.dy = (int)dims[1]; gs.dz = (int)dims[2];
                break;
            }
        }
        cell *= 2.0f;       // coarser cells stay exact, they only add candidates
        if (!(cell < 1e30f)) return fail(h, LVREG_ERR_INVALID, "map extent is not finite");
    }
    CK(ms.cell_pts.reserve((size_t)m * 16));
    CK(L.keys[0].reserve((size_t)m * 4));
    CK(L.vals[0].reserve((size_t)m * 4));
    if (ms.hashed) {
        uint32_t tsize = 1024;
        while (tsize < 2u * m) tsize <<= 1;
        ms.hmask = tsize - 1;
        CK(ms.hkeys.reserve((size_t)tsize * 8));
        CK(L.scan_in.reserve((size_t)(tsize + 9) * 4));
        CK(ms.cell_start.reserve((size_t)(tsize + 9) * 4));
        CK(L.scan_temp.reserve((size_t)(scan_num_tiles(tsize + 1) + 2) * 4));
        CK(cudaMemsetAsync(ms.hkeys.p, 0xff, (size_t)tsize * 8, L.st));
        CK(cudaMemsetAsync(L.scan_in.p, 0, (size_t)(tsize + 1) * 4, L.st));
        hash_count_kernel<<<nblk(m, 256), 256, 0, L.st>>>(pts, m, gs, ms.hkeys.as<unsigned long long>(), ms.hmask,
                                                          L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                                                          L.scan_in.as<uint32_t>());
        launched(h);
        exclusive_scan(CountIn{L.scan_in.as<uint32_t>()}, StartOut{ms.cell_start.as<uint32_t>()}, tsize + 1,
                       L.scan_temp.as<uint32_t>(), L.small.as<uint32_t>() + SM_TOTAL, L.st, &h->call_launches);
        cell_scatter_kernel<<<nblk(m, 256), 256, 0, L.st>>>(pts, L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                                                            ms.cell_start.as<uint32_t>(), m, ms.cell_pts.as<float4>());
        launched(h);
        CK(cudaGetLastError());
        ms.gs = gs;
        return LVREG_OK;
    }
    const uint32_t ncells = (uint32_t)gs.dx * gs.dy * gs.dz;
    CK(L.scan_in.reserve((size_t)(ncells + 9) * 4));
    CK(ms.cell_start.reserve((size_t)(ncells + 9) * 4));
    CK(L.scan_temp.reserve((size_t)(scan_num_tiles(ncells + 1) + 2) * 4));
    // counting sort by cell: count (the atomic also ranks the point inside its cell), scan, scatter
    CK(cudaMemsetAsync(L.scan_in.p, 0, (size_t)(ncells + 1) * 4, L.st));
    if (m <= kGridMidMaxPoints && ncells <= kGridMidMaxCells && h->vg_mid_enabled) {       // small map: one cooperative launch
        uint32_t blocks = nblk(m > ncells ? m : ncells, kGridMidThreads * 4);
        if (blocks > (uint32_t)h->num_sms) blocks = (uint32_t)h->num_sms;
        if (blocks < 1) blocks = 1;
        CK(L.scan_temp.reserve((size_t)(blocks + 2) * 4 + (size_t)(scan_num_tiles(ncells + 1) + 2) * 4));
        const float4* pts_c = pts;
        uint32_t* keys_p = L.keys[0].as<uint32_t>();
        uint32_t* ranks_p = L.vals[0].as<uint32_t>();
        uint32_t* counts_p = L.scan_in.as<uint32_t>();
        uint32_t* chunk_p = L.scan_temp.as<uint32_t>();
        uint32_t* start_p = ms.cell_start.as<uint32_t>();
        float4* out_p = ms.cell_pts.as<float4>();
        uint32_t m_c = m, nc_c = ncells;
        void* kargs[] = {&pts_c, &m_c, &gs, &nc_c, &keys_p, &ranks_p, &counts_p, &chunk_p, &start_p, &out_p};
        CK(cudaLaunchCooperativeKernel((void*)grid_build_mid_kernel, dim3(blocks), dim3(kGridMidThreads), kargs, 0, L.st));
        launched(h);
        CK(cudaGetLastError());
        ms.gs = gs;
        return LVREG_OK;
    }
    cell_count_kernel<<<nblk(m, 256), 256, 0, L.st>>>(pts, m, gs, L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                                                      L.scan_in.as<uint32_t>());
    launched(h);
    exclusive_scan(CountIn{L.scan_in.as<uint32_t>()}, StartOut{ms.cell_start.as<uint32_t>()}, ncells + 1,
                   L.scan_temp.as<uint32_t>(), L.small.as<uint32_t>() + SM_TOTAL, L.st, &h->call_launches);
    cell_scatter_kernel<<<nblk(m, 256), 256, 0, L.st>>>(pts, L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                                                        ms.cell_start.as<uint32_t>(), m, ms.cell_pts.as<float4>());
    launched(h);
    CK(cudaGetLastError());
    ms.gs = gs;
    return LVREG_OK;
}

GridView grid_view(const MapSide& ms) {
    GridView g;
    g.pts = ms.cell_pts.as<float4>();
    g.cell_start = ms.cell_start.as<uint32_t>();
    g.hkeys = ms.hkeys.as<unsigned long long>();
    g.hmask = ms.hmask;
    g.hashed = ms.hashed ? 1 : 0;
    g.ox = ms.gs.ox; g.oy = ms.gs.oy; g.oz = ms.gs.oz;
    g.inv = ms.gs.inv;
    g.cell = ms.gs.cell;
    g.dx = ms.gs.dx; g.dy = ms.gs.dy; g.dz = ms.gs.dz;
    g.m = ms.m;
    return g;
}

RegParams reg_params(const lvreg_handle* h) {
    RegParams P;
    P.knn_gate_sq = h->prm.knn_gate_sq;
    P.line_eig_ratio = h->prm.line_eig_ratio;
    P.plane_tol = h->prm.plane_tol;
    P.min_weight = h->prm.min_weight;
    P.degeneracy_eig = h->prm.degeneracy_eig;
    P.conv_deg = h->prm.conv_deg;
    P.conv_cm = h->prm.conv_cm;
    P.min_matches = h->prm.min_matches;
    P.max_iters = h->prm.max_iters;
    P.reference_quirks = h->prm.reference_quirks;
    return P;
}

void fill_map_info(const lvreg_handle* h, lvreg_map_info* info) {
    if (!info) return;
    memset(info, 0, sizeof(*info));
    info->n_corner_in = h->map[0].n_in;
    info->n_surf_in = h->map[1].n_in;
    info->n_corner_ds = h->map[0].m;
    info->n_surf_ds = h->map[1].m;
    for (int s = 0; s < 2; ++s) {
        info->grid_dims[s][0] = h->map[s].gs.dx;
        info->grid_dims[s][1] = h->map[s].gs.dy;
        info->grid_dims[s][2] = h->map[s].gs.dz;
        info->grid_cell[s] = h->map[s].gs.cell;
    }
}

// host restatement of pcl::getTransformation (PCL common/impl/eigen.hpp), fp32, libm sinf/cosf:
// computed on the host so that the kernels receive bit-identical matrices
void pose_to_affine_host(const float pose[6], float T[12]) {
    const float roll = pose[0], pitch = pose[1], yaw = pose[2];
    float A = cosf(yaw), B = sinf(yaw), C = cosf(pitch), D = sinf(pitch), E = cosf(roll), F = sinf(roll);
    float DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2]  = B * F + A * DE;  T[3]  = pose[3];
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6]  = B * DE - A * F;  T[7]  = pose[4];
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = pose[5];
}

// PCL's VoxelGrid bounds of a cloud with the given bounding box; false when PCL would pass the cloud through (index
// overflow) or the voxel coordinates leave the range in which the fp32 arithmetic of voxel_key is exact
bool voxel_spec_from_bbox(const float* mn, const float* mx, float leaf, VoxelSpec* vs) {
    const float inv = 1.0f / leaf;
    int64_t d[3];
    for (int a = 0; a < 3; ++a) {
        if (!(fabsf(mn[a]) * inv < 4194304.f) || !(fabsf(mx[a]) * inv < 4194304.f)) return false;
        d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    }
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) return false;
    vs->inv = inv;
    int div_b[3];
    for (int a = 0; a < 3; ++a) {
        vs->min_b[a] = (int)floorf(mn[a] * inv);
        const int max_b = (int)floorf(mx[a] * inv);
        div_b[a] = max_b - vs->min_b[a] + 1;
    }
    vs->mul[0] = 1;
    vs->mul[1] = div_b[0];
    vs->mul[2] = div_b[0] * div_b[1];
    vs->key_bits = bits_for((uint64_t)div_b[0] * div_b[1] * div_b[2] - 1);
    return true;
}

// laserCloudMapContainer (MO:942-954): transforms the not-yet-cached clouds of the listed keyframes under their
// current poses and measures their bounding boxes -- one batch of launches and ONE host synchronisation for all of
// them; a keyframe is transformed once per pose, as in the reference.  Runs on the main stream, before the lanes fork.
// With the bucketed VoxelGrid enabled the cached cloud is additionally ORDERED by its voxel index under the local
// map's leaf size (stable: points of one voxel keep their scan order, so the sort-based filter still gives the same
// sums on it); see voxelgrid_bucket.cuh.
int ensure_world_cache(lvreg_handle* h, const int32_t* ids, size_t n_ids) {
    const float leaf[2] = {h->prm.corner_leaf, h->prm.surf_leaf};
    const bool want_sorted = h->vg_bucket_enabled;
    std::vector<Keyframe*> todo;
    for (size_t i = 0; i < n_ids; ++i) {
        Keyframe* kf = h->kfs[ids[i]];
        bool stale = !kf->world_ok;
        for (int s = 0; s < 2; ++s)
            if (kf->n[s] && kf->wsorted[s] && kf->wleaf[s] != leaf[s]) stale = true;     // ordered for another leaf size
        if (stale && std::find(todo.begin(), todo.end(), kf) == todo.end()) todo.push_back(kf);
    }
    const size_t batch = 1024;                                   // 1024 keyframes x 2 clouds x 6 words = 48 KB of pinned memory
    uint32_t* host_mm = (uint32_t*)((char*)h->pinned + 8192);
    Lane& L = h->lane[LANE_MAP_SURF];                            // its sort scratch is idle: the lanes have not forked yet
    for (size_t b0 = 0; b0 < todo.size(); b0 += batch) {
        const size_t nb = std::min(batch, todo.size() - b0);
        CK(h->kfmm.reserve(nb * 2 * 6 * 4));
        bbox_slots_init_kernel<<<nblk((uint32_t)(nb * 12), 256), 256, 0, h->st>>>(h->kfmm.as<uint32_t>(), (uint32_t)(nb * 2));
        launched(h);
        for (size_t k = 0; k < nb; ++k) {
            Keyframe* kf = todo[b0 + k];
            Affine T;
            pose_to_affine_host(kf->pose, T.m);                  // pclPointToAffine3f(cloudKeyPoses6D[id])
            for (int s = 0; s < 2; ++s) {
                if (kf->n[s] == 0) continue;
                CK(kf->world[s].reserve((size_t)kf->n[s] * 16));
                const uint32_t blocks = min(nblk(kf->n[s], 256), (uint32_t)h->num_sms * 4);
                uint32_t* mm = h->kfmm.as<uint32_t>() + (k * 2 + s) * 6;
                if (want_sorted)
                    bbox_tf_kernel<<<blocks, 256, 0, h->st>>>(kf->cloud[s].as<float4>(), kf->n[s], T, mm);
                else
                    transform_bbox_kernel<<<blocks, 256, 0, h->st>>>(kf->cloud[s].as<float4>(), kf->n[s], T,
                                                                     kf->world[s].as<float4>(), mm);
                launched(h);
            }
        }
        CK(cudaMemcpyAsync(host_mm, h->kfmm.p, nb * 2 * 6 * 4, cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        for (size_t k = 0; k < nb; ++k) {
            Keyframe* kf = todo[b0 + k];
            for (int s = 0; s < 2; ++s) {
                for (int a = 0; a < 3; ++a) {
                    kf->wmn[s][a] = ordered_to_float(host_mm[(k * 2 + s) * 6 + a]);
                    kf->wmx[s][a] = ordered_to_float(host_mm[(k * 2 + s) * 6 + 3 + a]);
                }
                kf->wsorted[s] = false;
                kf->wleaf[s] = 0.f;
            }
            if (want_sorted) {
                Affine T;
                pose_to_affine_host(kf->pose, T.m);
                for (int s = 0; s < 2; ++s) {
                    const uint32_t n = kf->n[s];
                    if (n == 0) continue;
                    VoxelSpec vs;
                    const bool ok = leaf[s] > 0.f && voxel_spec_from_bbox(kf->wmn[s], kf->wmx[s], leaf[s], &vs);
                    bool packs = false;                          // the packed voxel coordinates have 11 | 11 | 10 bits
                    if (ok) {
                        const int dxk = vs.mul[1], dyk = vs.mul[1] ? vs.mul[2] / vs.mul[1] : 0;
                        const int dzk = (int)floorf(kf->wmx[s][2] * vs.inv) - vs.min_b[2] + 1;
                        packs = dxk <= (1 << kVgbPackX) && dyk <= (1 << kVgbPackY) && dzk <= (1 << kVgbPackZ);
                    }
                    kf->wpacked[s] = false;
                    if (!ok) {                                   // cannot be ordered with 32-bit keys: cached as it is
                        transform_kernel<<<nblk(n, 256), 256, 0, h->st>>>(kf->cloud[s].as<float4>(), n, T, kf->world[s].as<float4>());
                        launched(h);
                        continue;
                    }
                    CKS(ensure_sort_buffers(h, L, n));
                    voxel_keys_tf_kernel<<<nblk(n, 256), 256, 0, h->st>>>(kf->cloud[s].as<float4>(), n, T, vs,
                                                                          L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>());
                    launched(h);
                    const int cur = radix_sort_pairs(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), L.keys[1].as<uint32_t>(),
                                                     L.vals[1].as<uint32_t>(), n, vs.key_bits, L.sort_scratch.as<uint32_t>(), h->st,
                                                     &h->call_launches);
                    if (packs) CK(kf->wkey[s].reserve((size_t)n * 4));
                    gather_tf_kernel<<<nblk(n, 256), 256, 0, h->st>>>(kf->cloud[s].as<float4>(), L.vals[cur].as<uint32_t>(), n, T, vs,
                                                                      kf->world[s].as<float4>(),
                                                                      packs ? kf->wkey[s].as<uint32_t>() : nullptr);
                    kf->wpacked[s] = packs;
                    launched(h);
                    for (int a = 0; a < 3; ++a) kf->wminb[s][a] = vs.min_b[a];
                    kf->wsorted[s] = true;
                    kf->wleaf[s] = leaf[s];
                }
            }
            kf->world_ok = true;
        }
        CK(cudaGetLastError());
    }
    return LVREG_OK;
}

// fills the two local-map jobs (lanes 0/1) from the keyframe id list (extractCloud MO:931-957)
int prepare_map_jobs(lvreg_handle* h, const int32_t* ids, size_t n_ids, VgJob* jobs) {
    if (h->kfs.empty()) return fail(h, LVREG_ERR_NO_KEYFRAMES, "no keyframes");
    for (size_t i = 0; i < n_ids; ++i)
        if (ids[i] < 0 || (size_t)ids[i] >= h->kfs.size()) return fail(h, LVREG_ERR_INVALID, "keyframe id out of range");
    // which class goes through the cached world-frame clouds: the large ones (the single-launch filter for small maps
    // transforms in-kernel), when the (segment, offset) payload can address them
    bool cached[2] = {false, false};
    bool any_cached = false;
    for (int s = 0; s < 2; ++s) {
        uint64_t total = 0, nseg = 0, biggest = 0;
        for (size_t i = 0; i < n_ids; ++i) {
            const Keyframe* kf = h->kfs[ids[i]];
            if (kf->n[s] == 0) continue;
            total += kf->n[s];
            ++nseg;
            if (kf->n[s] > biggest) biggest = kf->n[s];
        }
        cached[s] = h->kf_cache_enabled && (total > (uint64_t)kVgMidMax || !h->vg_mid_enabled) &&
                    nseg <= (1u << (32 - kSegShift)) && biggest <= (uint64_t)kSegOffMask;
        any_cached = any_cached || cached[s];
    }
    if (any_cached) CKS(ensure_world_cache(h, ids, n_ids));
    for (int s = 0; s < 2; ++s) {
        Lane& L = h->lane[s];
        MapSide& ms = h->map[s];
        L.seg_host.clear();
        uint64_t total = 0;
        float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        for (size_t i = 0; i < n_ids; ++i) {
            const Keyframe* kf = h->kfs[ids[i]];
            if (kf->n[s] == 0) continue;          // empty clouds contribute nothing
            Segment sg;
            sg.src = cached[s] ? kf->world[s].as<float4>() : kf->cloud[s].as<float4>();
            sg.begin = (uint32_t)total;
            sg.n = kf->n[s];
            sg.wkey = cached[s] && kf->wsorted[s] && kf->wpacked[s] && !h->dbg_nowkey ? kf->wkey[s].as<uint32_t>() : nullptr;
            for (int a = 0; a < 3; ++a) sg.kminb[a] = kf->wminb[s][a];
            if (cached[s]) memset(&sg.T, 0, sizeof(sg.T));   // world-frame cloud: no transform
            else pose_to_affine_host(kf->pose, sg.T.m);      // pclPointToAffine3f(cloudKeyPoses6D[id])
            L.seg_host.push_back(sg);
            total += kf->n[s];
            for (int a = 0; a < 3; ++a) {
                mn[a] = fminf(mn[a], kf->wmn[s][a]);
                mx[a] = fmaxf(mx[a], kf->wmx[s][a]);
            }
        }
        if (total > 0x7fffffffull) return fail(h, LVREG_ERR_INVALID, "local map too large");
        ms.n_in = total;
        ms.valid = false;
        VgJob& J = jobs[s];
        J = VgJob();
        J.lane = s;
        J.n = (uint32_t)total;
        J.from_segments = true;
        J.cached = cached[s] && total > 0;
        if (J.cached)
            for (int a = 0; a < 3; ++a) { J.mn[a] = mn[a]; J.mx[a] = mx[a]; }
        J.leaf = s == 0 ? h->prm.corner_leaf : h->prm.surf_leaf;
        J.bucket = J.cached && h->vg_bucket_enabled;
        for (size_t i = 0; i < n_ids && J.bucket; ++i) {
            const Keyframe* kf = h->kfs[ids[i]];
            if (kf->n[s] && !(kf->wsorted[s] && kf->wleaf[s] == J.leaf)) J.bucket = false;
        }
        J.out = &ms.ds;
        J.n_out = &ms.m;
        J.mark_map_stage = true;
    }
    return LVREG_OK;
}

int reg_variant(const lvreg_handle* h);

// uploads the raw scan clouds on lanes 2/3 and fills their jobs (downsampleCurrentScan MO:987-999)
int prepare_scan_jobs(lvreg_handle* h, const lvreg_cloud* corner_raw, const lvreg_cloud* surf_raw, VgJob* jobs) {
    const lvreg_cloud* c[2] = {corner_raw, surf_raw};
    for (int s = 0; s < 2; ++s) {
        Lane& L = h->lane[LANE_SCAN_CORNER + s];
        CKS(upload_cloud(h, c[s], L.raw, L.stage, L.st));
        VgJob& J = jobs[s];
        J = VgJob();
        J.lane = LANE_SCAN_CORNER + s;
        J.pts = L.raw.as<float4>();
        J.n = (uint32_t)c[s]->n;
        J.leaf = s == 0 ? h->prm.corner_leaf : h->prm.surf_leaf;
        J.out = &h->scan_ds[s];
        J.n_out = &h->n_scan[s];
    }
    // large scans: the registration kernels want the queries in Morton order (want_sorted_scan); the single-launch
    // filter produces that copy as its last phase, before any host synchronisation
    if (reg_variant(h) >= 2 && c[0]->n + c[1]->n >= 65536)
        for (int s = 0; s < 2; ++s) jobs[s].morton_out = &h->scan_sorted[s];
    return LVREG_OK;
}

// search grids of both maps on lanes 0/1, from the bounding boxes the VoxelGrid jobs measured
int build_map_grids(lvreg_handle* h, VgJob* map_jobs) {
    for (int s = 0; s < 2; ++s) {
        const bool have_bb = map_jobs && map_jobs[s].n > 0;
        MapSide& ms = h->map[s];
        CKS(build_grid(h, h->lane[s], ms, have_bb ? map_jobs[s].mn : nullptr, have_bb ? map_jobs[s].mx : nullptr,
                       0.f, nullptr, true));
        ms.valid = true;
    }
    return LVREG_OK;
}

// Morton-orders laserCloud{Corner,Surf}LastDS for the registration kernels (knn.cuh) on stream `st`, with the
// scratch of the scan's lane.  The down-sampled cloud itself keeps PCL's order (it is API-visible and becomes
// a keyframe).
int sort_scan_for_search(lvreg_handle* h, int s, cudaStream_t st) {
    Lane& L = h->lane[LANE_SCAN_CORNER + s];
    const uint32_t n = h->n_scan[s];
    h->scan_sorted_ok[s] = false;
    CK(h->scan_sorted[s].reserve((size_t)(n ? n : 1) * 16));
    if (n == 0) { h->scan_sorted_ok[s] = true; return LVREG_OK; }
    CKS(ensure_sort_buffers(h, L, n));
    morton_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(h->scan_ds[s].as<float4>(), n, L.keys[0].as<uint32_t>(),
                                                     L.vals[0].as<uint32_t>());
    launched(h);
    const int cur = radix_sort_pairs(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), L.keys[1].as<uint32_t>(),
                                     L.vals[1].as<uint32_t>(), n, 24, L.sort_scratch.as<uint32_t>(), st, &h->call_launches);
    gather_points_kernel<<<nblk(n, 256), 256, 0, st>>>(h->scan_ds[s].as<float4>(), L.vals[cur].as<uint32_t>(), n,
                                                       h->scan_sorted[s].as<float4>());
    launched(h);
    CK(cudaGetLastError());
    h->scan_sorted_ok[s] = true;
    return LVREG_OK;
}

float clampf(float v, float lim) {
    if (v < -lim) v = -lim;
    if (v > lim) v = lim;
    return v;
}

// transformUpdate (MO:1345-1375) in the reference's order: the IMU blend of roll and pitch first, then ONE clamp of
// roll, pitch and z (constraintTransformation, MO:1377-1385)
void transform_update_host(const lvreg_params& prm, float pose[6], int imu_available, float imu_roll, float imu_pitch) {
    if (imu_available && fabs((double)imu_pitch) < 1.4) {
        // tf2 slerp of two rotations about one axis == interpolation of the angle along the
        // shortest arc (MO:1349-1366), evaluated in double like tf2
        const double w = prm.imu_rpy_weight;
        auto slerp_angle = [](double a, double b, double t) {
            double d = remainder(b - a, 2.0 * M_PI);
            return remainder(a + t * d, 2.0 * M_PI);
        };
        pose[0] = (float)slerp_angle(pose[0], imu_roll, w);
        pose[1] = (float)slerp_angle(pose[1], imu_pitch, w);
    }
    pose[0] = clampf(pose[0], prm.rotation_tolerance);
    pose[1] = clampf(pose[1], prm.rotation_tolerance);
    pose[5] = clampf(pose[5], prm.z_tolerance);
}

template <int LPQ, int TILE>
int launch_register(lvreg_handle* h, RegArgs& args, int grid) {
    void* kargs[] = {&args};
    CK(cudaLaunchCooperativeKernel((void*)register_kernel<LPQ, TILE>, dim3(grid), dim3(kRegThreads), kargs, 0, h->st));
    return LVREG_OK;
}

template <int LPQ, int TILE>
int reg_occupancy(lvreg_handle* h, int* nb) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, register_kernel<LPQ, TILE>, kRegThreads, 0));
    return LVREG_OK;
}

// (lanes per query, queries per warp tile) variants compiled in; selected by LVREG_LPQ / LVREG_TILE
#define LVREG_REG_DISPATCH(FN, ...)                                             \
    do {                                                                        \
        const int key_ = h->lpq * 100 + h->tile;                                \
        switch (key_) {                                                         \
            case 816: CKS((FN<8, 16>(__VA_ARGS__))); break;                     \
            default: CKS((FN<4, 16>(__VA_ARGS__))); break;                      \
        }                                                                       \
    } while (0)

// Morton-ordering the queries pays when there are several tiles per SM to balance (C3: 2200 tiles); on a small
// scan its ~16 extra launches cost more than the ordering returns (C1: 240 tiles, launch-bound).  The staged
// search always needs compact tiles.
int reg_variant(const lvreg_handle* h);
bool want_sorted_scan(const lvreg_handle* h, int variant) {
    if (variant == 2) return true;
    return variant == 3 && h->n_scan[0] + h->n_scan[1] >= 32768u;
}

// kernel variant: 3 = warm (thread per query, warm-started search radius, static tiles; the default),
// 2 = staged (shared-memory search), 1 = thread per query with dynamic tiles, 0 = lane groups.
// LVREG_REG / LVREG_TPQ override the choice (tests, experiments).
int reg_variant(const lvreg_handle* h) {
    int variant = 3;
    if (h->force_tpq >= 0) variant = h->force_tpq;
    if (h->reg_variant_env >= 0) variant = h->reg_variant_env;
    return variant;
}

int scan2map_impl(lvreg_handle* h, float pose[6], lvreg_result* res) {
    if (res) {
        memset(res, 0, sizeof(*res));
        res->n_corner_ds = (int)h->n_scan[0];
        res->n_surf_ds = (int)h->n_scan[1];
        res->n_corner_map = (int)h->map[0].m;
        res->n_surf_map = (int)h->map[1].m;
    }
    if (!h->map[0].valid || !h->map[1].valid)
        return fail(h, h->kfs.empty() ? LVREG_ERR_NO_KEYFRAMES : LVREG_ERR_NO_MAP, "no local map");
    if (!((int)h->n_scan[0] > h->prm.edge_min_valid && (int)h->n_scan[1] > h->prm.surf_min_valid)) {
        if (res) { /* isDegenerate keeps its previous value (MO:131) */
            LmState st;
            cudaMemcpyAsync(&st, h->lmstate.p, sizeof(int), cudaMemcpyDeviceToHost, h->st);
            cudaStreamSynchronize(h->st);
            res->degenerate = st.is_degenerate;
        }
        return fail(h, LVREG_ERR_NOT_ENOUGH_FEATURES, "not enough features");
    }
    float* hp = (float*)((char*)h->pinned + 2048);
    memcpy(hp, pose, 6 * sizeof(float));
    CK(cudaMemcpyAsync(h->posebuf.p, hp, 6 * sizeof(float), cudaMemcpyHostToDevice, h->st));

    RegArgs args;
    for (int s = 0; s < 2; ++s) {
        args.grid[s] = grid_view(h->map[s]);
        args.map[s] = h->map[s].ds.as<float4>();
        args.scan[s] = h->scan_ds[s].as<float4>();
        args.n[s] = h->n_scan[s];
    }
    args.dealt = h->dbg_dealt ? 1 : 0;
    if (want_sorted_scan(h, reg_variant(h)) && !args.dealt)     // spatially compact tiles: similar paths, shared cache lines
        for (int s = 0; s < 2; ++s) {
            if (!h->scan_sorted_ok[s]) CKS(sort_scan_for_search(h, s, h->st));
            args.scan[s] = h->scan_sorted[s].as<float4>();
        }
    args.prm = reg_params(h);
    args.pose_in = h->posebuf.as<float>();
    args.out = h->regout.as<RegOut>();
    args.lm = h->lmstate.as<LmState>();

    const int variant = reg_variant(h);
    if (variant != h->reg_occ_variant) h->reg_max_blocks_per_sm = 0;
    h->reg_occ_variant = variant;
    if (h->reg_max_blocks_per_sm == 0) {
        int nb = 0;
        if (variant == 3) {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, register_warm_kernel_ptr(h->warm_tile), kRegThreads, 0));
        } else if (variant == 2) {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, register_staged_kernel, kRegThreads, register_staged_smem_bytes()));
        } else if (variant == 1) {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, register_tpq_kernel, kRegThreads, 0));
        } else {
            LVREG_REG_DISPATCH(reg_occupancy, h, &nb);
        }
        if (nb < 1) return fail(h, LVREG_ERR_CUDA, "the registration kernel does not fit on an SM");
        h->reg_max_blocks_per_sm = nb;
    }
    const int tile_q = variant == 0 ? h->tile : (variant == 3 ? h->warm_tile : 32);
    const uint32_t tiles = nblk(h->n_scan[0], tile_q) + nblk(h->n_scan[1], tile_q);
    const int max_grid = h->reg_max_blocks_per_sm * h->num_sms;
    // static tiles: tile t runs on block t % grid, so a small scan spreads over all SMs, one tile per warp
    int grid = variant >= 2 ? (int)nblk(tiles, (uint32_t)h->reg_tiles_per_block) : (int)nblk(tiles, kRegWarps);
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    CK(h->partials.reserve((size_t)2 * grid * kRegTerms * sizeof(double)));
    args.partials = h->partials.as<double>();

    CK(cudaMemsetAsync(h->tilectr.p, 0, LVREG_MAX_ITERS * sizeof(uint32_t), h->st));
    args.tile_counter = h->tilectr.as<uint32_t>();
    args.tile_ns = nullptr;
    if (h->debug_tiles) {
        CK(h->tilens.reserve((size_t)(tiles + 1) * 4));
        CK(cudaMemsetAsync(h->tilens.p, 0, (size_t)(tiles + 1) * 4, h->st));
        args.tile_ns = h->tilens.as<uint32_t>();
        h->debug_ntiles = tiles;
    }
    args.stage_stats = nullptr;
    args.nn_prev[0] = args.nn_prev[1] = nullptr;
    if (variant == 3) {
        for (int s = 0; s < 2; ++s) {
            CK(h->nnprev[s].reserve((size_t)(nblk(h->n_scan[s], 32) + 1) * 32 * 5 * sizeof(int32_t)));
            args.nn_prev[s] = h->nnprev[s].as<int32_t>();
        }
        void* kargs[] = {&args};
        CK(cudaLaunchCooperativeKernel(register_warm_kernel_ptr(h->warm_tile), dim3(grid), dim3(kRegThreads), kargs, 0, h->st));
    } else if (variant == 2) {
        if (h->debug_tiles) {
            CK(cudaMemsetAsync(h->stagestats.p, 0, 8 * sizeof(uint32_t), h->st));
            args.stage_stats = h->stagestats.as<uint32_t>();
        }
        void* kargs[] = {&args};
        CK(cudaLaunchCooperativeKernel((void*)register_staged_kernel, dim3(grid), dim3(kRegThreads), kargs,
                                       register_staged_smem_bytes(), h->st));
    } else if (variant == 1) {
        void* kargs[] = {&args};
        CK(cudaLaunchCooperativeKernel((void*)register_tpq_kernel, dim3(grid), dim3(kRegThreads), kargs, 0, h->st));
    } else {
        LVREG_REG_DISPATCH(launch_register, h, args, grid);
    }
    launched(h);
    RegOut* ho = (RegOut*)((char*)h->pinned + 4096);
    CK(cudaMemcpyAsync(ho, h->regout.p, sizeof(RegOut), cudaMemcpyDeviceToHost, h->st));
    mark(h, EV_REG);
    CK(cudaStreamSynchronize(h->st));
    for (int i = 0; i < 6; ++i) pose[i] = ho->pose[i];
    // transformUpdate (MO:1339): IMU blend (lvreg_set_imu_prior) first, then the clamps, as the reference orders them
    transform_update_host(h->prm, pose, h->imu_available, h->imu_roll, h->imu_pitch);
    if (res) {
        res->iterations = ho->iterations;
        res->converged = ho->converged;
        res->degenerate = ho->degenerate;
        const int k = h->prm.max_iters < LVREG_MAX_ITERS ? h->prm.max_iters : LVREG_MAX_ITERS;
        for (int i = 0; i < k; ++i) {
            res->n_sel[i] = i < ho->iterations ? ho->n_sel[i] : 0;
            res->cost[i] = i < ho->iterations ? ho->cost[i] : 0.f;
            for (int j = 0; j < 6; ++j) res->pose_iter[i][j] = i < ho->iterations ? ho->pose_iter[i][j] : 0.f;
        }
    }
    return LVREG_OK;
}

void finish_timings(lvreg_handle* h) {
    // stages are recorded in call order; absent stages read 0
    int last = EV_BEGIN;
    for (int i = 0; i < EV_COUNT; ++i) if (h->ev_set[i]) last = i;
    h->last.total_ms = span(h, EV_BEGIN, last);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

void lvreg_default_params(lvreg_params* p) {
    memset(p, 0, sizeof(*p));
    p->corner_leaf = 0.2f;
    p->surf_leaf = 0.4f;
    p->edge_min_valid = 10;
    p->surf_min_valid = 100;
    p->max_iters = 20;
    p->knn_gate_sq = 1.0f;
    p->line_eig_ratio = 3.0f;
    p->plane_tol = 0.2f;
    p->min_weight = 0.1f;
    p->min_matches = 50;
    p->degeneracy_eig = 100.0f;
    p->conv_deg = 0.05f;
    p->conv_cm = 0.05f;
    p->reference_quirks = 1;
    p->rotation_tolerance = 1000.0f;
    p->z_tolerance = 1000.0f;
    p->imu_rpy_weight = 0.01f;
}

int lvreg_version(void) { return 100; }

const char* lvreg_status_string(int s) {
    switch (s) {
        case LVREG_OK: return "ok";
        case LVREG_ERR_INVALID: return "invalid argument";
        case LVREG_ERR_CUDA: return "CUDA error";
        case LVREG_ERR_NOT_ENOUGH_FEATURES: return "not enough features";
        case LVREG_ERR_NO_KEYFRAMES: return "no keyframes";
        case LVREG_ERR_NO_MAP: return "no local map";
        case LVREG_ERR_CAPACITY: return "output buffer too small";
        default: return "unknown status";
    }
}

const char* lvreg_last_error(const lvreg_handle* h) { return h ? h->err.c_str() : "null handle"; }

void* lvreg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void lvreg_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int lvreg_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return LVREG_ERR_INVALID;
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? LVREG_OK : LVREG_ERR_CUDA;
}
int lvreg_host_unregister(void* p) {
    if (!p) return LVREG_ERR_INVALID;
    return cudaHostUnregister(p) == cudaSuccess ? LVREG_OK : LVREG_ERR_CUDA;
}

int lvreg_create(const lvreg_params* p, int device, void* cuda_stream, lvreg_handle** out) {
    if (!out) return LVREG_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
        return LVREG_ERR_CUDA;                      // no CPU fallback
    lvreg_handle* h = new lvreg_handle();
    if (p) h->prm = *p; else lvreg_default_params(&h->prm);
    if (h->prm.max_iters < 1) h->prm.max_iters = 1;
    if (h->prm.max_iters > LVREG_MAX_ITERS) h->prm.max_iters = LVREG_MAX_ITERS;
    h->device = device;
    auto bail = [&](int code) { lvreg_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(LVREG_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(LVREG_ERR_CUDA);
    if (!prop.cooperativeLaunch) return bail(LVREG_ERR_CUDA);
    h->num_sms = prop.multiProcessorCount;
    if (cuda_stream) {
        h->st = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) return bail(LVREG_ERR_CUDA);
        h->own_stream = true;
    }
    for (int i = 0; i < EV_COUNT; ++i) {
        h->ev[i] = nullptr;
        if (cudaEventCreate(&h->ev[i]) != cudaSuccess) return bail(LVREG_ERR_CUDA);
        h->ev_set[i] = false;
    }
    if (cudaMallocHost(&h->pinned, 65536) != cudaSuccess) return bail(LVREG_ERR_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess) return bail(LVREG_ERR_CUDA);
    for (int l = 0; l < kLanes; ++l) {
        Lane& L = h->lane[l];
        if (cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&L.ev, cudaEventDisableTiming) != cudaSuccess ||
            L.small.reserve(SM_WORDS * 4) != cudaSuccess)
            return bail(LVREG_ERR_CUDA);
        L.pinned = (uint32_t*)h->pinned + 64 * l;
        cudaMemsetAsync(L.small.p, 0, SM_WORDS * 4, h->st);
    }
    if (h->regout.reserve(sizeof(RegOut)) != cudaSuccess ||
        h->lmstate.reserve(sizeof(LmState)) != cudaSuccess || h->posebuf.reserve(256) != cudaSuccess ||
        h->tilectr.reserve(LVREG_MAX_ITERS * sizeof(uint32_t)) != cudaSuccess ||
        h->stagestats.reserve(8 * sizeof(uint32_t)) != cudaSuccess)
        return bail(LVREG_ERR_CUDA);
    cudaMemsetAsync(h->stagestats.p, 0, 8 * sizeof(uint32_t), h->st);
    // the staged registration kernel keeps a 12.5 KB candidate tile per warp
    if (cudaFuncSetAttribute(register_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)register_staged_smem_bytes()) != cudaSuccess)
        return bail(LVREG_ERR_CUDA);
    cudaFuncSetAttribute(register_staged_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (cudaFuncSetAttribute(knn5_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)register_staged_smem_bytes()) != cudaSuccess)
        return bail(LVREG_ERR_CUDA);
    cudaFuncSetAttribute(knn5_staged_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaMemsetAsync(h->lmstate.p, 0, sizeof(LmState), h->st);
    // the sort pass keeps 42 KB of staging per block: ask for the large shared-memory carveout
    cudaFuncSetAttribute(vgb_bucket_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(vgb_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vgb_smem_bytes(kVgbMaxSegs));
    cudaFuncSetAttribute(rs_onesweep_kernel<kSortThreads>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(rs_onesweep_kernel<kSortThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem_bytes(kSortThreads));
    cudaFuncSetAttribute(rs_onesweep_kernel<kSortThreadsBig>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(rs_onesweep_kernel<kSortThreadsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem_bytes(kSortThreadsBig));
    const char* e = getenv("LVREG_LPQ");
    if (e) {
        int v = atoi(e);
        if (v == 4 || v == 8 || v == 16 || v == 32) h->lpq = v;
    }
    e = getenv("LVREG_DEBUG_PHASES");
    if (e && atoi(e)) {
        h->debug_phases = 1;
        for (int l = 0; l < kLanes; ++l)
            for (int k = 0; k < 6; ++k) cudaEventCreate(&h->dbg_ev[l][k]);
    }
    e = getenv("LVREG_KF_CACHE");
    if (e) h->kf_cache_enabled = atoi(e) != 0;
    e = getenv("LVREG_VG_MID");
    if (e) h->vg_mid_enabled = atoi(e) != 0;
    e = getenv("LVREG_VG_BUCKET");
    if (e) h->vg_bucket_enabled = atoi(e) != 0;
    h->dbg_nowkey = getenv("LVREG_DEBUG_NOWKEY") != nullptr;
    h->dbg_dealt = getenv("LVREG_DEALT") != nullptr;
    h->dbg_alloc = getenv("LVREG_DEBUG_ALLOC") != nullptr;
    e = getenv("LVREG_VG_CACHED_FIRST");
    if (e) h->vg_cached_first = atoi(e) != 0;
    e = getenv("LVREG_VG_BUCKET_CAP");
    if (e && atoi(e) > 0) h->vg_bucket_cap = (uint32_t)atoi(e);
    e = getenv("LVREG_DEBUG_TILES");
    if (e) h->debug_tiles = atoi(e);
    e = getenv("LVREG_TPQ");
    if (e) h->force_tpq = atoi(e) ? 1 : 0;
    e = getenv("LVREG_REG_TPB");
    if (e && atoi(e) >= 1 && atoi(e) <= kRegWarps) h->reg_tiles_per_block = atoi(e);
    e = getenv("LVREG_REG");
    if (e) {
        if (!strcmp(e, "staged")) h->reg_variant_env = 2;
        else if (!strcmp(e, "warm")) h->reg_variant_env = 3;
        else if (!strcmp(e, "tpq")) h->reg_variant_env = 1;
        else if (!strcmp(e, "grouped")) h->reg_variant_env = 0;
    }
    e = getenv("LVREG_TILE");
    if (e) {
        int v = atoi(e);
        if (v == 16) h->tile = v;
    }
    if (cudaStreamSynchronize(h->st) != cudaSuccess) return bail(LVREG_ERR_CUDA);
    *out = h;
    return LVREG_OK;
}

void lvreg_destroy(lvreg_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    for (Keyframe* kf : h->kfs) {
        kf->cloud[0].release();
        kf->cloud[1].release();
        delete kf;
    }
    for (Keyframe* kf : h->kf_free) {
        kf->cloud[0].release();
        kf->cloud[1].release();
        delete kf;
    }
    for (int l = 0; l < kLanes; ++l) {
        Lane& L = h->lane[l];
        if (L.st) cudaStreamSynchronize(L.st);
        DevBuf* lb[] = {&L.stage, &L.raw, &L.concat, &L.keys[0], &L.keys[1], &L.vals[0], &L.vals[1], &L.sort_scratch,
                        &L.scan_temp, &L.scan_in, &L.vox_start, &L.vox_keys, &L.segs, &L.small, &L.bsoff, &L.bstatus, &L.bout};
        for (DevBuf* b : lb) b->release();
        if (L.seg_pin) cudaFreeHost(L.seg_pin);
        if (L.ev) cudaEventDestroy(L.ev);
        if (L.st) cudaStreamDestroy(L.st);
    }
    for (int s = 0; s < 2; ++s) {
        h->map[s].ds.release(); h->map[s].cell_pts.release(); h->map[s].cell_start.release(); h->map[s].hkeys.release();
        h->scan_ds[s].release();
        h->scan_sorted[s].release();
        h->icp_cloud[s].ds.release(); h->icp_cloud[s].cell_pts.release(); h->icp_cloud[s].cell_start.release();
    }
    h->icp_coarse.cell_pts.release(); h->icp_coarse.cell_start.release();
    for (DepthEntry* e : h->depth_queue) { e->pts.release(); delete e; }
    for (DepthEntry* e : h->depth_free) { e->pts.release(); delete e; }
    {
        DevBuf* pb[] = {&h->proj_raw, &h->proj_imu, &h->proj_pts, &h->proj_rangein, &h->proj_colin, &h->proj_small,
                        &h->proj_owner, &h->proj_cloud, &h->proj_range, &h->proj_col, &h->proj_rings};
        for (DevBuf* b : pb) b->release();
    }
    {
        DevBuf* db[] = {&h->depth_cloud, &h->depth_concat, &h->depth_bins, &h->depth_local, &h->depth_unit, &h->depth_feat,
                        &h->depth_out, &h->depth_f3d};
        for (DevBuf* b : db) b->release();
    }
    DevBuf* bufs[] = {&h->icp_cur, &h->icp_partials, &h->icp_state, &h->icp_idx, &h->icp_d2, &h->feat_pts, &h->feat_range, &h->feat_col, &h->feat_rings, &h->feat_curv, &h->feat_picked,
                      &h->feat_label, &h->feat_flag, &h->feat_ringof, &h->feat_cidx, &h->feat_ccnt, &h->feat_pos,
                      &h->feat_cand, &h->feat_spec, &h->feat_idx, &h->feat_pidx, &h->feat_corner, &h->feat_surf,
                      &h->kfmm, &h->stagestats, &h->nnprev[0], &h->nnprev[1], &h->vgout, &h->partials, &h->regout, &h->lmstate, &h->posebuf, &h->tilectr, &h->tilens, &h->qbuf,
                      &h->idxbuf, &h->d2buf, &h->brute_partial, &h->coeffbuf, &h->flagbuf};
    for (DevBuf* b : bufs) b->release();
    h->kf_arena.release_all();
    for (int l = 0; l < 2; ++l)
        for (int k = 0; k < 2; ++k)
            if (h->kt_ev[l][k]) cudaEventDestroy(h->kt_ev[l][k]);
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (int i = 0; i < EV_COUNT; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream && h->st) cudaStreamDestroy(h->st);
    delete h;
}


// Pre-sizes every per-call buffer for the given workload so that no later call allocates device memory
// (cudaMalloc / cudaFree synchronise the device; single calls were measured to stall for 10-100 ms, and
// growth in the middle of a sequence shows up as a latency outlier).  All arguments are upper bounds;
// 0 skips that group.  Without this call the buffers grow geometrically on demand.
int lvreg_reserve(lvreg_handle* h, size_t map_points_corner, size_t map_points_surf, size_t scan_points_corner,
                  size_t scan_points_surf, size_t max_grid_cells) {
    if (!h) return LVREG_ERR_INVALID;
    const size_t lim = 0x7fffffffull;
    if (map_points_corner > lim || map_points_surf > lim || scan_points_corner > lim || scan_points_surf > lim ||
        max_grid_cells > ((size_t)1 << 26))
        return fail(h, LVREG_ERR_INVALID, "reserve: size out of range");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->st));
    const size_t npts[kLanes] = {map_points_corner, map_points_surf, scan_points_corner, scan_points_surf};
    for (int l = 0; l < kLanes; ++l) {
        const size_t n = npts[l];
        if (!n) continue;
        Lane& L = h->lane[l];
        CKS(ensure_sort_buffers(h, L, (uint32_t)n));
        CK(L.vox_start.reserve(n * 4));
        CK(L.stage.reserve(n * 32));
        if (l < 2) {
            CK(L.concat.reserve(n * 16));
            CK(h->map[l].ds.reserve(n * 16));
            CK(h->map[l].cell_pts.reserve(n * 16));
            if (max_grid_cells) {
                CK(L.scan_in.reserve((max_grid_cells + 9) * 4));
                CK(h->map[l].cell_start.reserve((max_grid_cells + 9) * 4));
                CK(L.scan_temp.reserve((size_t)(scan_num_tiles((uint32_t)max_grid_cells + 1) + 2) * 4));
            }
        } else {
            CK(L.raw.reserve(n * 16));
            CK(h->scan_ds[l - 2].reserve(n * 16));
            CK(h->scan_sorted[l - 2].reserve(n * 16));
            CK(h->nnprev[l - 2].reserve((n / 32 + 2) * 32 * 5 * sizeof(int32_t)));
        }
    }
    CK(cudaStreamSynchronize(h->st));
    return LVREG_OK;
}

// ---- keyframes -----------------------------------------------------------------------------------
namespace {
Keyframe* take_keyframe(lvreg_handle* h) {
    if (h->kf_free.empty()) {
        Keyframe* kf = new Keyframe();
        kf->cloud[0].arena = kf->cloud[1].arena = &h->kf_arena;
        kf->world[0].arena = kf->world[1].arena = &h->kf_arena;
        kf->wkey[0].arena = kf->wkey[1].arena = &h->kf_arena;
        return kf;
    }
    Keyframe* kf = h->kf_free.back();
    h->kf_free.pop_back();
    kf->n[0] = kf->n[1] = 0;
    kf->world_ok = false;
    return kf;
}
}  // namespace

int lvreg_add_keyframe(lvreg_handle* h, const lvreg_cloud* corner, const lvreg_cloud* surf,
                       const float pose[6], int32_t* id_out) {
    if (!h || !pose) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    Keyframe* kf = take_keyframe(h);
    int s0 = upload_cloud(h, corner, kf->cloud[0], h->lane[LANE_SCAN_CORNER].stage, h->st);
    int s1 = s0 == LVREG_OK ? upload_cloud(h, surf, kf->cloud[1], h->lane[LANE_SCAN_SURF].stage, h->st) : s0;
    if (s1 != LVREG_OK || cudaStreamSynchronize(h->st) != cudaSuccess) {
        kf->cloud[0].release();
        kf->cloud[1].release();
        delete kf;
        return s1 != LVREG_OK ? s1 : LVREG_ERR_CUDA;
    }
    kf->n[0] = (uint32_t)corner->n;
    kf->n[1] = (uint32_t)surf->n;
    memcpy(kf->pose, pose, sizeof(kf->pose));
    h->kfs.push_back(kf);
    if (id_out) *id_out = (int32_t)h->kfs.size() - 1;
    end_call(h);
    return LVREG_OK;
}

int lvreg_add_keyframe_from_scan(lvreg_handle* h, const float pose[6], int32_t* id_out) {
    if (!h || !pose) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    Keyframe* kf = take_keyframe(h);
    for (int s = 0; s < 2; ++s) {
        const uint32_t n = h->n_scan[s];
        cudaError_t e = kf->cloud[s].reserve((size_t)(n ? n : 1) * 16);
        if (e == cudaSuccess && n)
            e = cudaMemcpyAsync(kf->cloud[s].p, h->scan_ds[s].p, (size_t)n * 16, cudaMemcpyDeviceToDevice, h->st);
        if (e != cudaSuccess) {
            kf->cloud[0].release();
            kf->cloud[1].release();
            delete kf;
            h->err = cudaGetErrorString(e);
            return LVREG_ERR_CUDA;
        }
        kf->n[s] = n;
    }
    // no host synchronisation: every later use of the keyframe is ordered behind these copies on the handle's stream
    // (the lanes fork from it)
    memcpy(kf->pose, pose, sizeof(kf->pose));
    h->kfs.push_back(kf);
    if (id_out) *id_out = (int32_t)h->kfs.size() - 1;
    return LVREG_OK;
}

int lvreg_update_keyframe_poses(lvreg_handle* h, const float* poses, size_t n) {
    if (!h || (!poses && n)) return LVREG_ERR_INVALID;
    if (n > h->kfs.size()) return fail(h, LVREG_ERR_INVALID, "more poses than keyframes");
    for (size_t i = 0; i < n; ++i) {
        memcpy(h->kfs[i]->pose, poses + 6 * i, 6 * sizeof(float));
        h->kfs[i]->world_ok = false;                      // laserCloudMapContainer.clear(), MO:1623
    }
    // the reference drops laserCloudMapContainer here (MO:1623); this library re-transforms on
    // every build, so only the current local map becomes stale
    h->map[0].valid = h->map[1].valid = false;
    return LVREG_OK;
}

int lvreg_num_keyframes(const lvreg_handle* h, size_t* n) {
    if (!h || !n) return LVREG_ERR_INVALID;
    *n = h->kfs.size();
    return LVREG_OK;
}

int lvreg_clear_keyframes(lvreg_handle* h) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->st));
    // no keyframe stays live: hand every cloud back to the arena at once (its slabs are kept and refilled from the
    // start, so replaying session after session on one handle neither allocates nor grows) and recycle the objects
    for (Keyframe* kf : h->kfs) {
        if (h->kf_free.size() < 4096) h->kf_free.push_back(kf);
        else delete kf;
    }
    for (Keyframe* kf : h->kf_free) {
        for (DevBuf* b : {&kf->cloud[0], &kf->cloud[1], &kf->world[0], &kf->world[1], &kf->wkey[0], &kf->wkey[1]}) { b->p = nullptr; b->cap = 0; }
        kf->world_ok = false;
    }
    h->kf_arena.reset();
    h->kfs.clear();
    h->map[0].valid = h->map[1].valid = false;
    return LVREG_OK;
}

// ---- local map -----------------------------------------------------------------------------------
int lvreg_build_local_map(lvreg_handle* h, const int32_t* ids, size_t n, lvreg_map_info* info) {
    if (!h || (!ids && n)) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    VgJob jobs[2];
    CallGuard guard(h, false, true);
    CKS(prepare_map_jobs(h, ids, n, jobs));
    lanes_fork(h, 0x3);
    CKS(voxelgrid_batch(h, jobs, 2));
    lanes_join(h, 0x3);
    if (!h->ev_set[EV_MAP]) mark(h, EV_MAP);
    CKS(build_map_grids(h, jobs));
    lanes_join(h, 0x3);
    mark(h, EV_GRID);
    CK(cudaStreamSynchronize(h->st));
    guard.commit();
    h->last.map_build_ms = span(h, EV_BEGIN, EV_MAP);
    h->last.grid_build_ms = span(h, EV_MAP, EV_GRID);
    finish_timings(h);
    end_call(h);
    fill_map_info(h, info);
    return LVREG_OK;
}

int lvreg_set_local_map(lvreg_handle* h, const lvreg_cloud* corner_ds, const lvreg_cloud* surf_ds,
                        lvreg_map_info* info) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    const lvreg_cloud* c[2] = {corner_ds, surf_ds};
    CallGuard guard(h, false, true);
    lanes_fork(h, 0x3);
    for (int s = 0; s < 2; ++s) {
        Lane& L = h->lane[s];
        h->map[s].valid = false;
        CKS(upload_cloud(h, c[s], h->map[s].ds, L.stage, L.st));
        h->map[s].m = (uint32_t)c[s]->n;
        h->map[s].n_in = c[s]->n;
    }
    lanes_join(h, 0x3);
    if (!h->ev_set[EV_MAP]) mark(h, EV_MAP);
    CKS(build_map_grids(h, nullptr));
    lanes_join(h, 0x3);
    mark(h, EV_GRID);
    CK(cudaStreamSynchronize(h->st));
    guard.commit();
    h->last.upload_ms = span(h, EV_BEGIN, EV_MAP);
    h->last.grid_build_ms = span(h, EV_MAP, EV_GRID);
    finish_timings(h);
    end_call(h);
    fill_map_info(h, info);
    return LVREG_OK;
}

int lvreg_get_local_map(lvreg_handle* h, int which, lvreg_cloud_out* out, size_t* n) {
    if (!h || which < 0 || which > 1) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->map[which].valid) return fail(h, LVREG_ERR_NO_MAP, "no local map");
    if (n) *n = h->map[which].m;
    if (!out) return LVREG_OK;
    return download_cloud(h, h->map[which].ds.as<float4>(), h->map[which].m, out);
}

// ---- current scan --------------------------------------------------------------------------------
int lvreg_downsample_scan(lvreg_handle* h, const lvreg_cloud* corner_raw, const lvreg_cloud* surf_raw,
                          size_t* nc, size_t* ns) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    VgJob jobs[2];
    CallGuard guard(h, true, false);
    lanes_fork(h, 0xc);
    CKS(prepare_scan_jobs(h, corner_raw, surf_raw, jobs));
    CKS(voxelgrid_batch(h, jobs, 2));
    for (int s = 0; s < 2; ++s) h->scan_sorted_ok[s] = jobs[s].morton_done;
    if (want_sorted_scan(h, reg_variant(h)))
        for (int s = 0; s < 2; ++s)
            if (!h->scan_sorted_ok[s]) CKS(sort_scan_for_search(h, s, h->lane[LANE_SCAN_CORNER + s].st));
    lanes_join(h, 0xc);
    mark(h, EV_DS);
    CK(cudaStreamSynchronize(h->st));
    guard.commit();
    h->last.downsample_ms = span(h, EV_BEGIN, EV_DS);      // includes the H2D + pack of the two clouds
    finish_timings(h);
    end_call(h);
    if (nc) *nc = h->n_scan[0];
    if (ns) *ns = h->n_scan[1];
    return LVREG_OK;
}

int lvreg_set_scan_ds(lvreg_handle* h, const lvreg_cloud* corner_ds, const lvreg_cloud* surf_ds) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    CallGuard guard(h, true, false);
    CKS(upload_cloud(h, corner_ds, h->scan_ds[0], h->lane[LANE_SCAN_CORNER].stage, h->st));
    CKS(upload_cloud(h, surf_ds, h->scan_ds[1], h->lane[LANE_SCAN_SURF].stage, h->st));
    CK(cudaStreamSynchronize(h->st));
    guard.commit();
    h->n_scan[0] = (uint32_t)corner_ds->n;
    h->n_scan[1] = (uint32_t)surf_ds->n;
    h->scan_sorted_ok[0] = h->scan_sorted_ok[1] = false;      // ordered by the next lvreg_scan2map
    end_call(h);
    return LVREG_OK;
}

int lvreg_get_scan_ds(lvreg_handle* h, int which, lvreg_cloud_out* out, size_t* n) {
    if (!h || which < 0 || which > 1) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (n) *n = h->n_scan[which];
    if (!out) return LVREG_OK;
    return download_cloud(h, h->scan_ds[which].as<float4>(), h->n_scan[which], out);
}

// ---- registration --------------------------------------------------------------------------------
int lvreg_scan2map(lvreg_handle* h, float pose[6], lvreg_result* res) {
    if (!h || !pose) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    int s = scan2map_impl(h, pose, res);
    if (s == LVREG_OK) {
        h->last.register_ms = span(h, EV_BEGIN, EV_REG);
        finish_timings(h);
    }
    end_call(h);
    return s;
}

int lvreg_register_scan(lvreg_handle* h, const lvreg_cloud* corner_raw, const lvreg_cloud* surf_raw,
                        const int32_t* ids, size_t n_ids, float pose[6], lvreg_result* res) {
    if (!h || !pose) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    if (h->debug_phases) h->dbg_host_us[0] = host_now_us();
    // extractSurroundingKeyFrames (MO:318) and downsampleCurrentScan (MO:320) are independent: all
    // four VoxelGrid filters run concurrently on the lanes and share two host synchronisations
    VgJob jobs[4];
    int nj = 0;
    unsigned mask = 0xc;
    CallGuard guard(h, true, ids != nullptr);
    if (ids) {
        CKS(prepare_map_jobs(h, ids, n_ids, jobs));
        nj = 2;
        mask = 0xf;
    }
    lanes_fork(h, mask);
    CKS(prepare_scan_jobs(h, corner_raw, surf_raw, jobs + nj));
    nj += 2;
    if (h->debug_phases) h->dbg_host_us[1] = host_now_us();
    CKS(voxelgrid_batch(h, jobs, nj));
    if (h->debug_phases) h->dbg_host_us[4] = host_now_us();
    // large scans are Morton-ordered for the search: normally by the single-launch filter itself (morton_done)
    for (int s = 0; s < 2; ++s) h->scan_sorted_ok[s] = jobs[nj - 2 + s].morton_done;
    if (want_sorted_scan(h, reg_variant(h)))
        for (int s = 0; s < 2; ++s)
            if (!h->scan_sorted_ok[s]) CKS(sort_scan_for_search(h, s, h->lane[LANE_SCAN_CORNER + s].st));
    if (h->debug_phases)
        for (int l = 0; l < kLanes; ++l) cudaEventRecord(h->dbg_ev[l][5], h->lane[l].st);
    lanes_join(h, mask);
    if (!h->ev_set[EV_MAP]) mark(h, EV_MAP);
    if (ids) {
        CKS(build_map_grids(h, jobs));
        lanes_join(h, 0x3);
    }
    mark(h, EV_GRID);
    if (h->dbg_alloc && g_alloc_calls) {
        fprintf(stderr, "[lvreg] allocations in this call: %d calls, %.1f MB requested, %.2f ms\n", g_alloc_calls,
                g_alloc_bytes / 1048576.0, g_alloc_ms);
        g_alloc_ms = 0.0; g_alloc_bytes = 0; g_alloc_calls = 0;
    }
    guard.commit();                                            // scan and map are consistent from here on
    if (h->debug_phases) h->dbg_host_us[5] = host_now_us();
    int s = scan2map_impl(h, pose, res);                      // scan2MapOptimization MO:322
    if (h->debug_phases) h->dbg_host_us[6] = host_now_us();
    if (s == LVREG_OK || s == LVREG_ERR_NOT_ENOUGH_FEATURES) {
        CK(cudaStreamSynchronize(h->st));
        // the four filters overlap, so they are reported together: map_build_ms covers H2D + pack +
        // local-map VoxelGrid + scan down-sampling; downsample_ms stays 0 in this fused call
        h->last.map_build_ms = span(h, EV_BEGIN, EV_MAP);
        h->last.grid_build_ms = span(h, EV_MAP, EV_GRID);
        h->last.register_ms = span(h, EV_GRID, EV_REG);
        finish_timings(h);
        if (h->debug_phases) {
            for (int l = 0; l < kLanes; ++l) {
                if (!(h->dbg_ev_used & (1u << l))) continue;
                float t[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&t[k], h->ev[EV_BEGIN], h->dbg_ev[l][k]);
                fprintf(stderr, "[lvreg] lane %d: keys start %.3f, sort start %.3f, sort done %.3f, heads done %.3f, centroid done %.3f ms "
                        "(map stage ends %.3f, grids %.3f, registration %.3f)\n", l, t[0], t[1], t[2], t[3], t[4],
                        h->last.map_build_ms, h->last.map_build_ms + h->last.grid_build_ms,
                        h->last.map_build_ms + h->last.grid_build_ms + h->last.register_ms);
            }
            h->dbg_ev_used = 0;
            float e[kLanes];
            for (int l = 0; l < kLanes; ++l) cudaEventElapsedTime(&e[l], h->ev[EV_BEGIN], h->dbg_ev[l][5]);
            const double* t = h->dbg_host_us;
            fprintf(stderr, "[lvreg] host clock (ms after the call began): VoxelGrid batch entered %.3f, all chains enqueued %.3f, "
                    "synchronised %.3f, batch left %.3f, grids enqueued %.3f, registration enqueued %.3f\n", (t[1] - t[0]) * 1e-3,
                    (t[2] - t[0]) * 1e-3, (t[3] - t[0]) * 1e-3, (t[4] - t[0]) * 1e-3, (t[5] - t[0]) * 1e-3, (t[6] - t[0]) * 1e-3);
            fprintf(stderr, "[lvreg] lanes idle at: map corner %.3f, map surf %.3f, scan corner %.3f, scan surf %.3f ms\n", e[0], e[1], e[2], e[3]);
        }
    }
    end_call(h);
    return s;
}

int lvreg_transform_update(const lvreg_handle* h, float pose[6], int imu_available, float imu_roll,
                           float imu_pitch) {
    if (!h || !pose) return LVREG_ERR_INVALID;
    transform_update_host(h->prm, pose, imu_available, imu_roll, imu_pitch);
    return LVREG_OK;
}

int lvreg_set_imu_prior(lvreg_handle* h, int imu_available, float imu_roll, float imu_pitch) {
    if (!h) return LVREG_ERR_INVALID;
    h->imu_available = imu_available ? 1 : 0;
    h->imu_roll = imu_roll;
    h->imu_pitch = imu_pitch;
    return LVREG_OK;
}

int lvreg_get_degenerate(const lvreg_handle* hc, int* is_degenerate) {
    lvreg_handle* h = const_cast<lvreg_handle*>(hc);
    if (!h || !is_degenerate) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->pinned, h->lmstate.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    *is_degenerate = *(int*)h->pinned;
    return LVREG_OK;
}

int lvreg_reset_lm_state(lvreg_handle* h) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->lmstate.p, 0, sizeof(LmState), h->st));
    CK(cudaStreamSynchronize(h->st));
    return LVREG_OK;
}

// ---- stage-level ---------------------------------------------------------------------------------
void lvreg_pose_to_affine(const float pose[6], float T[12]) { pose_to_affine_host(pose, T); }

int lvreg_transform_cloud(lvreg_handle* h, const lvreg_cloud* in, const float pose[6], lvreg_cloud_out* out) {
    if (!h || !in || !pose || !out) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    Lane& L = h->lane[LANE_SCAN_CORNER];
    CKS(upload_cloud(h, in, L.raw, L.stage, h->st));
    const uint32_t n = (uint32_t)in->n;
    CK(h->vgout.reserve((size_t)(n ? n : 1) * 16));
    Affine T;
    pose_to_affine_host(pose, T.m);
    if (n) {
        transform_kernel<<<nblk(n, 256), 256, 0, h->st>>>(L.raw.as<float4>(), n, T, h->vgout.as<float4>());
        launched(h);
    }
    int s = download_cloud(h, h->vgout.as<float4>(), n, out);
    end_call(h);
    return s;
}

int lvreg_voxelgrid(lvreg_handle* h, const lvreg_cloud* in, float leaf, lvreg_cloud_out* out, size_t* n_out,
                    uint32_t* voxel_keys_out, int* passthrough) {
    if (!h || !in || !out) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    Lane& L = h->lane[LANE_SCAN_CORNER];
    lanes_fork(h, 0x4);
    CKS(upload_cloud(h, in, L.raw, L.stage, L.st));
    uint32_t m = 0;
    VgJob J;
    J.lane = LANE_SCAN_CORNER;
    J.pts = L.raw.as<float4>();
    J.n = (uint32_t)in->n;
    J.leaf = leaf;
    J.out = &h->vgout;
    J.n_out = &m;
    J.want_out_keys = voxel_keys_out != nullptr;
    CKS(voxelgrid_batch(h, &J, 1));
    lanes_join(h, 0x4);
    mark(h, EV_DS);
    if (passthrough) *passthrough = J.passthrough;
    if (n_out) *n_out = m;
    if (voxel_keys_out && m && !J.passthrough) {
        CK(cudaMemcpyAsync(voxel_keys_out, L.vox_keys.p, (size_t)m * 4, cudaMemcpyDeviceToHost, h->st));
    } else if (voxel_keys_out && m) {
        memset(voxel_keys_out, 0, (size_t)m * 4);
    }
    int s = download_cloud(h, h->vgout.as<float4>(), m, out);
    CK(cudaStreamSynchronize(h->st));
    h->last.downsample_ms = span(h, EV_BEGIN, EV_DS);
    finish_timings(h);
    end_call(h);
    return s;
}

int lvreg_voxel_keys(lvreg_handle* h, const lvreg_cloud* in, float leaf, uint32_t* keys_out) {
    if (!h || !in || (!keys_out && in->n)) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    Lane& L = h->lane[LANE_SCAN_CORNER];
    const uint32_t n = (uint32_t)in->n;
    if (n == 0) return LVREG_OK;
    CK(h->idxbuf.reserve((size_t)n * 4));
    lanes_fork(h, 0x4);
    CKS(upload_cloud(h, in, L.raw, L.stage, L.st));
    uint32_t m = 0;
    VgJob J;
    J.lane = LANE_SCAN_CORNER;
    J.pts = L.raw.as<float4>();
    J.n = n;
    J.leaf = leaf;
    J.out = &h->vgout;
    J.n_out = &m;
    J.d_point_keys = h->idxbuf.as<uint32_t>();
    CKS(voxelgrid_batch(h, &J, 1));
    lanes_join(h, 0x4);
    CK(cudaMemcpyAsync(keys_out, h->idxbuf.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    end_call(h);
    return LVREG_OK;
}

static int knn5_launch(lvreg_handle* h, int which, const float4* q, uint32_t nq, int variant, int32_t* d_idx,
                       float* d_d2) {
    const MapSide& ms = h->map[which];
    if (variant == LVREG_KNN_BRUTE) {
        const uint32_t qblocks = nblk(nq, kBruteQpb);
        const uint32_t mblocks = nblk(nq, 256);
        uint32_t splits = (uint32_t)(4 * h->num_sms) / (qblocks ? qblocks : 1);
        if (splits < 1) splits = 1;
        uint32_t max_splits = nblk(ms.m ? ms.m : 1, kBruteTile);
        if (splits > max_splits) splits = max_splits;
        uint32_t chunk = nblk(ms.m ? ms.m : 1, splits);
        chunk = nblk(chunk, kBruteTile) * kBruteTile;
        splits = nblk(ms.m ? ms.m : 1, chunk);
        CK(h->brute_partial.reserve((size_t)splits * nq * 5 * sizeof(u64)));
        knn5_brute_kernel<<<dim3(qblocks, splits), 256, 0, h->st>>>(ms.ds.as<float4>(), ms.m, q, nq, chunk,
                                                                   h->brute_partial.as<u64>());
        knn5_brute_merge_kernel<<<mblocks, 256, 0, h->st>>>(h->brute_partial.as<u64>(), nq, splits, d_idx, d_d2);
        launched(h, 2);
    } else if (variant == LVREG_KNN_GRID_STAGED) {
        // the search of the registration kernel (shared-memory tiles filled by bulk copies), materialised
        const GridView g = grid_view(ms);
        uint32_t blocks = nblk(nq, 32);
        if (blocks > (uint32_t)h->num_sms * 2) blocks = (uint32_t)h->num_sms * 2;
        uint32_t* stats = nullptr;
        if (h->debug_tiles) {
            CK(cudaMemsetAsync(h->stagestats.p, 0, 8 * sizeof(uint32_t), h->st));
            stats = h->stagestats.as<uint32_t>();
        }
        knn5_staged_kernel<<<blocks, kRegThreads, register_staged_smem_bytes(), h->st>>>(g, q, nq, h->prm.knn_gate_sq, d_idx,
                                                                                        d_d2, stats);
        launched(h);
    } else {
        const GridView g = grid_view(ms);
        const int exact = variant == LVREG_KNN_GRID_EXACT ? 1 : 0;
        const float gate = h->prm.knn_gate_sq;
        const uint32_t cap = (uint32_t)h->num_sms * 8;
        switch (h->lpq) {
            case 4: knn5_grid_kernel<4><<<min(nblk(nq, 64), cap), 256, 0, h->st>>>(g, q, nq, exact, gate, d_idx, d_d2); break;
            case 16: knn5_grid_kernel<16><<<min(nblk(nq, 16), cap), 256, 0, h->st>>>(g, q, nq, exact, gate, d_idx, d_d2); break;
            case 32: knn5_grid_kernel<32><<<min(nblk(nq, 8), cap), 256, 0, h->st>>>(g, q, nq, exact, gate, d_idx, d_d2); break;
            default: knn5_grid_kernel<8><<<min(nblk(nq, 32), cap), 256, 0, h->st>>>(g, q, nq, exact, gate, d_idx, d_d2); break;
        }
        launched(h);
    }
    CK(cudaGetLastError());
    return LVREG_OK;
}

int lvreg_knn5(lvreg_handle* h, int which, const lvreg_cloud* queries, int variant, int32_t* idx_out, float* d2_out) {
    if (!h || which < 0 || which > 1 || !queries || variant < 0 || variant > 3) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->map[which].valid) return fail(h, LVREG_ERR_NO_MAP, "no local map");
    begin_call(h);
    CKS(upload_cloud(h, queries, h->qbuf, h->lane[LANE_SCAN_CORNER].stage, h->st));
    const uint32_t nq = (uint32_t)queries->n;
    if (nq == 0) return LVREG_OK;
    CK(h->idxbuf.reserve((size_t)nq * 5 * 4));
    CK(h->d2buf.reserve((size_t)nq * 5 * 4));
    mark(h, EV_BEGIN);
    CKS(knn5_launch(h, which, h->qbuf.as<float4>(), nq, variant, h->idxbuf.as<int32_t>(), h->d2buf.as<float>()));
    mark(h, EV_REG);
    if (idx_out) CK(cudaMemcpyAsync(idx_out, h->idxbuf.p, (size_t)nq * 5 * 4, cudaMemcpyDeviceToHost, h->st));
    if (d2_out) CK(cudaMemcpyAsync(d2_out, h->d2buf.p, (size_t)nq * 5 * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->last.register_ms = span(h, EV_BEGIN, EV_REG);
    finish_timings(h);
    end_call(h);
    return LVREG_OK;
}

int lvreg_bench_knn5(lvreg_handle* h, int which, const lvreg_cloud* queries, int variant, int repeats, float* ms) {
    if (!h || which < 0 || which > 1 || !queries || repeats < 1 || variant < 0 || variant > 3) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->map[which].valid) return fail(h, LVREG_ERR_NO_MAP, "no local map");
    begin_call(h);
    CKS(upload_cloud(h, queries, h->qbuf, h->lane[LANE_SCAN_CORNER].stage, h->st));
    const uint32_t nq = (uint32_t)queries->n;
    if (nq == 0) return fail(h, LVREG_ERR_INVALID, "no queries");
    CK(h->idxbuf.reserve((size_t)nq * 5 * 4));
    CK(h->d2buf.reserve((size_t)nq * 5 * 4));
    for (int i = 0; i < 3; ++i)
        CKS(knn5_launch(h, which, h->qbuf.as<float4>(), nq, variant, h->idxbuf.as<int32_t>(), h->d2buf.as<float>()));
    mark(h, EV_BEGIN);
    for (int i = 0; i < repeats; ++i)
        CKS(knn5_launch(h, which, h->qbuf.as<float4>(), nq, variant, h->idxbuf.as<int32_t>(), h->d2buf.as<float>()));
    mark(h, EV_REG);
    CK(cudaStreamSynchronize(h->st));
    if (ms) *ms = span(h, EV_BEGIN, EV_REG) / (float)repeats;
    end_call(h);
    return LVREG_OK;
}

// C4 "with fused residual": search + fit + residual of surfOptimization / cornerOptimization in one kernel
// (residual_kernel), device-resident queries, CUDA-event timing; nothing is copied back
int lvreg_bench_residuals(lvreg_handle* h, int which, const lvreg_cloud* queries, const float pose[6], int repeats, float* ms) {
    if (!h || which < 0 || which > 1 || !queries || !pose || repeats < 1) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->map[which].valid) return fail(h, LVREG_ERR_NO_MAP, "no local map");
    begin_call(h);
    CKS(upload_cloud(h, queries, h->qbuf, h->lane[LANE_SCAN_CORNER].stage, h->st));
    const uint32_t n = (uint32_t)queries->n;
    if (n == 0) return fail(h, LVREG_ERR_INVALID, "no queries");
    CK(h->coeffbuf.reserve((size_t)n * 16));
    CK(h->flagbuf.reserve((size_t)n));
    CK(h->idxbuf.reserve((size_t)n * 5 * 4));
    Affine T;
    pose_to_affine_host(pose, T.m);
    const GridView g = grid_view(h->map[which]);
    const RegParams P = reg_params(h);
    const uint32_t blocks = min(nblk(nblk(n, 32), kRegWarps), (uint32_t)h->num_sms * 8);
    for (int i = 0; i < repeats + 3; ++i) {
        if (i == 3) mark(h, EV_BEGIN);
        residual_kernel<8><<<blocks, kRegThreads, 0, h->st>>>(g, h->map[which].ds.as<float4>(), h->qbuf.as<float4>(), n, which, T, P,
                                                              h->coeffbuf.as<float4>(), h->flagbuf.as<uint8_t>(), h->idxbuf.as<int32_t>());
        launched(h);
    }
    mark(h, EV_REG);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->st));
    if (ms) *ms = span(h, EV_BEGIN, EV_REG) / (float)repeats;
    end_call(h);
    return LVREG_OK;
}

static int residuals_api(lvreg_handle* h, int cls, const lvreg_cloud* pts, const float pose[6], float* coeff_out,
                         uint8_t* flag_out, int32_t* knn_idx_out) {
    if (!h || !pts || !pose) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->map[cls].valid) return fail(h, LVREG_ERR_NO_MAP, "no local map");
    begin_call(h);
    CKS(upload_cloud(h, pts, h->qbuf, h->lane[LANE_SCAN_CORNER].stage, h->st));
    const uint32_t n = (uint32_t)pts->n;
    if (n == 0) return LVREG_OK;
    CK(h->coeffbuf.reserve((size_t)n * 16));
    CK(h->flagbuf.reserve((size_t)n));
    CK(h->idxbuf.reserve((size_t)n * 5 * 4));
    Affine T;
    pose_to_affine_host(pose, T.m);
    const GridView g = grid_view(h->map[cls]);
    const RegParams P = reg_params(h);
    const uint32_t blocks = min(nblk(nblk(n, 32), kRegWarps), (uint32_t)h->num_sms * 8);
    const float4* map = h->map[cls].ds.as<float4>();
    const float4* q = h->qbuf.as<float4>();
    float4* co = h->coeffbuf.as<float4>();
    uint8_t* fl = h->flagbuf.as<uint8_t>();
    int32_t* nn = h->idxbuf.as<int32_t>();
    switch (h->lpq) {
        case 4: residual_kernel<4><<<blocks, kRegThreads, 0, h->st>>>(g, map, q, n, cls, T, P, co, fl, nn); break;
        case 16: residual_kernel<16><<<blocks, kRegThreads, 0, h->st>>>(g, map, q, n, cls, T, P, co, fl, nn); break;
        case 32: residual_kernel<32><<<blocks, kRegThreads, 0, h->st>>>(g, map, q, n, cls, T, P, co, fl, nn); break;
        default: residual_kernel<8><<<blocks, kRegThreads, 0, h->st>>>(g, map, q, n, cls, T, P, co, fl, nn); break;
    }
    launched(h);
    CK(cudaGetLastError());
    if (coeff_out) CK(cudaMemcpyAsync(coeff_out, co, (size_t)n * 16, cudaMemcpyDeviceToHost, h->st));
    if (flag_out) CK(cudaMemcpyAsync(flag_out, fl, (size_t)n, cudaMemcpyDeviceToHost, h->st));
    if (knn_idx_out) CK(cudaMemcpyAsync(knn_idx_out, nn, (size_t)n * 5 * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    end_call(h);
    return LVREG_OK;
}

int lvreg_corner_residuals(lvreg_handle* h, const lvreg_cloud* pts, const float pose[6], float* coeff_out,
                           uint8_t* flag_out, int32_t* knn_idx_out) {
    return residuals_api(h, 0, pts, pose, coeff_out, flag_out, knn_idx_out);
}
int lvreg_surf_residuals(lvreg_handle* h, const lvreg_cloud* pts, const float pose[6], float* coeff_out,
                         uint8_t* flag_out, int32_t* knn_idx_out) {
    return residuals_api(h, 1, pts, pose, coeff_out, flag_out, knn_idx_out);
}

int lvreg_lm_step(lvreg_handle* h, const float* ori, const float* coeff, size_t n_sel, int iter, float pose[6],
                  float AtA_out[36], float Atb_out[6], float x_out[6], int* converged) {
    if (!h || !pose || (n_sel && (!ori || !coeff))) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    if (converged) *converged = 0;
    if ((int)n_sel < h->prm.min_matches) return LVREG_OK;       // MO:1209-1212: no update
    const uint32_t n = (uint32_t)n_sel;
    CK(h->qbuf.reserve((size_t)n * 16));
    CK(h->coeffbuf.reserve((size_t)n * 16));
    CK(cudaMemcpyAsync(h->qbuf.p, ori, (size_t)n * 16, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->coeffbuf.p, coeff, (size_t)n * 16, cudaMemcpyHostToDevice, h->st));
    float* hp = (float*)h->pinned;
    memcpy(hp, pose, 6 * sizeof(float));
    float* dscr = h->posebuf.as<float>();        // [0..5] pose, [8..43] AtA, [44..49] Atb, [50..55] x, [56] conv
    CK(cudaMemcpyAsync(dscr, hp, 6 * sizeof(float), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemsetAsync(dscr + 8, 0, 49 * sizeof(float), h->st));
    Trig g;
    // the reference takes float sin/cos of the pose on the host (MO:1202-1207)
    g.srx = sinf(pose[1]); g.crx = cosf(pose[1]);
    g.sry = sinf(pose[2]); g.cry = cosf(pose[2]);
    g.srz = sinf(pose[0]); g.crz = cosf(pose[0]);
    lm_step_kernel<<<1, 256, 0, h->st>>>(h->qbuf.as<float4>(), h->coeffbuf.as<float4>(), n, iter, g, reg_params(h), dscr,
                                         h->lmstate.as<LmState>(), dscr + 8, dscr + 44, dscr + 50, (int*)(dscr + 56));
    launched(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hp, dscr, 57 * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    memcpy(pose, hp, 6 * sizeof(float));
    if (AtA_out) memcpy(AtA_out, hp + 8, 36 * sizeof(float));
    if (Atb_out) memcpy(Atb_out, hp + 44, 6 * sizeof(float));
    if (x_out) memcpy(x_out, hp + 50, 6 * sizeof(float));
    if (converged) *converged = *(int*)(hp + 56);
    end_call(h);
    return LVREG_OK;
}

// ---- "next" row: FeatureExtraction --------------------------------------------------------------
int lvreg_extract_features(lvreg_handle* h, const lvreg_cloud* deskewed, const lvreg_scan_info* info, float edge_th,
                           float surf_th, float surf_leaf, lvreg_cloud_out* corner, size_t* n_corner,
                           lvreg_cloud_out* surf, size_t* n_surf, int32_t* label_out) {
    if (!h || !deskewed || !info || !info->start_ring_index || !info->end_ring_index) return LVREG_ERR_INVALID;
    if (info->n_scan < 1 || info->n_scan > 256) return fail(h, LVREG_ERR_INVALID, "n_scan must be in [1, 256]");
    if (deskewed->n && (!info->point_col_ind || !info->point_range)) return LVREG_ERR_INVALID;
    if (!(surf_leaf > 0.f)) return fail(h, LVREG_ERR_INVALID, "leaf size must be positive");
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    Lane& L = h->lane[LANE_SCAN_SURF];
    const uint32_t n = (uint32_t)deskewed->n;
    const int ns = info->n_scan;
    h->n_feat[0] = h->n_feat[1] = 0;
    if (n_corner) *n_corner = 0;
    if (n_surf) *n_surf = 0;
    CKS(upload_cloud(h, deskewed, h->feat_pts, L.stage, h->st));
    CK(h->feat_corner.reserve((size_t)ns * kFeCornerStride * 16));
    if (n == 0) return LVREG_OK;
    // shared-memory budget of the per-ring kernel from the ring / sector lengths
    // A ring is processed when end - start >= 1 (cloudExtraction, imageProjection.cpp:617-640, writes
    // start = first + 4, end = last - 5: a ring of fewer than 11 points has end < start and is skipped).  A processed
    // ring indexes [start - 5, end + 4] clipped to the cloud, its sectors [start, end]: both ends must lie inside the
    // cloud, and processed rings must be disjoint and ascending (each block owns its ring's labels and flags).
    int cap = 16, max_sector = 1;
    long long prev_end = -1;
    for (int r = 0; r < ns; ++r) {
        const long long a = info->start_ring_index[r], b = info->end_ring_index[r];
        if (b - a < 1) continue;
        if (a < 0 || b > (long long)n - 1) return fail(h, LVREG_ERR_INVALID, "ring index out of range");
        if (a <= prev_end) return fail(h, LVREG_ERR_INVALID, "rings overlap or are not ascending");
        prev_end = b;
        long long lo = a - 5, hi = b + 4;
        if (lo < 0) lo = 0;
        if (hi > (long long)n - 1) hi = (long long)n - 1;
        if (hi - lo + 1 > cap) cap = (int)(hi - lo + 1);
        if ((b - a) / 6 + 2 > max_sector) max_sector = (int)((b - a) / 6 + 2);
    }
    int sort_cap = 2;
    while (sort_cap < max_sector) sort_cap <<= 1;
    const size_t smem = fe_ring_smem_bytes(cap, sort_cap);
    if (smem > 220 * 1024) return fail(h, LVREG_ERR_INVALID, "a ring is too long for the feature-extraction kernel");
    CK(cudaFuncSetAttribute(fe_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    CK(h->feat_range.reserve((size_t)n * 4));
    CK(h->feat_col.reserve((size_t)n * 4));
    CK(h->feat_rings.reserve((size_t)ns * 8));
    CK(h->feat_curv.reserve((size_t)n * 4));
    CK(h->feat_picked.reserve(n));
    CK(h->feat_label.reserve(n));
    CK(h->feat_flag.reserve((size_t)(n + 8) * 4));
    CK(h->feat_ringof.reserve(n));
    CK(h->feat_cidx.reserve((size_t)ns * kFeCornerStride * 4));
    CK(h->feat_ccnt.reserve((size_t)ns * 4 + 16));
    CK(h->feat_pos.reserve((size_t)(n + 8) * 4));
    CK(h->feat_cand.reserve((size_t)(n + 8) * 4));
    CK(h->feat_spec.reserve((size_t)ns * sizeof(RingSpec)));
    CKS(ensure_sort_buffers(h, L, n));
    // host or device pointers (lvreg_get_projection hands out device ones): unified addressing decides
    CK(cudaMemcpyAsync(h->feat_range.p, info->point_range, (size_t)n * 4, cudaMemcpyDefault, h->st));
    CK(cudaMemcpyAsync(h->feat_col.p, info->point_col_ind, (size_t)n * 4, cudaMemcpyDefault, h->st));
    int32_t* d_start = h->feat_rings.as<int32_t>();
    int32_t* d_end = d_start + ns;
    CK(cudaMemcpyAsync(d_start, info->start_ring_index, (size_t)ns * 4, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(d_end, info->end_ring_index, (size_t)ns * 4, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemsetAsync(h->feat_picked.p, 0, n, h->st));
    CK(cudaMemsetAsync(h->feat_label.p, 0, n, h->st));
    CK(cudaMemsetAsync(h->feat_flag.p, 0, (size_t)(n + 8) * 4, h->st));
    CK(cudaMemsetAsync(h->feat_ringof.p, 0, n, h->st));
    int32_t* d_err = h->feat_ccnt.as<int32_t>() + ns;
    CK(cudaMemsetAsync(d_err, 0, 4, h->st));
    uint32_t* d_small = L.small.as<uint32_t>();

    fe_smooth_kernel<<<nblk(n, 256), 256, 0, h->st>>>(h->feat_range.as<float>(), h->feat_col.as<int32_t>(), (int)n,
                                                      h->feat_curv.as<float>(), h->feat_picked.as<uint8_t>());
    fe_ring_kernel<<<ns, 256, smem, h->st>>>(h->feat_curv.as<float>(), h->feat_picked.as<uint8_t>(), h->feat_col.as<int32_t>(),
                                             (int)n, d_start, d_end, edge_th, surf_th, cap, sort_cap,
                                             h->feat_label.as<int8_t>(), h->feat_flag.as<uint32_t>(),
                                             h->feat_ringof.as<uint8_t>(), h->feat_cidx.as<int32_t>(),
                                             h->feat_ccnt.as<int32_t>(), d_err);
    fe_corner_gather_kernel<<<ns, 256, 0, h->st>>>(h->feat_pts.as<float4>(), h->feat_cidx.as<int32_t>(),
                                                  h->feat_ccnt.as<int32_t>(), ns, h->feat_corner.as<float4>(), d_small + SM_TOTAL);
    launched(h, 3);
    exclusive_scan(FlagIn{h->feat_flag.as<uint32_t>()}, CompactOut{h->feat_pos.as<uint32_t>(), h->feat_cand.as<uint32_t>()}, n,
                   L.scan_temp.as<uint32_t>(), d_small + SM_NVOX, h->st, &h->call_launches);
    uint32_t* host = L.pinned;
    CK(cudaMemcpyAsync(host, d_small + SM_TOTAL, 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(host + 1, d_small + SM_NVOX, 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(host + 2, d_err, 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (host[2]) return fail(h, LVREG_ERR_INVALID, "ring layout exceeds the feature-extraction kernel's limits");
    const uint32_t nc = host[0], n_cand = host[1];
    uint32_t nsurf = 0;
    if (n_cand) {
        CK(h->feat_idx.reserve((size_t)n_cand * 4));
        CK(h->feat_pidx.reserve((size_t)n_cand * 4));
        CK(L.vox_start.reserve((size_t)n_cand * 4));
        fe_ring_bbox_kernel<<<ns, 256, 0, h->st>>>(h->feat_pts.as<float4>(), h->feat_flag.as<uint32_t>(), h->feat_pos.as<uint32_t>(),
                                                   d_start, d_end, (int)n, surf_leaf, h->feat_spec.as<RingSpec>());
        fe_surf_keys_kernel<<<nblk(n_cand, 256), 256, 0, h->st>>>(h->feat_pts.as<float4>(), h->feat_cand.as<uint32_t>(), n_cand,
                                                                  h->feat_ringof.as<uint8_t>(), h->feat_spec.as<RingSpec>(),
                                                                  L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(),
                                                                  h->feat_idx.as<uint32_t>());
        launched(h, 2);
        int cur = radix_sort_pairs(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), L.keys[1].as<uint32_t>(),
                                   L.vals[1].as<uint32_t>(), n_cand, 32, L.sort_scratch.as<uint32_t>(), h->st, &h->call_launches);
        // second, stable sort by ring: (ring, idx, input order)
        uint32_t* k_a = L.keys[cur].as<uint32_t>();
        uint32_t* v_a = L.vals[cur].as<uint32_t>();
        uint32_t* k_b = L.keys[cur ^ 1].as<uint32_t>();
        uint32_t* v_b = L.vals[cur ^ 1].as<uint32_t>();
        fe_ring_keys_kernel<<<nblk(n_cand, 256), 256, 0, h->st>>>(v_a, n_cand, h->feat_cand.as<uint32_t>(),
                                                                  h->feat_ringof.as<uint8_t>(), k_a);
        launched(h);
        int cur2 = radix_sort_pairs(k_a, v_a, k_b, v_b, n_cand, bits_for((uint64_t)(ns > 1 ? ns - 1 : 1)),
                                    L.sort_scratch.as<uint32_t>(), h->st, &h->call_launches);
        const uint32_t* rk = cur2 ? k_b : k_a;
        const uint32_t* rv = cur2 ? v_b : v_a;
        exclusive_scan(RingVoxelHeadIn{rk, rv, h->feat_idx.as<uint32_t>()}, VoxelStartOut{L.vox_start.as<uint32_t>()}, n_cand,
                       L.scan_temp.as<uint32_t>(), d_small + SM_NVOX, h->st, &h->call_launches);
        CK(cudaMemcpyAsync(host, d_small + SM_NVOX, 4, cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        nsurf = host[0];
        CK(h->feat_surf.reserve((size_t)(nsurf ? nsurf : 1) * 16));
        fe_point_index_kernel<<<nblk(n_cand, 256), 256, 0, h->st>>>(rv, n_cand, h->feat_cand.as<uint32_t>(), h->feat_pidx.as<uint32_t>());
        centroid_kernel<<<nblk(nsurf, 128), 128, 0, h->st>>>(h->feat_pts.as<float4>(), rk, h->feat_pidx.as<uint32_t>(),
                                                             L.vox_start.as<uint32_t>(), d_small + SM_NVOX, n_cand,
                                                             h->feat_surf.as<float4>(), nullptr);
        launched(h, 2);
    } else {
        CK(h->feat_surf.reserve(16));
    }
    CK(cudaGetLastError());
    mark(h, EV_DS);
    h->n_feat[0] = nc;
    h->n_feat[1] = nsurf;
    if (n_corner) *n_corner = nc;
    if (n_surf) *n_surf = nsurf;
    if (label_out) {
        std::vector<int8_t> tmp(n);
        CK(cudaMemcpyAsync(tmp.data(), h->feat_label.p, n, cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        for (uint32_t i = 0; i < n; ++i) label_out[i] = tmp[i];
    }
    if (corner) CKS(download_cloud(h, h->feat_corner.as<float4>(), nc, corner));
    if (surf) CKS(download_cloud(h, h->feat_surf.as<float4>(), nsurf, surf));
    CK(cudaStreamSynchronize(h->st));
    h->last.downsample_ms = span(h, EV_BEGIN, EV_DS);
    finish_timings(h);
    end_call(h);
    return LVREG_OK;
}

int lvreg_get_feature_clouds(const lvreg_handle* h, lvreg_cloud* corner, lvreg_cloud* surf) {
    if (!h || !corner || !surf) return LVREG_ERR_INVALID;
    lvreg_cloud* c[2] = {corner, surf};
    const DevBuf* b[2] = {&h->feat_corner, &h->feat_surf};
    for (int s = 0; s < 2; ++s) {
        c[s]->data = b[s]->p;
        c[s]->n = h->n_feat[s];
        c[s]->stride = 16;
        c[s]->intensity_offset = 12;
        c[s]->on_device = 1;
        c[s]->reserved = 0;
    }
    return LVREG_OK;
}

__global__ void __launch_bounds__(256) bench_fill_kernel(uint32_t* keys, uint32_t* vals, uint32_t n, uint32_t mask) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t x = i * 2654435761u + 0x9e3779b9u;      // cheap integer hash
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    keys[i] = x & mask;
    vals[i] = i;
}

int lvreg_bench_sort(lvreg_handle* h, size_t n_, int key_bits, int repeats, float* ms_per_sort, int* passes) {
    if (!h || n_ == 0 || n_ > 0x7fffffffull || key_bits < 1 || key_bits > 32 || repeats < 1) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    const uint32_t n = (uint32_t)n_;
    Lane& L = h->lane[LANE_MAP_SURF];
    CKS(ensure_sort_buffers(h, L, n));
    const uint32_t mask = key_bits >= 32 ? 0xffffffffu : ((1u << key_bits) - 1u);
    float total = 0.f;
    for (int r = 0; r < repeats + 2; ++r) {           // 2 warm-up rounds
        bench_fill_kernel<<<nblk(n, 256), 256, 0, h->st>>>(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), n, mask);
        mark(h, EV_BEGIN);
        radix_sort_pairs(L.keys[0].as<uint32_t>(), L.vals[0].as<uint32_t>(), L.keys[1].as<uint32_t>(),
                         L.vals[1].as<uint32_t>(), n, key_bits, L.sort_scratch.as<uint32_t>(), h->st, &h->call_launches);
        mark(h, EV_REG);
        CK(cudaStreamSynchronize(h->st));
        if (r >= 2) total += span(h, EV_BEGIN, EV_REG);
    }
    if (ms_per_sort) *ms_per_sort = total / (float)repeats;
    if (passes) *passes = (key_bits + 7) / 8;
    end_call(h);
    return LVREG_OK;
}

#ifdef LVREG_SORT_PROF
extern "C" int lvreg_debug_sort_prof(unsigned long long* out, int reset) {
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_sort_prof, z, sizeof(z)); return 0; }
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_sort_prof, 16 * sizeof(unsigned long long));
    return 0;
}
#endif


// ---- loop closure ---------------------------------------------------------------------
namespace {

constexpr float kIcpCell = 1.0f;          // fine search-grid cell of the ICP target, metres
constexpr float kIcpCoarseCell = 6.0f;    // coarse grid for the queries the fine grid cannot resolve in 27 cells

// both search grids of the ICP target, on lane 1, from the VoxelGrid job's bounding box when there is one
int build_icp_grids(lvreg_handle* h, const float* mn, const float* mx) {
    Lane& L = h->lane[1];
    MapSide& t = h->icp_cloud[1];
    float bmn[3], bmx[3];
    if (!mn && t.m) {
        // one bounding-box pass serves both grids
        uint32_t* mm = L.small.as<uint32_t>() + SM_MM;
        CK(cudaMemsetAsync(mm, 0xff, 3 * sizeof(uint32_t), L.st));
        CK(cudaMemsetAsync(mm + 3, 0, 3 * sizeof(uint32_t), L.st));
        minmax_kernel<<<min(nblk(t.m, 256), (uint32_t)h->num_sms * 8), 256, 0, L.st>>>(t.ds.as<float4>(), t.m, mm);
        launched(h);
        CK(cudaMemcpyAsync(L.pinned, mm, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
        CK(cudaStreamSynchronize(L.st));
        for (int a = 0; a < 3; ++a) {
            bmn[a] = ordered_to_float(L.pinned[a]);
            bmx[a] = ordered_to_float(L.pinned[3 + a]);
        }
        mn = bmn;
        mx = bmx;
    }
    CKS(build_grid(h, L, t, mn, mx, kIcpCell));
    h->icp_coarse.m = t.m;
    CKS(build_grid(h, L, h->icp_coarse, mn, mx, kIcpCoarseCell, t.ds.as<float4>()));
    return LVREG_OK;
}
constexpr size_t kPinnedIcpState = 48 * 1024;

// segment list + VoxelGrid job of a keyframe submap on lane `slot`: the clouds selected by `which`
// (bit 0 corner, bit 1 surf; corner first) of every listed keyframe under its stored pose, in list order
int prepare_submap_job(lvreg_handle* h, const int32_t* ids, size_t n_ids, int which, float leaf, int slot, VgJob& J) {
    Lane& L = h->lane[slot];
    MapSide& ms = h->icp_cloud[slot];
    L.seg_host.clear();
    uint64_t total = 0;
    const int K = (int)h->kfs.size();
    for (size_t i = 0; i < n_ids; ++i) {
        const int k = ids[i];
        if (k < 0 || k >= K) return fail(h, LVREG_ERR_INVALID, "keyframe id out of range");
        const Keyframe* kf = h->kfs[k];
        for (int s = 0; s < 2; ++s) {              // corner, then surf, of every keyframe (MO:730-731, 500-501)
            if (!(which & (1 << s)) || kf->n[s] == 0) continue;
            Segment sg;
            sg.src = kf->cloud[s].as<float4>();
            sg.begin = (uint32_t)total;
            sg.n = kf->n[s];
            sg.wkey = nullptr;
            sg.kminb[0] = sg.kminb[1] = sg.kminb[2] = 0;
            pose_to_affine_host(kf->pose, sg.T.m);
            L.seg_host.push_back(sg);
            total += kf->n[s];
        }
    }
    if (total > 0x7fffffffull) return fail(h, LVREG_ERR_INVALID, "submap too large");
    ms.n_in = total;
    ms.valid = false;
    J = VgJob();
    J.lane = slot;
    J.n = (uint32_t)total;
    J.from_segments = true;
    J.leaf = leaf;
    J.out = &ms.ds;
    J.n_out = &ms.m;
    return LVREG_OK;
}

// loopFindNearKeyframes for one slot (MO:719-741): keyframes [key - n, key + n], downSizeFilterICP (MO:249)
int prepare_loop_job(lvreg_handle* h, int key, int search_num, int slot, VgJob& J) {
    std::vector<int32_t> ids;
    const int K = (int)h->kfs.size();
    for (int i = -search_num; i <= search_num; ++i) {
        const int k = key + i;
        if (k < 0 || k >= K) continue;
        ids.push_back(k);
    }
    return prepare_submap_job(h, ids.data(), ids.size(), 3, h->prm.surf_leaf, slot, J);
}

IcpParams icp_params_dev(const lvreg_icp_params* p) {
    IcpParams P;
    P.max_d2 = (double)p->max_corr_dist * (double)p->max_corr_dist;
    P.rot_thr = 1.0 - p->transformation_epsilon;
    P.trans_thr = p->transformation_epsilon;
    P.rel_mse = p->euclidean_fitness_epsilon;
    P.abs_mse = 1e-12;
    P.max_iterations = p->max_iterations;
    return P;
}

int icp_align_impl(lvreg_handle* h, const lvreg_icp_params* prm, lvreg_icp_result* res) {
    memset(res, 0, sizeof(*res));
    for (int i = 0; i < 4; ++i) res->final_transformation[i * 5] = 1.f;
    res->fitness = 1.7976931348623157e308;
    const uint32_t ns = h->icp_cloud[0].m, nt = h->icp_cloud[1].m;
    if (!h->icp_cloud[0].valid || !h->icp_cloud[1].valid) return fail(h, LVREG_ERR_NO_MAP, "ICP source / target not set");
    if (ns == 0 || nt == 0) { res->state = ICP_NO_INPUT; return LVREG_OK; }
    if (prm->max_iterations < 1) return fail(h, LVREG_ERR_INVALID, "max_iterations < 1");
    const IcpParams P = icp_params_dev(prm);
    const uint32_t nb = nblk(ns, kIcpThreads);
    CK(h->icp_cur.reserve((size_t)ns * 16));
    CK(h->icp_partials.reserve((size_t)nb * kIcpMoments * 8));
    CK(h->icp_state.reserve(sizeof(IcpState)));
    IcpState* hs = (IcpState*)((char*)h->pinned + kPinnedIcpState);
    memset(hs, 0, sizeof(*hs));
    hs->mse_prev = 1.7976931348623157e308;
    for (int i = 0; i < 4; ++i) hs->T_final[i * 5] = hs->T_inc[i * 5] = 1.f;
    CK(cudaMemcpyAsync(h->icp_state.p, hs, sizeof(IcpState), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->icp_cur.p, h->icp_cloud[0].ds.p, (size_t)ns * 16, cudaMemcpyDeviceToDevice, h->st));
    const GridView g = grid_view(h->icp_cloud[1]), gc = grid_view(h->icp_coarse);
    IcpState* ds = h->icp_state.as<IcpState>();
    // the per-iteration decision is taken on the device; the host polls every few iterations
    const int batch = 4;
    for (int it = 0; it < prm->max_iterations;) {
        for (int b = 0; b < batch && it < prm->max_iterations; ++b, ++it) {
            icp_correspond_kernel<<<nb, kIcpThreads, 0, h->st>>>(h->icp_cur.as<float4>(), ns, g, gc, P, ds,
                                                               h->icp_partials.as<double>());
            icp_update_kernel<<<1, kIcpThreads, 0, h->st>>>(h->icp_partials.as<double>(), nb, P, ds);
            icp_transform_kernel<<<nb, kIcpThreads, 0, h->st>>>(h->icp_cur.as<float4>(), ns, ds);
            launched(h, 3);
        }
        CK(cudaMemcpyAsync(hs, ds, sizeof(IcpState), cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        if (hs->done) break;
    }
    // getFitnessScore: the original source under the final transformation, no range cap
    icp_fitness_kernel<<<nb, kIcpThreads, 0, h->st>>>(h->icp_cloud[0].ds.as<float4>(), ns, g, gc, ds,
                                                    h->icp_partials.as<double>());
    icp_fitness_reduce_kernel<<<1, kIcpThreads, 0, h->st>>>(h->icp_partials.as<double>(), nb, ds);
    launched(h, 2);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hs, ds, sizeof(IcpState), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    res->iterations = hs->iterations;
    res->state = hs->state;
    res->converged = (hs->state == ICP_ITERATIONS || hs->state == ICP_TRANSFORM || hs->state == ICP_ABS_MSE ||
                      hs->state == ICP_REL_MSE) ? 1 : 0;
    res->n_correspondences = hs->n_corr;
    res->mse = hs->mse;
    res->fitness = hs->fitness_sum / (double)ns;
    memcpy(res->final_transformation, hs->T_final, sizeof(hs->T_final));
    return LVREG_OK;
}

// 4x4 float product with the coefficient sums taken over k in order
void mat4_mul_host(const float* A, const float* B, float* C) {
    float R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = A[i * 4 + 0] * B[0 * 4 + j];
            s = s + A[i * 4 + 1] * B[1 * 4 + j];
            s = s + A[i * 4 + 2] * B[2 * 4 + j];
            s = s + A[i * 4 + 3] * B[3 * 4 + j];
            R[i * 4 + j] = s;
        }
    memcpy(C, R, sizeof(R));
}

}  // namespace

void lvreg_icp_default_params(lvreg_icp_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->max_corr_dist = 30.0f;                 // historyKeyframeSearchRadius * 2 (MO:580, utility.h:296)
    p->max_iterations = 100;                  // MO:581
    p->transformation_epsilon = 1e-6;         // MO:582
    p->euclidean_fitness_epsilon = 1e-6;      // MO:583
}

int lvreg_loop_find_near_keyframes(lvreg_handle* h, int key, int search_num, int slot, size_t* n_out) {
    if (!h || slot < 0 || slot > 1 || search_num < 0) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (h->kfs.empty()) return fail(h, LVREG_ERR_NO_KEYFRAMES, "no keyframes");
    begin_call(h);
    mark(h, EV_BEGIN);
    VgJob J;
    CKS(prepare_loop_job(h, key, search_num, slot, J));
    lanes_fork(h, 1u << slot);
    CKS(voxelgrid_batch(h, &J, 1));
    lanes_join(h, 1u << slot);
    if (!h->ev_set[EV_MAP]) mark(h, EV_MAP);
    if (slot == 1) {
        CKS(build_icp_grids(h, J.n ? J.mn : nullptr, J.n ? J.mx : nullptr));
        lanes_join(h, 1u << slot);
    }
    mark(h, EV_GRID);
    CK(cudaStreamSynchronize(h->st));
    h->icp_cloud[slot].valid = true;
    h->last.map_build_ms = span(h, EV_BEGIN, EV_MAP);
    h->last.grid_build_ms = span(h, EV_MAP, EV_GRID);
    finish_timings(h);
    end_call(h);
    if (n_out) *n_out = h->icp_cloud[slot].m;
    return LVREG_OK;
}

// global map for visualisation / saving (publishGlobalMap MO:493-508, saveMapService MO:199-231): the selected
// clouds of the listed keyframes under their stored poses, concatenated in list order, then one VoxelGrid
int lvreg_build_global_map(lvreg_handle* h, const int32_t* ids, size_t n_ids, int which, float leaf, size_t* n_out) {
    if (!h || (!ids && n_ids) || which < 1 || which > 3 || !(leaf >= 0.f)) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (h->kfs.empty()) return fail(h, LVREG_ERR_NO_KEYFRAMES, "no keyframes");
    begin_call(h);
    mark(h, EV_BEGIN);
    VgJob J;
    CKS(prepare_submap_job(h, ids, n_ids, which, leaf > 0.f ? leaf : 1.0f, 0, J));
    if (leaf == 0.f) {
        // saveMapService with resolution 0 (MO:219-225): the transformed clouds, concatenated, no VoxelGrid
        Lane& L = h->lane[0];
        MapSide& ms = h->icp_cloud[0];
        CK(ms.ds.reserve((size_t)(J.n ? J.n : 1) * 16));
        if (J.n) {
            uint32_t* mm = L.small.as<uint32_t>() + SM_MM;
            CKS(upload_segs(h, L, h->st));
            transform_concat_kernel<<<min(nblk(J.n, 2048), (uint32_t)h->num_sms * 16), 256, 0, h->st>>>(
                L.segs.as<Segment>(), (uint32_t)L.seg_host.size(), J.n, ms.ds.as<float4>(), mm);
            launched(h);
        }
        CK(cudaStreamSynchronize(h->st));
        ms.m = J.n;
        ms.valid = true;
        end_call(h);
        if (n_out) *n_out = ms.m;
        return LVREG_OK;
    }
    lanes_fork(h, 0x1);
    CKS(voxelgrid_batch(h, &J, 1));
    lanes_join(h, 0x1);
    if (!h->ev_set[EV_MAP]) mark(h, EV_MAP);
    CK(cudaStreamSynchronize(h->st));
    h->icp_cloud[0].valid = true;                 // read back with lvreg_icp_get_cloud(h, 0, ...)
    h->last.map_build_ms = span(h, EV_BEGIN, EV_MAP);
    finish_timings(h);
    end_call(h);
    if (n_out) *n_out = h->icp_cloud[0].m;
    return LVREG_OK;
}

int lvreg_icp_set_cloud(lvreg_handle* h, int slot, const lvreg_cloud* cloud) {
    if (!h || slot < 0 || slot > 1 || !cloud) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    Lane& L = h->lane[slot];
    MapSide& ms = h->icp_cloud[slot];
    ms.valid = false;
    lanes_fork(h, 1u << slot);
    CKS(upload_cloud(h, cloud, ms.ds, L.stage, L.st));
    ms.m = (uint32_t)cloud->n;
    ms.n_in = cloud->n;
    if (slot == 1) CKS(build_icp_grids(h, nullptr, nullptr));
    CK(lanes_sync(h, 1u << slot));
    ms.valid = true;
    end_call(h);
    return LVREG_OK;
}

int lvreg_icp_get_cloud(lvreg_handle* h, int slot, lvreg_cloud_out* out, size_t* n) {
    if (!h || slot < 0 || slot > 1) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->icp_cloud[slot].valid) return fail(h, LVREG_ERR_NO_MAP, "ICP cloud not set");
    if (n) *n = h->icp_cloud[slot].m;
    if (!out) return LVREG_OK;
    return download_cloud(h, h->icp_cloud[slot].ds.as<float4>(), h->icp_cloud[slot].m, out);
}

int lvreg_nn1(lvreg_handle* h, const lvreg_cloud* queries, float max_dist, int32_t* idx_out, float* d2_out) {
    if (!h || !queries || (queries->n && (!idx_out || !d2_out))) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!h->icp_cloud[1].valid) return fail(h, LVREG_ERR_NO_MAP, "ICP target not set");
    begin_call(h);
    const uint32_t n = (uint32_t)queries->n;
    if (n == 0) { end_call(h); return LVREG_OK; }
    CKS(upload_cloud(h, queries, h->qbuf, h->lane[LANE_SCAN_CORNER].stage, h->st));
    CK(h->icp_idx.reserve((size_t)n * 4));
    CK(h->icp_d2.reserve((size_t)n * 4));
    const float max_d2 = max_dist > 0.f && max_dist < 1e18f ? max_dist * max_dist : INFINITY;
    nn1_kernel<<<nblk(n, kIcpThreads), kIcpThreads, 0, h->st>>>(h->qbuf.as<float4>(), n, grid_view(h->icp_cloud[1]),
                                                             grid_view(h->icp_coarse), max_d2,
                                                             h->icp_idx.as<int32_t>(), h->icp_d2.as<float>());
    launched(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(idx_out, h->icp_idx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(d2_out, h->icp_d2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    end_call(h);
    return LVREG_OK;
}

int lvreg_icp_align(lvreg_handle* h, const lvreg_icp_params* prm, lvreg_icp_result* res) {
    if (!h || !prm || !res) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    int s = icp_align_impl(h, prm, res);
    mark(h, EV_REG);
    CK(cudaStreamSynchronize(h->st));
    h->last.register_ms = span(h, EV_BEGIN, EV_REG);
    finish_timings(h);
    end_call(h);
    return s;
}

int lvreg_correct_pose(const float* correction4x4, const float pose[6], float out[6]) {
    if (!correction4x4 || !pose || !out) return LVREG_ERR_INVALID;
    float T12[12], W[16], C[16];
    pose_to_affine_host(pose, T12);              // tWrong = pclPointToAffine3f(copy_cloudKeyPoses6D[loopKeyCur]), MO:602
    memcpy(W, T12, sizeof(T12));
    W[12] = W[13] = W[14] = 0.f; W[15] = 1.f;
    mat4_mul_host(correction4x4, W, C);          // tCorrect = correctionLidarFrame * tWrong, MO:604
    out[3] = C[3]; out[4] = C[7]; out[5] = C[11];   // pcl::getTranslationAndEulerAngles, MO:605
    out[0] = atan2f(C[9], C[10]);
    out[1] = (float)asin((double)-C[8]);
    out[2] = atan2f(C[4], C[0]);
    return LVREG_OK;
}

int lvreg_perform_loop_closure(lvreg_handle* h, int key_cur, int key_pre, int search_num, const lvreg_icp_params* prm,
                               float fitness_gate, lvreg_loop_result* out) {
    if (!h || !prm || !out || search_num < 0) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const int K = (int)h->kfs.size();
    if (K == 0) return fail(h, LVREG_ERR_NO_KEYFRAMES, "no keyframes");
    if (key_cur < 0 || key_cur >= K || key_pre < 0 || key_pre >= K) return fail(h, LVREG_ERR_INVALID, "keyframe id out of range");
    memset(out, 0, sizeof(*out));
    begin_call(h);
    mark(h, EV_BEGIN);
    // both submaps at once, one lane each (MO:569-571)
    VgJob jobs[2];
    CKS(prepare_loop_job(h, key_cur, 0, 0, jobs[0]));
    CKS(prepare_loop_job(h, key_pre, search_num, 1, jobs[1]));
    lanes_fork(h, 0x3);
    CKS(voxelgrid_batch(h, jobs, 2));
    lanes_join(h, 0x3);
    if (!h->ev_set[EV_MAP]) mark(h, EV_MAP);
    CKS(build_icp_grids(h, jobs[1].n ? jobs[1].mn : nullptr, jobs[1].n ? jobs[1].mx : nullptr));
    lanes_join(h, 0x2);
    mark(h, EV_GRID);
    h->icp_cloud[0].valid = h->icp_cloud[1].valid = true;
    out->n_source = (int32_t)h->icp_cloud[0].m;
    out->n_target = (int32_t)h->icp_cloud[1].m;
    int s = LVREG_OK;
    if (out->n_source < 300 || out->n_target < 1000) {          // MO:572
        out->status = LVREG_LOOP_SUBMAP_TOO_SMALL;
    } else {
        s = icp_align_impl(h, prm, &out->icp);
        if (s == LVREG_OK) {
            if (!out->icp.converged) out->status = LVREG_LOOP_NOT_CONVERGED;                 // MO:592
            else if (out->icp.fitness > (double)fitness_gate) out->status = LVREG_LOOP_FITNESS_TOO_HIGH;
            else {
                lvreg_correct_pose(out->icp.final_transformation, h->kfs[key_cur]->pose, out->pose_from);
                memcpy(out->pose_to, h->kfs[key_pre]->pose, 6 * sizeof(float));              // MO:611
                out->noise = (float)out->icp.fitness;                                        // MO:613
                out->status = LVREG_LOOP_OK;
            }
        }
    }
    mark(h, EV_REG);
    CK(cudaStreamSynchronize(h->st));
    h->last.map_build_ms = span(h, EV_BEGIN, EV_MAP);
    h->last.grid_build_ms = span(h, EV_MAP, EV_GRID);
    h->last.register_ms = span(h, EV_GRID, EV_REG);
    finish_timings(h);
    end_call(h);
    return s;
}

// ---- LiDAR depth for visual features -----------------------------------------------------
namespace {
Affine affine_from(const float T[12]) {
    Affine A;
    for (int i = 0; i < 12; ++i) A.m[i] = T[i];
    return A;
}
}  // namespace

int lvreg_depth_clear(lvreg_handle* h) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->st));
    for (DepthEntry* e : h->depth_queue) h->depth_free.push_back(e);
    h->depth_queue.clear();
    h->n_depth = 0;
    return LVREG_OK;
}

int lvreg_depth_add_cloud(lvreg_handle* h, const lvreg_cloud* cloud, const float T_now[12], double stamp, size_t* n_depth) {
    if (!h || !cloud || !T_now) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    // 2. 0.2 m VoxelGrid of the new cloud (FN:303-309)
    Lane& L = h->lane[LANE_SCAN_CORNER];
    lanes_fork(h, 0x4);
    CKS(upload_cloud(h, cloud, L.raw, L.stage, L.st));
    uint32_t n1 = 0;
    VgJob J;
    J.lane = LANE_SCAN_CORNER;
    J.pts = L.raw.as<float4>();
    J.n = (uint32_t)cloud->n;
    J.leaf = 0.2f;
    J.out = &h->vgout;
    J.n_out = &n1;
    CKS(voxelgrid_batch(h, &J, 1));
    lanes_join(h, 0x4);
    // 3. + 5. camera-view filter and transform into the odometry frame, order kept (FN:313-331)
    DepthEntry* e;
    if (!h->depth_free.empty()) { e = h->depth_free.back(); h->depth_free.pop_back(); }
    else e = new DepthEntry();
    e->stamp = stamp;
    e->n = 0;
    h->depth_queue.push_back(e);             // owned by the queue from here on (error paths included)
    if (n1) {
        CK(e->pts.reserve((size_t)n1 * 16));
        CK(L.scan_temp.reserve((size_t)(scan_num_tiles(n1) + 2) * 4));
        uint32_t* d_total = L.small.as<uint32_t>() + SM_TOTAL;
        exclusive_scan(ViewFlagIn{h->vgout.as<float4>()}, ViewCompactOut{h->vgout.as<float4>(), affine_from(T_now), e->pts.as<float4>()},
                       n1, L.scan_temp.as<uint32_t>(), d_total, h->st, &h->call_launches);
        CK(cudaMemcpyAsync(L.pinned, d_total, 4, cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        e->n = L.pinned[0];
    }
    // 6. + 7. queue, drop clouds older than 5 s (FN:341-357)
    while (!h->depth_queue.empty() && stamp - h->depth_queue.front()->stamp > 5.0) {
        h->depth_free.push_back(h->depth_queue.front());
        h->depth_queue.pop_front();
    }
    // 8. + 9. fuse and down-sample again (FN:360-371)
    uint64_t total = 0;
    for (DepthEntry* q : h->depth_queue) total += q->n;
    if (total > 0x7fffffffull) return fail(h, LVREG_ERR_INVALID, "depth cloud too large");
    CK(h->depth_concat.reserve((size_t)(total ? total : 1) * 16));
    uint64_t at = 0;
    for (DepthEntry* q : h->depth_queue) {
        if (q->n) CK(cudaMemcpyAsync(h->depth_concat.as<float4>() + at, q->pts.p, (size_t)q->n * 16, cudaMemcpyDeviceToDevice, h->st));
        at += q->n;
    }
    lanes_fork(h, 0x4);
    VgJob J2;
    J2.lane = LANE_SCAN_CORNER;
    J2.pts = h->depth_concat.as<float4>();
    J2.n = (uint32_t)total;
    J2.leaf = 0.2f;
    J2.out = &h->depth_cloud;
    J2.n_out = &h->n_depth;
    CKS(voxelgrid_batch(h, &J2, 1));
    lanes_join(h, 0x4);
    mark(h, EV_DS);
    CK(cudaStreamSynchronize(h->st));
    h->last.downsample_ms = span(h, EV_BEGIN, EV_DS);
    finish_timings(h);
    end_call(h);
    if (n_depth) *n_depth = h->n_depth;
    return LVREG_OK;
}

int lvreg_depth_set_cloud(lvreg_handle* h, const lvreg_cloud* depth_cloud) {
    if (!h || !depth_cloud) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    CKS(upload_cloud(h, depth_cloud, h->depth_cloud, h->lane[LANE_SCAN_CORNER].stage, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->n_depth = (uint32_t)depth_cloud->n;
    end_call(h);
    return LVREG_OK;
}

int lvreg_depth_get_cloud(lvreg_handle* h, int which, lvreg_cloud_out* out, size_t* n) {
    if (!h || which < 0 || which > 1) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    const uint32_t cnt = which == 0 ? h->n_depth : h->n_depth_local;
    if (n) *n = cnt;
    if (!out) return LVREG_OK;
    return download_cloud(h, (which == 0 ? h->depth_cloud : h->depth_local).as<float4>(), cnt, out);
}

int lvreg_get_depth(lvreg_handle* h, const float T_inv[12], const float* features_xyz, size_t n, int num_bins,
                    float* depth_out, float* features_3d_out) {
    if (!h || !T_inv || (n && (!features_xyz || !depth_out)) || num_bins < 1 || num_bins > 4096 || n > 0x7fffffffull)
        return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    h->n_depth_local = 0;
    for (size_t i = 0; i < n; ++i) depth_out[i] = -1.f;              // FT:121-123
    const uint32_t m = h->n_depth;
    const uint32_t nb = (uint32_t)num_bins * (uint32_t)num_bins;
    Lane& L = h->lane[LANE_SCAN_CORNER];
    uint32_t* d_total = L.small.as<uint32_t>() + SM_TOTAL;
    const Affine Ti = affine_from(T_inv);
    CK(h->depth_bins.reserve((size_t)nb * 8));
    CK(h->depth_local.reserve((size_t)nb * 16));
    CK(h->depth_unit.reserve((size_t)nb * 16));
    CK(cudaMemsetAsync(h->depth_bins.p, 0xff, (size_t)nb * 8, h->st));
    if (m) {
        depth_bin_kernel<<<nblk(m, 256), 256, 0, h->st>>>(h->depth_cloud.as<float4>(), m, Ti, num_bins,
                                                         h->depth_bins.as<unsigned long long>());
        launched(h);
    }
    CK(L.scan_temp.reserve((size_t)(scan_num_tiles(nb) + 2) * 4));
    exclusive_scan(BinFlagIn{h->depth_bins.as<unsigned long long>()},
                   BinCompactOut{h->depth_bins.as<unsigned long long>(), h->depth_cloud.as<float4>(), Ti,
                                 h->depth_local.as<float4>(), h->depth_unit.as<float4>()},
                   nb, L.scan_temp.as<uint32_t>(), d_total, h->st, &h->call_launches);
    CK(cudaMemcpyAsync(L.pinned, d_total, 4, cudaMemcpyDeviceToHost, h->st));
    if (n) {
        CK(h->depth_feat.reserve(n * 12));
        CK(h->depth_out.reserve(n * 4));
        CK(h->depth_f3d.reserve(n * 16));
        CK(cudaMemcpyAsync(h->depth_feat.p, features_xyz, n * 12, cudaMemcpyHostToDevice, h->st));
        const float bin_res = 180.0 / (float)num_bins;
        const float thr = (float)pow(sin(bin_res / 180.0 * M_PI) * 5.0, 2);                   // FT:234
        depth_feature_kernel<<<(uint32_t)n, kDepthThreads, 0, h->st>>>(h->depth_feat.as<float>(), (uint32_t)n,
                                                                      h->depth_unit.as<float4>(), d_total, thr,
                                                                      h->depth_out.as<float>(), h->depth_f3d.as<float4>());
        launched(h);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(depth_out, h->depth_out.p, n * 4, cudaMemcpyDeviceToHost, h->st));
        if (features_3d_out) CK(cudaMemcpyAsync(features_3d_out, h->depth_f3d.p, n * 16, cudaMemcpyDeviceToHost, h->st));
    }
    mark(h, EV_REG);
    CK(cudaStreamSynchronize(h->st));
    h->n_depth_local = L.pinned[0];
    h->last.register_ms = span(h, EV_BEGIN, EV_REG);
    finish_timings(h);
    end_call(h);
    return LVREG_OK;
}

// ---- deskew + range-image projection -----------------------------------------------------
int lvreg_project_cloud(lvreg_handle* h, const lvreg_raw_cloud* in, const lvreg_projection_params* prm, size_t* n_extracted) {
    if (!h || !in || !prm) return LVREG_ERR_INVALID;
    if (in->n && !in->data) return fail(h, LVREG_ERR_INVALID, "raw cloud has n > 0 but no data");
    if (in->n > 0x7fffffffull) return fail(h, LVREG_ERR_INVALID, "raw cloud too large");
    if (in->stride < 16 || (in->stride & 3) || in->intensity_offset + 4 > in->stride || (in->intensity_offset & 3) ||
        in->ring_offset + 2 > in->stride || (in->ring_offset & 1) || in->time_offset + 4 > in->stride || (in->time_offset & 3))
        return fail(h, LVREG_ERR_INVALID, "bad raw cloud layout");
    if (prm->n_scan < 1 || prm->n_scan > 255 || prm->horizon_scan < 1 || prm->downsample_rate < 1 ||
        (int64_t)prm->n_scan * prm->horizon_scan > (1 << 26))
        return fail(h, LVREG_ERR_INVALID, "bad N_SCAN / Horizon_SCAN / downsampleRate");
    if (prm->sensor < 0 || prm->sensor > 2) return fail(h, LVREG_ERR_INVALID, "unknown sensor type");
    if (prm->deskew && (prm->imu_pointer_cur < 0 || !prm->imu_time || !prm->imu_rot_x || !prm->imu_rot_y || !prm->imu_rot_z))
        return fail(h, LVREG_ERR_INVALID, "deskew requested without IMU rotation samples");
    CK(cudaSetDevice(h->device));
    begin_call(h);
    mark(h, EV_BEGIN);
    const uint32_t n = (uint32_t)in->n;
    const int ns = prm->n_scan, H = prm->horizon_scan;
    const uint32_t cells = (uint32_t)ns * (uint32_t)H;
    Lane& L = h->lane[LANE_SCAN_SURF];
    h->n_proj = 0;
    h->proj_n_scan = ns;
    h->proj_start_host.assign(ns, 4);        // empty scan: count = 0 everywhere (IP:631, 646)
    h->proj_end_host.assign(ns, -6);
    if (n_extracted) *n_extracted = 0;
    // inputs
    RawLayout raw;
    raw.stride = in->stride; raw.intensity_off = in->intensity_offset; raw.ring_off = in->ring_offset; raw.time_off = in->time_offset;
    if (in->on_device) raw.data = (const uint8_t*)in->data;
    else {
        CK(h->proj_raw.reserve((size_t)(n ? n : 1) * in->stride));
        if (n) CK(cudaMemcpyAsync(h->proj_raw.p, in->data, (size_t)n * in->stride, cudaMemcpyHostToDevice, h->st));
        raw.data = h->proj_raw.as<uint8_t>();
    }
    ProjParams P;
    P.n_scan = ns; P.horizon = H; P.downsample_rate = prm->downsample_rate; P.sensor = prm->sensor;
    P.min_range = prm->lidar_min_range; P.max_range = prm->lidar_max_range;
    P.deskew = prm->deskew ? 1 : 0; P.imu_pointer_cur = prm->imu_pointer_cur; P.time_scan_cur = prm->time_scan_cur;
    P.imu_time = P.imu_rx = P.imu_ry = P.imu_rz = nullptr;
    if (P.deskew) {
        const size_t cnt = (size_t)prm->imu_pointer_cur + 1;
        CK(h->proj_imu.reserve(cnt * 4 * 8));
        double* d = h->proj_imu.as<double>();
        CK(cudaMemcpyAsync(d, prm->imu_time, cnt * 8, cudaMemcpyHostToDevice, h->st));
        CK(cudaMemcpyAsync(d + cnt, prm->imu_rot_x, cnt * 8, cudaMemcpyHostToDevice, h->st));
        CK(cudaMemcpyAsync(d + 2 * cnt, prm->imu_rot_y, cnt * 8, cudaMemcpyHostToDevice, h->st));
        CK(cudaMemcpyAsync(d + 3 * cnt, prm->imu_rot_z, cnt * 8, cudaMemcpyHostToDevice, h->st));
        P.imu_time = d; P.imu_rx = d + cnt; P.imu_ry = d + 2 * cnt; P.imu_rz = d + 3 * cnt;
    }
    // scratch: [0,256) ring counts, [256,512) ring starts, [512] first index, [513] total, [520..) row prefixes,
    // then transStartInverse
    const size_t small_words = 520 + (size_t)ns + 8;
    CK(h->proj_small.reserve(small_words * 4 + sizeof(Affine) + 64));
    uint32_t* sm = h->proj_small.as<uint32_t>();
    uint32_t* d_ring_count = sm, *d_ring_start = sm + 256, *d_first = sm + 512, *d_total = sm + 513, *d_row_prefix = sm + 520;
    Affine* d_start_inv = reinterpret_cast<Affine*>(reinterpret_cast<uint8_t*>(sm) + ((small_words * 4 + 63) & ~(size_t)63));
    CK(h->proj_pts.reserve((size_t)(n ? n : 1) * 16));
    CK(h->proj_rangein.reserve((size_t)(n ? n : 1) * 4));
    CK(h->proj_colin.reserve((size_t)(n ? n : 1) * 4));
    CK(h->proj_owner.reserve((size_t)(cells + 8) * 4));
    CK(h->proj_cloud.reserve((size_t)cells * 16));
    CK(h->proj_range.reserve((size_t)cells * 4));
    CK(h->proj_col.reserve((size_t)cells * 4));
    CK(h->proj_rings.reserve((size_t)ns * 8));
    CKS(ensure_sort_buffers(h, L, n ? n : 1));
    CK(L.scan_temp.reserve((size_t)(scan_num_tiles(cells) + 2) * 4));
    CK(cudaMemsetAsync(sm, 0, 512 * 4, h->st));
    CK(cudaMemsetAsync(d_first, 0xff, 4, h->st));
    CK(cudaMemsetAsync(h->proj_owner.p, 0xff, (size_t)(cells + 8) * 4, h->st));
    if (n) {
        uint32_t* keys = L.keys[0].as<uint32_t>();
        uint32_t* vals = L.vals[0].as<uint32_t>();
        proj_classify_kernel<<<nblk(n, 256), 256, 0, h->st>>>(raw, n, P, h->proj_pts.as<float4>(), h->proj_rangein.as<float>(), keys,
                                                            vals, h->proj_colin.as<int32_t>(), d_ring_count);
        launched(h);
        if (P.sensor == 2) {
            // columnIdnCountVec: rank of the point among the earlier accepted points of its ring
            proj_ring_starts_kernel<<<1, 32, 0, h->st>>>(d_ring_count, ns, d_ring_start);
            // the sort permutes (keys, vals); the claim kernel needs the unsorted keys: sort a copy
            CK(cudaMemcpyAsync(L.keys[1].p, keys, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->st));
            CK(cudaMemcpyAsync(L.vals[1].p, vals, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->st));
            CK(h->feat_idx.reserve((size_t)n * 4));
            CK(h->feat_pidx.reserve((size_t)n * 4));
            int cur = radix_sort_pairs(L.keys[1].as<uint32_t>(), L.vals[1].as<uint32_t>(), h->feat_idx.as<uint32_t>(),
                                       h->feat_pidx.as<uint32_t>(), n, 8, L.sort_scratch.as<uint32_t>(), h->st, &h->call_launches);
            const uint32_t* sk = cur ? h->feat_idx.as<uint32_t>() : L.keys[1].as<uint32_t>();
            const uint32_t* sv = cur ? h->feat_pidx.as<uint32_t>() : L.vals[1].as<uint32_t>();
            proj_livox_columns_kernel<<<nblk(n, 256), 256, 0, h->st>>>(sk, sv, n, d_ring_start, h->proj_colin.as<int32_t>());
            launched(h, 2);
        }
        proj_claim_kernel<<<nblk(n, 256), 256, 0, h->st>>>(keys, h->proj_colin.as<int32_t>(), n, H, h->proj_owner.as<uint32_t>(), d_first);
        proj_start_kernel<<<1, 32, 0, h->st>>>(P, raw, d_first, d_start_inv);
        launched(h, 2);
    }
    CellExtractOut out;
    out.owner = h->proj_owner.as<uint32_t>(); out.pts = h->proj_pts.as<float4>(); out.range = h->proj_rangein.as<float>();
    out.in = raw; out.P = P; out.start_inv = d_start_inv; out.extracted = h->proj_cloud.as<float4>();
    out.point_range = h->proj_range.as<float>(); out.point_col_ind = h->proj_col.as<int32_t>(); out.row_prefix = d_row_prefix;
    exclusive_scan(CellFlagIn{h->proj_owner.as<uint32_t>()}, out, cells, L.scan_temp.as<uint32_t>(), d_total, h->st, &h->call_launches);
    int32_t* d_start = h->proj_rings.as<int32_t>();
    int32_t* d_end = d_start + ns;
    proj_ring_index_kernel<<<nblk((uint32_t)ns, 128), 128, 0, h->st>>>(d_row_prefix, d_total, ns, d_start, d_end);
    launched(h);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(L.pinned, d_total, 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->proj_start_host.data(), d_start, (size_t)ns * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->proj_end_host.data(), d_end, (size_t)ns * 4, cudaMemcpyDeviceToHost, h->st));
    mark(h, EV_DS);
    CK(cudaStreamSynchronize(h->st));
    h->n_proj = L.pinned[0];
    h->last.downsample_ms = span(h, EV_BEGIN, EV_DS);
    finish_timings(h);
    end_call(h);
    if (n_extracted) *n_extracted = h->n_proj;
    return LVREG_OK;
}

int lvreg_get_projection(const lvreg_handle* h, lvreg_cloud* extracted, lvreg_scan_info* info) {
    if (!h || !extracted || !info) return LVREG_ERR_INVALID;
    if (h->proj_n_scan == 0) return LVREG_ERR_INVALID;
    extracted->data = h->proj_cloud.p;
    extracted->n = h->n_proj;
    extracted->stride = 16;
    extracted->intensity_offset = 12;
    extracted->on_device = 1;
    extracted->reserved = 0;
    info->start_ring_index = h->proj_start_host.data();
    info->end_ring_index = h->proj_end_host.data();
    info->n_scan = h->proj_n_scan;
    info->reserved = 0;
    info->point_col_ind = h->proj_col.as<int32_t>();
    info->point_range = h->proj_range.as<float>();
    return LVREG_OK;
}

int lvreg_download_projection(lvreg_handle* h, lvreg_cloud_out* extracted, float* point_range, int32_t* point_col_ind,
                              int32_t* start_ring_index, int32_t* end_ring_index, size_t* n) {
    if (!h) return LVREG_ERR_INVALID;
    if (h->proj_n_scan == 0) return fail(h, LVREG_ERR_INVALID, "no projected scan");
    CK(cudaSetDevice(h->device));
    if (n) *n = h->n_proj;
    const uint32_t m = h->n_proj;
    if (start_ring_index) memcpy(start_ring_index, h->proj_start_host.data(), (size_t)h->proj_n_scan * 4);
    if (end_ring_index) memcpy(end_ring_index, h->proj_end_host.data(), (size_t)h->proj_n_scan * 4);
    if (point_range && m) CK(cudaMemcpyAsync(point_range, h->proj_range.p, (size_t)m * 4, cudaMemcpyDeviceToHost, h->st));
    if (point_col_ind && m) CK(cudaMemcpyAsync(point_col_ind, h->proj_col.p, (size_t)m * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (extracted) return download_cloud(h, h->proj_cloud.as<float4>(), m, extracted);
    return LVREG_OK;
}

// ---- measurement ---------------------------------------------------------------------------------
int lvreg_get_timings(const lvreg_handle* h, lvreg_timings* t) {
    if (!h || !t) return LVREG_ERR_INVALID;
    *t = h->last;
    return LVREG_OK;
}

int lvreg_get_iteration_profile(const lvreg_handle* h, float* us, int* iterations) {
    if (!h || !us) return LVREG_ERR_INVALID;
    const RegOut* ho = (const RegOut*)((const char*)h->pinned + 4096);    // last D2H copy of RegOut
    int n = ho->iterations;
    if (n < 0) n = 0;
    if (n > h->prm.max_iters) n = h->prm.max_iters;
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 4; ++k)
            us[i * 4 + k] = (float)((double)(ho->stamp[i][k + 1] - ho->stamp[i][k]) * 1e-3);
    if (iterations) *iterations = n;
    return LVREG_OK;
}

int lvreg_debug_tile_times(lvreg_handle* h, uint32_t* ns_out, size_t cap, size_t* n_tiles) {
    if (!h || !n_tiles) return LVREG_ERR_INVALID;
    *n_tiles = h->debug_tiles ? h->debug_ntiles : 0;
    if (!ns_out || !*n_tiles) return LVREG_OK;
    CK(cudaSetDevice(h->device));
    size_t n = *n_tiles < cap ? *n_tiles : cap;
    CK(cudaMemcpyAsync(ns_out, h->tilens.p, n * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return LVREG_OK;
}

int lvreg_enable_kernel_timing(lvreg_handle* h, int on) {
    if (!h) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (on && !h->kt_ev[0][0])
        for (int l = 0; l < 2; ++l)
            for (int k = 0; k < 2; ++k) CK(cudaEventCreate(&h->kt_ev[l][k]));
    h->kernel_timing = on != 0;
    return LVREG_OK;
}

int lvreg_get_bucket_kernel_ms(lvreg_handle* h, float ms[2], uint32_t n_in[2], uint32_t n_out[2]) {
    if (!h || !ms || !n_in || !n_out) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->st));
    for (int l = 0; l < 2; ++l) {
        ms[l] = 0.f;
        n_in[l] = h->kt_set[l] ? h->kt_n_in[l] : 0u;
        n_out[l] = h->kt_set[l] ? h->map[l].m : 0u;
        if (h->kt_set[l]) CK(cudaEventElapsedTime(&ms[l], h->kt_ev[l][0], h->kt_ev[l][1]));
    }
    return LVREG_OK;
}

int lvreg_debug_stage_stats(lvreg_handle* h, uint32_t out[8]) {
    if (!h || !out) return LVREG_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, h->stagestats.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    out[5] = (uint32_t)((const RegOut*)((const char*)h->pinned + 4096))->pad;   // 1 = Cholesky shortcut, 2 = 6x6 Jacobi
    out[6] = h->vg_bucket_fallbacks;      // bucketed VoxelGrid jobs redone by the device-wide sort (a bucket overflowed)
    out[7] = h->vg_bucket_jobs;           // VoxelGrid jobs that went through the bucketed path
    return LVREG_OK;
}

int lvreg_get_launch_count(const lvreg_handle* h, uint64_t* n) {
    if (!h || !n) return LVREG_ERR_INVALID;
    *n = h->total_launches;
    return LVREG_OK;
}

}  // extern "C"
