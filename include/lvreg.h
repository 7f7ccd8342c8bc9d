/*
 * lvreg.h -- C ABI of the B200-native scan-to-map registration library (liblvreg.so).
 *
 * Drop-in boundary for ONE path of valentinomario/LiDAR-Visual-Inertial-SLAM: the
 * LIO-SAM-derived scan-to-map registration inside class mapOptimization.  The reference has
 * no plugin/FFI interface; the hot path is a set of void member functions that talk through
 * public data members.  Every entry point below names the reference
 * member function / data member it replaces.  "MO:" = lidar_odometry/src/mapOptimization.cpp:
 *
 * Conventions
 *   - points: any AoS layout with float x,y,z at byte offsets 0/4/8 and a float intensity at
 *     `intensity_offset`; pcl::PointXYZI (utility.h:64) is {stride 32, intensity_offset 16},
 *     packed float4 is {16, 12}.  Pass PCL clouds as cloud->points.data() without a copy.
 *   - poses: float[6] = {roll, pitch, yaw, x, y, z} = transformTobeMapped order (MO:126).
 *     Keyframe poses (PointTypePose, MO:29-46) are passed in the same order.
 *   - every call is synchronous with respect to the host; all device work of one handle runs
 *     on one CUDA stream.  A handle is not thread-safe; different handles are independent
 *     (one per robot / sequence / GPU), mirroring the reference's one-mutex design (MO:309).
 *   - status codes: 0 = ok; the two "soft" outcomes of the reference (not enough features,
 *     MO:1340-1342; no key poses, MO:1317) return a distinct code and leave the pose untouched.
 *   - there is no CPU fallback: without a CUDA device lvreg_create fails.
 */
#ifndef LVREG_H
#define LVREG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVREG_OK                       0
#define LVREG_ERR_INVALID              1   /* bad argument */
#define LVREG_ERR_CUDA                 2   /* CUDA runtime error, see lvreg_last_error */
#define LVREG_ERR_NOT_ENOUGH_FEATURES  3   /* MO:1320/1340: Nc <= 10 or Ns <= 100; pose unchanged */
#define LVREG_ERR_NO_KEYFRAMES         4   /* MO:1317: cloudKeyPoses3D empty; pose unchanged */
#define LVREG_ERR_NO_MAP               5   /* scan2map before a local map was built */
#define LVREG_ERR_CAPACITY             6   /* output buffer too small */

#define LVREG_MAX_ITERS 32

#define LVREG_CORNER 0
#define LVREG_SURF   1

/* kNN variants of lvreg_knn5 */
#define LVREG_KNN_GRID_GATED 0   /* 27-cell search: exact whenever the 5th neighbour is inside the gate */
#define LVREG_KNN_GRID_EXACT 1   /* ring expansion until provably exact (unbounded) */
#define LVREG_KNN_BRUTE      2   /* exhaustive, FP32-pipe bound */
#define LVREG_KNN_GRID_STAGED 3  /* GATED semantics through the registration kernel's search: 32-query tiles, the
                                    cell box of a tile staged in shared memory by bulk copies (cp.async.bulk) */

typedef struct lvreg_handle lvreg_handle;

/* A read-only point cloud in host (on_device = 0) or device (on_device = 1) memory. */
typedef struct lvreg_cloud {
    const void* data;
    size_t      n;
    uint32_t    stride;            /* bytes between consecutive points (>= 16, multiple of 4) */
    uint32_t    intensity_offset;  /* byte offset of the float intensity inside a point */
    int32_t     on_device;
    int32_t     reserved;
} lvreg_cloud;

/* A writable point buffer. */
typedef struct lvreg_cloud_out {
    void*    data;
    size_t   capacity;             /* in points */
    uint32_t stride;
    uint32_t intensity_offset;
    int32_t  on_device;
    int32_t  reserved;
} lvreg_cloud_out;

/* Mirrors the ParamServer fields and the hard-coded literals of the hot path. */
typedef struct lvreg_params {
    float   corner_leaf;         /* mappingCornerLeafSize   utility.h:266   0.2 */
    float   surf_leaf;           /* mappingSurfLeafSize     utility.h:268   0.4 */
    int32_t edge_min_valid;      /* edgeFeatureMinValidNum  utility.h:259   10  */
    int32_t surf_min_valid;      /* surfFeatureMinValidNum  utility.h:261   100 */
    int32_t max_iters;           /* MO:1325  20 (<= LVREG_MAX_ITERS) */
    float   knn_gate_sq;         /* MO:1025, MO:1121  1.0 */
    float   line_eig_ratio;      /* MO:1052  3 */
    float   plane_tol;           /* MO:1142  0.2 */
    float   min_weight;          /* MO:1088, MO:1159  0.1 */
    int32_t min_matches;         /* MO:1210  50 */
    float   degeneracy_eig;      /* MO:1272  100 */
    float   conv_deg;            /* MO:1309  0.05 */
    float   conv_cm;             /* MO:1309  0.05 */
    int32_t reference_quirks;    /* 1 = reproduce the shadowed matP (MO:1220 hides MO:132) */
    float   rotation_tolerance;  /* rotation_tollerance utility.h:273  1000 */
    float   z_tolerance;         /* z_tollerance        utility.h:271  1000 */
    float   imu_rpy_weight;      /* imuRPYWeight        utility.h:234  0.01 */
    int32_t reserved[7];
} lvreg_params;

/* Outcome of one scan2MapOptimization (MO:1315-1343). */
typedef struct lvreg_result {
    int32_t iterations;                       /* iterCount reached (1..max_iters) */
    int32_t converged;                        /* LMOptimization returned true */
    int32_t degenerate;                       /* isDegenerate (MO:131) after the call */
    int32_t n_corner_ds, n_surf_ds;           /* laserCloud{Corner,Surf}LastDSNum */
    int32_t n_corner_map, n_surf_map;         /* laserCloud{Corner,Surf}FromMapDSNum */
    int32_t n_sel[LVREG_MAX_ITERS];           /* laserCloudSelNum per iteration */
    float   pose_iter[LVREG_MAX_ITERS][6];    /* transformTobeMapped after each iteration */
    float   cost[LVREG_MAX_ITERS];            /* sum of squared residuals per iteration (diagnostic) */
} lvreg_result;

typedef struct lvreg_map_info {
    uint64_t n_corner_in, n_surf_in;          /* concatenated keyframe points (laserCloud*FromMap) */
    uint64_t n_corner_ds, n_surf_ds;          /* after VoxelGrid (laserCloud*FromMapDS) */
    int32_t  grid_dims[2][3];                 /* search-grid cells per axis, corner / surf */
    float    grid_cell[2];                    /* search-grid cell edge in metres */
} lvreg_map_info;

/* Device time (CUDA events on the handle's stream) of the stages of the last call, in ms. */
typedef struct lvreg_timings {
    float upload_ms;        /* H2D + pack of the call's input clouds */
    float downsample_ms;    /* downsampleCurrentScan */
    float map_build_ms;     /* transform + concat + VoxelGrid of the local map */
    float grid_build_ms;    /* search-grid build (replaces kdtree->setInputCloud) */
    float register_ms;      /* the on-device LM loop */
    float total_ms;         /* first event to last event of the call */
    int32_t kernel_launches;/* kernels launched by the last call */
    int32_t reserved;
} lvreg_timings;

/* ---- lifetime ---------------------------------------------------------------------------- */
void        lvreg_default_params(lvreg_params* p);
/* `cuda_stream` is a cudaStream_t (may be NULL = a private non-blocking stream). */
int         lvreg_create(const lvreg_params* p, int device, void* cuda_stream, lvreg_handle** out);
void        lvreg_destroy(lvreg_handle* h);
const char* lvreg_last_error(const lvreg_handle* h);
const char* lvreg_status_string(int status);
int         lvreg_version(void);
/* page-locked host memory for callers that want full-speed uploads */
void*       lvreg_host_alloc(size_t bytes);
void        lvreg_host_free(void* p);
/* page-lock / unlock memory the caller already owns (e.g. the points of a pcl::PointCloud): uploads from it are
 * then asynchronous DMA transfers instead of staged copies */
int         lvreg_host_register(void* p, size_t bytes);
int         lvreg_host_unregister(void* p);

/* ---- keyframe store: cornerCloudKeyFrames / surfCloudKeyFrames / cloudKeyPoses6D (MO:83-87) -- */
/* saveKeyFramesAndFactor's push_back of the DS feature clouds + pose (MO:1600-1610). */
/* Pre-sizes every per-call buffer (upper bounds; 0 skips a group) so that no later call allocates device
 * memory: map_points_* = points entering the local-map VoxelGrid per class (sum over the selected
 * keyframes), scan_points_* = raw feature points per scan, max_grid_cells = cells of a search grid
 * (<= 2^26).  Without it the buffers grow geometrically on demand (each growth synchronises the device). */
int lvreg_reserve(lvreg_handle* h, size_t map_points_corner, size_t map_points_surf, size_t scan_points_corner,
                  size_t scan_points_surf, size_t max_grid_cells);
int lvreg_add_keyframe(lvreg_handle* h, const lvreg_cloud* corner, const lvreg_cloud* surf,
                       const float pose_rpyxyz[6], int32_t* id_out);
/* Same, taking the device-resident laserCloud{Corner,Surf}LastDS of the current scan (the
 * pcl::copyPointCloud calls of MO:1600-1606) -- no host round trip. */
int lvreg_add_keyframe_from_scan(lvreg_handle* h, const float pose_rpyxyz[6], int32_t* id_out);
/* correctPoses after a loop closure (MO:1623-1640): replaces the first n keyframe poses. */
int lvreg_update_keyframe_poses(lvreg_handle* h, const float* poses_rpyxyz, size_t n);
int lvreg_num_keyframes(const lvreg_handle* h, size_t* n);
int lvreg_clear_keyframes(lvreg_handle* h);

/* ---- local map: extractCloud (MO:931-970) ------------------------------------------------- */
/* ids = keyframe indices in concatenation order (output of extractNearby MO:894-929 after the
 * distance filter MO:938); duplicates allowed, as in the reference.  Transforms each keyframe by
 * its pose (transformPointCloud MO:347-366), concatenates, VoxelGrid-filters corner/surf with
 * corner_leaf/surf_leaf (MO:959-965) and builds the 5-NN search grids (replaces MO:1322-1323). */
int lvreg_build_local_map(lvreg_handle* h, const int32_t* ids, size_t n, lvreg_map_info* info);
/* Install already down-sampled maps (laserCloud{Corner,Surf}FromMapDS) and build the grids. */
int lvreg_set_local_map(lvreg_handle* h, const lvreg_cloud* corner_ds, const lvreg_cloud* surf_ds,
                        lvreg_map_info* info);
/* Read back laserCloud{Corner,Surf}FromMapDS (which = LVREG_CORNER / LVREG_SURF). */
int lvreg_get_local_map(lvreg_handle* h, int which, lvreg_cloud_out* out, size_t* n);

/* ---- current scan ------------------------------------------------------------------------- */
/* downsampleCurrentScan (MO:987-999): VoxelGrid of laserCloud{Corner,Surf}Last; the results
 * (laserCloud{Corner,Surf}LastDS) stay on the device for lvreg_scan2map. */
int lvreg_downsample_scan(lvreg_handle* h, const lvreg_cloud* corner_raw, const lvreg_cloud* surf_raw,
                          size_t* n_corner_ds, size_t* n_surf_ds);
/* Install already down-sampled feature clouds as laserCloud{Corner,Surf}LastDS. */
int lvreg_set_scan_ds(lvreg_handle* h, const lvreg_cloud* corner_ds, const lvreg_cloud* surf_ds);
int lvreg_get_scan_ds(lvreg_handle* h, int which, lvreg_cloud_out* out, size_t* n);

/* ---- registration ------------------------------------------------------------------------- */
/* scan2MapOptimization (MO:1315-1343) on the current DS scan and local map: up to max_iters of
 * {cornerOptimization, surfOptimization, combineOptimizationCoeffs, LMOptimization} in ONE
 * cooperative kernel launch, then transformUpdate's clamps (MO:1370-1372; IMU slerp via
 * lvreg_transform_update).  pose is transformTobeMapped, in/out. */
int lvreg_scan2map(lvreg_handle* h, float pose_rpyxyz[6], lvreg_result* res);
/* One call per incoming scan = extractCloud (if ids != NULL) + downsampleCurrentScan +
 * scan2MapOptimization, i.e. the body of laserCloudInfoHandler MO:318-322. */
int lvreg_register_scan(lvreg_handle* h, const lvreg_cloud* corner_raw, const lvreg_cloud* surf_raw,
                        const int32_t* ids, size_t n_ids, float pose_rpyxyz[6], lvreg_result* res);
/* transformUpdate (MO:1345-1375) as a stand-alone stage: optional IMU roll/pitch slerp, THEN the clamps of
 * roll / pitch / z (the reference's order).  Host arithmetic. */
int lvreg_transform_update(const lvreg_handle* h, float pose_rpyxyz[6], int imu_available,
                           float imu_roll_init, float imu_pitch_init);
/* cloudInfo.imu_available / imu_roll_init / imu_pitch_init of the scan about to be registered (MO:1347-1366).
 * lvreg_scan2map / lvreg_register_scan end with transformUpdate exactly like scan2MapOptimization (MO:1339): blend
 * towards this IMU attitude when available, then clamp once.  Sticky until changed; default: not available. */
int lvreg_set_imu_prior(lvreg_handle* h, int imu_available, float imu_roll_init, float imu_pitch_init);
/* isDegenerate (MO:131) persists across scans in the reference; read / reset it here. */
int lvreg_get_degenerate(const lvreg_handle* h, int* is_degenerate);
int lvreg_reset_lm_state(lvreg_handle* h);

/* ---- stage-level entry points (parity tests, micro-benchmarks) ----------------------------- */
/* pcl::getTransformation (MO:399-407) on the host, row-major 3x4. */
void lvreg_pose_to_affine(const float pose_rpyxyz[6], float T[12]);
/* transformPointCloud (MO:347-385). */
int lvreg_transform_cloud(lvreg_handle* h, const lvreg_cloud* in, const float pose_rpyxyz[6],
                          lvreg_cloud_out* out);
/* pcl::VoxelGrid::filter as used at MO:959-965 / MO:991-997.  voxel_keys_out (optional, host,
 * out->capacity entries) receives the voxel idx of every output point; passthrough is set when
 * PCL's "leaf size too small" rule returns the input unchanged. */
int lvreg_voxelgrid(lvreg_handle* h, const lvreg_cloud* in, float leaf, lvreg_cloud_out* out,
                    size_t* n_out, uint32_t* voxel_keys_out, int* passthrough);
/* per-input-point voxel idx (the sort key of VoxelGrid); keys_out is host memory, in->n entries */
int lvreg_voxel_keys(lvreg_handle* h, const lvreg_cloud* in, float leaf, uint32_t* keys_out);
/* kdtree->nearestKSearch(q, 5, ...) (MO:1019, MO:1111) for world-frame queries against the current
 * corner / surf map.  idx_out: n x 5 int32 (indices into the DS map), d2_out: n x 5 float, host. */
int lvreg_knn5(lvreg_handle* h, int which_map, const lvreg_cloud* queries, int variant,
               int32_t* idx_out, float* d2_out);
/* cornerOptimization / surfOptimization (MO:1006-1167) for sensor-frame points at `pose`:
 * coeff_out n x 4 (coeffSel: s*la, s*lb, s*lc, s*ld2; zero when rejected), flag_out n bytes
 * (laserCloudOri*Flag), knn_idx_out optional n x 5.  All host memory. */
int lvreg_corner_residuals(lvreg_handle* h, const lvreg_cloud* pts, const float pose_rpyxyz[6],
                           float* coeff_out, uint8_t* flag_out, int32_t* knn_idx_out);
int lvreg_surf_residuals(lvreg_handle* h, const lvreg_cloud* pts, const float pose_rpyxyz[6],
                         float* coeff_out, uint8_t* flag_out, int32_t* knn_idx_out);
/* LMOptimization(iterCount) (MO:1190-1313) on explicit laserCloudOri / coeffSel rows (packed
 * float4 host arrays).  Outputs optional.  *converged = return value of LMOptimization. */
int lvreg_lm_step(lvreg_handle* h, const float* ori_xyzi, const float* coeff_xyzi, size_t n_sel,
                  int iter_count, float pose_rpyxyz[6], float AtA_out[36], float Atb_out[6],
                  float x_out[6], int* converged);

/* ---- "next" row: FeatureExtraction on the device ----------------------------- */
/* The per-point side channels of lidar_odometry/msg/CloudInfo.msg that FeatureExtraction consumes
 * (imageProjection.cpp:624-647).  The ring indices are host arrays; point_col_ind / point_range may be
 * host or device pointers (lvreg_get_projection hands out device ones). */
typedef struct lvreg_scan_info {
    const int32_t* start_ring_index;   /* n_scan entries */
    const int32_t* end_ring_index;     /* n_scan entries */
    int32_t        n_scan;             /* N_SCAN, <= 256 */
    int32_t        reserved;
    const int32_t* point_col_ind;      /* one per point of the deskewed cloud */
    const float*   point_range;        /* one per point */
} lvreg_scan_info;
/* calculateSmoothness + markOccludedPoints + extractFeatures (featureExtraction.cpp:87-245) on the
 * deskewed, ring-ordered cloud: corner cloud (<= 40 per ring sector, ring / sector / pick order) and
 * surface cloud (per-ring VoxelGrid with `surf_leaf` = odometrySurfLeafSize, rings in order).
 * Unspecified reference behaviour is pinned: std::sort ties break on the point index; entries the
 * reference never initialises read as zero.  corner / surf (optional) receive host or device
 * copies; label_out (optional, host, one int32 per point) receives cloudLabel.  The two clouds also
 * stay on the device as laserCloudCornerLast / laserCloudSurfLast -- see lvreg_get_feature_clouds. */
int lvreg_extract_features(lvreg_handle* h, const lvreg_cloud* deskewed, const lvreg_scan_info* info,
                           float edge_threshold, float surf_threshold, float surf_leaf,
                           lvreg_cloud_out* corner, size_t* n_corner, lvreg_cloud_out* surf, size_t* n_surf,
                           int32_t* label_out);
/* Device-resident descriptors of the last extracted feature clouds; pass them to
 * lvreg_register_scan / lvreg_downsample_scan to go from the raw scan to the pose without a host
 * round trip of the features. */
int lvreg_get_feature_clouds(const lvreg_handle* h, lvreg_cloud* corner, lvreg_cloud* surf);

/* ---- "next" row: loop-closure registration --------------------------------------
 * Replaces, for the loop-closure thread (MO:523-535 -> performLoopClosure MO:549-628), everything
 * between the candidate pair and the pose constraint:
 *   loopFindNearKeyframes  MO:719-741   keyframes [key-n, key+n] (corner then surf of each, under
 *                                       the stored poses) concatenated + downSizeFilterICP
 *   pcl::IterativeClosestPoint MO:578-590  point-to-point, 1-NN correspondences within
 *                                       max_corr_dist, Umeyama/SVD estimate, PCL's
 *                                       DefaultConvergenceCriteria; align() with identity guess
 *   getFitnessScore / gates MO:572,592  submap sizes >= 300 / 1000, converged, fitness <= gate
 *   pose correction        MO:600-609   tCorrect = icp.getFinalTransformation() * tWrong
 * The gtsam::Pose3 algebra of MO:610-619 (poseFrom.between(poseTo), the noise model) and the
 * candidate search (MO:630-661, a radius search over a few hundred key poses) stay on the host:
 * see host/map_optimization.cpp.  The 3x3 cross-covariance is accumulated in double and decomposed
 * by a one-sided Jacobi SVD in double (Eigen::umeyama works in float; the difference is float
 * rounding -- stated in DESIGN.md). */
enum { LVREG_ICP_NOT_CONVERGED = 0, LVREG_ICP_ITERATIONS = 1, LVREG_ICP_TRANSFORM = 2, LVREG_ICP_ABS_MSE = 3,
       LVREG_ICP_REL_MSE = 4, LVREG_ICP_NO_CORRESPONDENCES = 5, LVREG_ICP_NO_INPUT = 6 };
enum { LVREG_LOOP_OK = 0, LVREG_LOOP_SUBMAP_TOO_SMALL = 1, LVREG_LOOP_NOT_CONVERGED = 2,
       LVREG_LOOP_FITNESS_TOO_HIGH = 3 };
typedef struct lvreg_icp_params {
    float max_corr_dist;               /* setMaxCorrespondenceDistance, historyKeyframeSearchRadius * 2 */
    int32_t max_iterations;            /* setMaximumIterations */
    double transformation_epsilon;     /* setTransformationEpsilon */
    double euclidean_fitness_epsilon;  /* setEuclideanFitnessEpsilon */
    float reserved[4];
} lvreg_icp_params;
typedef struct lvreg_icp_result {
    int32_t converged;                 /* hasConverged() */
    int32_t iterations;                /* nr_iterations_ */
    int32_t state;                     /* LVREG_ICP_*: which criterion ended the loop */
    int32_t n_correspondences;         /* kept pairs in the last iteration */
    double fitness;                    /* getFitnessScore() */
    double mse;                        /* mean squared correspondence distance of the last iteration */
    float final_transformation[16];    /* getFinalTransformation(), row-major 4x4 */
} lvreg_icp_result;
typedef struct lvreg_loop_result {
    int32_t status;                    /* LVREG_LOOP_* */
    int32_t n_source, n_target;        /* cureKeyframeCloud / prevKeyframeCloud sizes */
    int32_t reserved;
    lvreg_icp_result icp;
    float pose_from[6];                /* corrected pose of key_cur {roll,pitch,yaw,x,y,z}: Pose3(RzRyRx, Point3) of MO:606 */
    float pose_to[6];                  /* stored pose of key_pre, MO:611 */
    float noise;                       /* (float)getFitnessScore(), the diagonal of the constraint noise, MO:613 */
    float reserved2;
} lvreg_loop_result;
void lvreg_icp_default_params(lvreg_icp_params* p);
/* loopFindNearKeyframes into slot 0 (ICP source) or 1 (ICP target, also builds its search grid) */
int lvreg_loop_find_near_keyframes(lvreg_handle* h, int key, int search_num, int slot, size_t* n_out);
/* Global map for visualisation / saving (publishGlobalMap MO:493-508; saveMapService MO:199-231): the
 * clouds selected by `which` (1 corner, 2 surf, 3 corner then surf of each keyframe) of the listed
 * keyframes under their stored poses, concatenated in list order, then one VoxelGrid with `leaf`
 * (globalMapVisualizationLeafSize / the service's resolution; leaf = 0: no VoxelGrid, the concatenation itself,
 * as saveMapService does for resolution 0).  The result replaces slot 0; read it back
 * with lvreg_icp_get_cloud(h, 0, ...).  The key-pose selection (radius search + pose-density VoxelGrid,
 * MO:476-491) is host logic over a few hundred poses. */
int lvreg_build_global_map(lvreg_handle* h, const int32_t* ids, size_t n_ids, int which, float leaf, size_t* n_out);
/* setInputSource (slot 0) / setInputTarget (slot 1) from a caller-provided cloud */
int lvreg_icp_set_cloud(lvreg_handle* h, int slot, const lvreg_cloud* cloud);
int lvreg_icp_get_cloud(lvreg_handle* h, int slot, lvreg_cloud_out* out, size_t* n);
/* exact 1-NN of every query in the ICP target ((d2, index) tie-break); idx -1 / d2 +inf when the
 * target is empty or nothing lies within max_dist (<= 0: unbounded).  Stage-level parity entry. */
int lvreg_nn1(lvreg_handle* h, const lvreg_cloud* queries, float max_dist, int32_t* idx_out, float* d2_out);
/* icp.align() + icp.getFitnessScore() on the two slots */
int lvreg_icp_align(lvreg_handle* h, const lvreg_icp_params* prm, lvreg_icp_result* res);
/* tCorrect = correction * pclPointToAffine3f(pose) -> {roll,pitch,yaw,x,y,z}; host arithmetic */
int lvreg_correct_pose(const float* correction4x4, const float pose[6], float out[6]);
/* performLoopClosure from the submaps on (MO:566-613) for a given candidate pair */
int lvreg_perform_loop_closure(lvreg_handle* h, int key_cur, int key_pre, int search_num,
                               const lvreg_icp_params* prm, float fitness_gate, lvreg_loop_result* out);

/* ---- "next" row: LiDAR depth for tracked visual features --------------------------
 * Replaces the point-cloud work of the visual front end:
 *   lidar_callback  feature_tracker/src/feature_tracker_node.cpp:303-371  0.2 m VoxelGrid of the new
 *       cloud, camera-view filter, transform by transNow into the odometry frame, 5 s queue, fuse,
 *       0.2 m VoxelGrid -> depthCloud (stays on the device)
 *   DepthRegister::get_depth  feature_tracker/src/feature_tracker.h:150-283  depthCloud into the
 *       camera frame (T_inv = transNow.inverse(), row-major 3x4, computed by the caller from tf),
 *       num_bins x num_bins range image keeping the closest point per bin, unit-sphere projection,
 *       exact 3-NN per feature, ray / plane intersection with the reference's clamps.
 * features_xyz: n x 3 undistorted normalised image coordinates (z = 1), as features_2d.
 * depth_out[n]: depth_of_point.values (-1 = none).  features_3d_out (optional, n x 4 floats):
 * features_3d_sphere as published on /vins/depth/depth_feature.  atan2 is evaluated in double and
 * rounded to float. */
int lvreg_depth_clear(lvreg_handle* h);
int lvreg_depth_add_cloud(lvreg_handle* h, const lvreg_cloud* cloud, const float T_now[12], double stamp,
                          size_t* n_depth);
/* stage level: set depthCloud directly */
int lvreg_depth_set_cloud(lvreg_handle* h, const lvreg_cloud* depth_cloud);
/* which = 0: depthCloud; 1: depth_cloud_local after the range-image filter of the last lvreg_get_depth */
int lvreg_depth_get_cloud(lvreg_handle* h, int which, lvreg_cloud_out* out, size_t* n);
int lvreg_get_depth(lvreg_handle* h, const float T_inv[12], const float* features_xyz, size_t n, int num_bins,
                    float* depth_out, float* features_3d_out);

/* ---- "next" row: deskew + range-image projection -------------------------------------
 * Replaces the per-point work of lidar_odometry/src/imageProjection.cpp: projectPointCloud (571-623)
 * with deskewPoint (538-569) / findRotation (495-526), and cloudExtraction (625-647).  The IMU
 * integration that fills imuTime / imuRot{X,Y,Z} (340-408) and the odometry lookup stay on the host
 * and are inputs.  The result -- extractedCloud plus the CloudInfo side channels start_ring_index,
 * end_ring_index, point_col_ind, point_range (msg/CloudInfo.msg) -- stays on the device and feeds
 * lvreg_extract_features directly (lvreg_get_projection), so that a raw scan goes to a registered
 * pose with one upload. */
typedef struct lvreg_raw_cloud {       /* laserCloudIn, any AoS layout (PointXYZIRT) */
    const void* data;
    size_t      n;
    uint32_t    stride;                /* bytes per point; x, y, z are floats at offsets 0, 4, 8 */
    uint32_t    intensity_offset;      /* float */
    uint32_t    ring_offset;           /* uint16_t */
    uint32_t    time_offset;           /* float, seconds relative to the scan start */
    int32_t     on_device;
    int32_t     reserved;
} lvreg_raw_cloud;
typedef struct lvreg_projection_params {
    int32_t n_scan, horizon_scan, downsample_rate;   /* N_SCAN (<= 255), Horizon_SCAN, downsampleRate */
    int32_t sensor;                    /* SensorType: 0 VELODYNE, 1 OUSTER, 2 LIVOX */
    float   lidar_min_range, lidar_max_range;
    int32_t deskew;                    /* deskewFlag != -1 && cloudInfo.imu_available */
    int32_t imu_pointer_cur;           /* index of the last valid IMU sample (imuPointerCur after imuDeskewInfo) */
    double  time_scan_cur;             /* timeScanCur */
    const double* imu_time;            /* host arrays, imu_pointer_cur + 1 entries */
    const double* imu_rot_x;
    const double* imu_rot_y;
    const double* imu_rot_z;
} lvreg_projection_params;
int lvreg_project_cloud(lvreg_handle* h, const lvreg_raw_cloud* in, const lvreg_projection_params* prm,
                        size_t* n_extracted);
/* device-resident extractedCloud + side channels of the last projection, in the form
 * lvreg_extract_features takes (ring indices: host arrays owned by the handle, valid until the next
 * lvreg_project_cloud; the per-point arrays: device pointers) */
int lvreg_get_projection(const lvreg_handle* h, lvreg_cloud* extracted, lvreg_scan_info* info);
/* host copies for the CloudInfo message; every output is optional */
int lvreg_download_projection(lvreg_handle* h, lvreg_cloud_out* extracted, float* point_range, int32_t* point_col_ind,
                              int32_t* start_ring_index, int32_t* end_ring_index, size_t* n);

/* ---- measurement -------------------------------------------------------------------------- */
int lvreg_get_timings(const lvreg_handle* h, lvreg_timings* t);
/* Phase profile of the last lvreg_scan2map / lvreg_register_scan launch, from block 0's
 * globaltimer stamps inside the cooperative kernel: for each executed iteration
 * us[iter][0..3] = {tile work (kNN + fits + reduction), wait at the grid barrier, grid reduction,
 * 6x6 solve + pose update} in microseconds.  `us` has room for LVREG_MAX_ITERS x 4 floats. */
int lvreg_get_iteration_profile(const lvreg_handle* h, float* us, int* iterations);
/* Diagnostics (set LVREG_DEBUG_TILES=1 before lvreg_create): duration in ns of every query tile of
 * iteration 1 of the last registration, in tile order (corner tiles first). */
int lvreg_debug_tile_times(lvreg_handle* h, uint32_t* ns_out, size_t cap, size_t* n_tiles);
/* Diagnostics (LVREG_DEBUG_TILES=1): how the shared-memory search of the last registration staged its
 * query tiles in iteration 0: out[0..2] = tiles staged whole / as halves / as quarters, out[3] = queries that
 * took the global-memory search, out[4] = copy-barrier time-outs (must be 0), out[5] = how iteration 0 of the
 * last registration decided degeneracy (1 = Cholesky shortcut, 2 = 6x6 Jacobi; any kernel variant),
 * out[6] = local-map VoxelGrid jobs that had to be redone by the device-wide sort because a bucket of the
 * sample-sort path overflowed, out[7] = jobs completed by the sample-sort path (since the handle was created). */
int lvreg_debug_stage_stats(lvreg_handle* h, uint32_t out[8]);
/* Measurement support (bench.py `roofline`): with timing enabled every call brackets the bucket kernel of each local-map
 * VoxelGrid (extractCloud, MO:959-965; voxelgrid_bucket.cuh) with a CUDA event pair on the stream it is launched on.
 * lvreg_get_bucket_kernel_ms returns, for the last call, [0] corner map / [1] surf map: the kernel's duration in ms
 * (0 if the map did not go through that kernel), its input points and its output voxels. */
int lvreg_enable_kernel_timing(lvreg_handle* h, int on);
int lvreg_get_bucket_kernel_ms(lvreg_handle* h, float ms[2], uint32_t n_in[2], uint32_t n_out[2]);
/* total kernels launched by this handle since creation */
int lvreg_get_launch_count(const lvreg_handle* h, uint64_t* n);
/* kNN micro-benchmark on device-resident data: runs `repeats` launches of the chosen variant on
 * the current map with the given device queries and returns the mean device time per launch. */
int lvreg_bench_knn5(lvreg_handle* h, int which_map, const lvreg_cloud* queries, int variant,
                     int repeats, float* ms_per_launch);

/* Radix-sort micro-benchmark: sorts n pseudo-random (key, index) pairs with `key_bits` significant
 * bits `repeats` times on the handle's stream; returns the mean device time of one whole sort and the
 * number of 8-bit passes it took.  Algorithmic traffic per pass = 16 B per pair. */
/* C4 "with fused residual": times the fused search + fit + residual kernel of one feature class on
 * device-resident queries (CUDA events, ms per launch); nothing is copied back */
int lvreg_bench_residuals(lvreg_handle* h, int which, const lvreg_cloud* queries, const float pose_rpyxyz[6],
                          int repeats, float* ms);
int lvreg_bench_sort(lvreg_handle* h, size_t n, int key_bits, int repeats, float* ms_per_sort, int* passes);

#ifdef __cplusplus
}
#endif
#endif
