/*
 * oracle.h -- CPU restatement of the reference's scan-to-map registration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lidar_visual_inertial_slam_b200/ may
 * include, link or load this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker and the
 * timed CPU baseline.
 *
 * What it restates (MO: = /root/reference/lidar_odometry/src/mapOptimization.cpp):
 *   MO:339-385   pointAssociateToMap / transformPointCloud
 *   MO:894-970   extractNearby / extractCloud (local-map build)
 *   MO:987-999   downsampleCurrentScan
 *   MO:1006-1343 cornerOptimization, surfOptimization, combineOptimizationCoeffs,
 *                LMOptimization, scan2MapOptimization
 *   MO:1345-1385 transformUpdate / constraintTransformation
 *
 * The arithmetic of that path lives in third-party libraries that are NOT vendored
 * under /root/reference and are not installed here (PCL 1.12 VoxelGrid/KdTreeFLANN,
 * FLANN 1.9 KDTreeSingleIndex, OpenCV 4 cv::eigen / cv::solve / gemm, Eigen 3.4
 * ColPivHouseholderQR; versions unpinned by the reference).
 * Their published algorithms are restated in oracle_math.cpp / oracle_cloud.cpp /
 * oracle_kdtree.cpp.
 *
 * PARITY PIN STATUS: the reference ships no tests, golden vectors or fixtures for
 * this path, and cannot be compiled here (needs ROS 2 + PCL + OpenCV C++ + GTSAM),
 * so end-to-end parity is UNPINNED by the reference.  What IS pinned, bit-exactly,
 * against the third-party code the reference calls (tests/test_oracle_pins.py):
 *   - orc_jacobi_eigen   == cv2.eigen            (OpenCV 4.13, Jacobi back end)
 *   - orc_qr_solve       == cv2.solve(DECOMP_QR)
 *   - orc_lu_solve       == cv2.solve(DECOMP_LU) (what `matV.inv()*matV2` lowers to)
 *   - orc_gemm_*         ~= cv2.gemm
 * and, to tolerance, orc_plane_fit vs numpy.linalg.lstsq, kNN vs scipy cKDTree and
 * numpy brute force, voxel keys vs numpy integer arithmetic.
 *
 * All clouds here are packed float4 rows {x, y, z, intensity}.  Poses are
 * float[6] = {roll, pitch, yaw, x, y, z} (transformTobeMapped order, MO:126).
 */
#ifndef LVREG_ORACLE_H
#define LVREG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
    float corner_leaf;        /* mappingCornerLeafSize  utility.h:266 (0.2)  */
    float surf_leaf;          /* mappingSurfLeafSize    utility.h:268 (0.4)  */
    int   edge_min_valid;     /* edgeFeatureMinValidNum utility.h:259 (10)   */
    int   surf_min_valid;     /* surfFeatureMinValidNum utility.h:261 (100)  */
    int   max_iters;          /* MO:1325 (20) */
    float knn_gate_sq;        /* MO:1025,1121 (1.0) */
    float line_eig_ratio;     /* MO:1052 (3) */
    float plane_tol;          /* MO:1142 (0.2) */
    float min_weight;         /* MO:1088,1159 (0.1) */
    int   min_matches;        /* MO:1210 (50) */
    float degeneracy_eig;     /* MO:1272 (100) */
    float conv_deg;           /* MO:1309 (0.05) */
    float conv_cm;            /* MO:1309 (0.05) */
    int   reference_quirks;   /* 1: reproduce the shadowed-matP behaviour */
    float keyframe_search_radius;  /* surroundingKeyframeSearchRadius utility.h:285 (50) */
    float keyframe_density;        /* surroundingKeyframeDensity      utility.h:287 (2.0) */
    float rotation_tolerance;      /* rotation_tollerance utility.h:273 (1000) */
    float z_tolerance;             /* z_tollerance        utility.h:271 (1000) */
    int   num_threads;             /* numberOfCores (OpenMP team size) */
} orc_params;

typedef struct orc_result {
    int   status;             /* 0 ok, 1 not enough features, 2 no map */
    int   iterations;         /* LM iterations executed (1..max_iters) */
    int   converged;
    int   degenerate;         /* isDegenerate after the call */
    int   n_sel[32];          /* laserCloudSelNum per iteration */
    float pose_iter[32][6];   /* pose after each iteration */
} orc_result;

void orc_default_params(orc_params* p);

/* ---- small dense math (restated OpenCV / Eigen), oracle_math.cpp ------------ */
/* cv::eigen for symmetric CV_32F n x n (Jacobi).  evals descending, evecs as rows. */
void orc_jacobi_eigen(const float* A, int n, float* evals, float* evecs);
/* cv::solve(A, b, x, DECOMP_QR) for square CV_32F, nrhs columns.  returns 0 if singular (x zeroed). */
int  orc_qr_solve(const float* A, const float* b, int n, int nrhs, float* x);
/* cv::solve(A, B, X, DECOMP_LU) for square CV_32F.  returns 0 if singular (X zeroed). */
int  orc_lu_solve(const float* A, const float* B, int n, int nrhs, float* X);
/* cv::gemm for CV_32F: C(m x n) = A(m x k) * B(k x n), double accumulators, row order. */
void orc_gemm(const float* A, const float* B, int m, int k, int n, float* C);
/* A^T A (6x6) and A^T b (6) for an N x 6 row matrix, as matAt*matA / matAt*matB MO:1257-1259 */
void orc_normal_equations(const float* A, const float* b, int nrows, float* AtA, float* Atb);
/* Eigen::Matrix<float,5,3>::colPivHouseholderQr().solve(b), MO:1128 */
void orc_colpiv_qr_solve_5x3(const float* A_rowmajor_5x3, const float* b5, float* x3);

/* ---- clouds, oracle_cloud.cpp ---------------------------------------------- */
/* pcl::getTransformation(x,y,z,roll,pitch,yaw) -> row-major 3x4.  MO:399-407 */
void orc_pose_to_affine(const float pose_rpyxyz[6], float T[12]);
/* transformPointCloud MO:347-385 */
void orc_transform_cloud(const float* in, size_t n, const float T[12], float* out, int num_threads);
/* pcl::VoxelGrid::filter with a STABLE sort.  out has room for n rows.
 * keys_out (optional, n entries) receives the per-input-point voxel idx; returns the
 * number of output points, or n with *passthrough=1 when the leaf-size overflow rule fires. */
size_t orc_voxelgrid(const float* in, size_t n, float leaf, float* out, uint32_t* keys_out,
                     uint32_t* out_keys /* optional, per output voxel */, int* passthrough);

/* ---- exact 5-NN, oracle_kdtree.cpp ----------------------------------------- */
/* brute force, (d2, index) lexicographic order; idx -1 / d2 +inf when the map has < 5 points */
void orc_knn5_brute(const float* map, size_t m, const float* queries, size_t nq,
                    int32_t* idx, float* d2, int num_threads);
typedef struct orc_kdtree orc_kdtree;
orc_kdtree* orc_kdtree_build(const float* map, size_t m);       /* FLANN-like single tree, leaf 15 */
void orc_kdtree_free(orc_kdtree* t);
void orc_kdtree_knn(const orc_kdtree* t, const float* queries, size_t nq, int k,
                    int32_t* idx, float* d2, int num_threads);
/* radiusSearch (sorted by (d2, index)); returns count, writes up to cap entries */
size_t orc_kdtree_radius(const orc_kdtree* t, const float q[3], float radius,
                         int32_t* idx, float* d2, size_t cap);

/* ---- registration, oracle_reg.cpp ------------------------------------------ */
/* cornerOptimization MO:1006-1096.  coeff (n x 4), flag (n), knn_idx optional (n x 5) */
void orc_corner_residuals(const float* map, size_t m, const orc_kdtree* tree,
                          const float* pts, size_t n, const float pose[6], const orc_params* p,
                          float* coeff, uint8_t* flag, int32_t* knn_idx);
/* surfOptimization MO:1098-1167 */
void orc_surf_residuals(const float* map, size_t m, const orc_kdtree* tree,
                        const float* pts, size_t n, const float pose[6], const orc_params* p,
                        float* coeff, uint8_t* flag, int32_t* knn_idx);
/* Jacobian rows of LMOptimization MO:1222-1255: A (n x 6), b (n) */
void orc_jacobian_rows(const float* ori, const float* coeff, size_t n, const float pose[6],
                       float* A, float* b);
/* persistent LM state across calls (isDegenerate, matP) */
typedef struct orc_lm_state { int is_degenerate; float matP[36]; } orc_lm_state;
/* LMOptimization MO:1190-1313.  returns 1 when converged.  pose updated in place. */
int orc_lm_step(const float* ori, const float* coeff, size_t n_sel, int iter, float pose[6],
                orc_lm_state* st, const orc_params* p, float AtA_out[36], float Atb_out[6],
                float x_out[6]);
/* scan2MapOptimization MO:1315-1343 on explicit DS maps / DS scan clouds */
void orc_scan2map(const float* corner_map, size_t mc, const float* surf_map, size_t ms,
                  const float* corner, size_t nc, const float* surf, size_t ns,
                  float pose[6], orc_lm_state* st, const orc_params* p, orc_result* res);
/* transformUpdate MO:1345-1375 */
void orc_transform_update(float pose[6], int imu_available, float imu_roll, float imu_pitch,
                          float imu_weight, const orc_params* p);

/* ---- mapOptimization-like object (keyframes + local map), oracle_reg.cpp ---- */
typedef struct orc_mo orc_mo;
orc_mo* orc_mo_create(const orc_params* p);
void    orc_mo_destroy(orc_mo* mo);
/* saveKeyFramesAndFactor without iSAM2: store clouds + pose {roll,pitch,yaw,x,y,z} + time */
int     orc_mo_add_keyframe(orc_mo* mo, const float* corner, size_t nc, const float* surf, size_t ns,
                            const float pose[6], double time);
size_t  orc_mo_num_keyframes(const orc_mo* mo);
/* extractNearby MO:894-929: writes the keyframe id list (in concatenation order) */
size_t  orc_mo_extract_nearby(orc_mo* mo, double time_now, int32_t* ids, size_t cap);
/* extractCloud MO:931-970 for an explicit id list; builds both DS maps and kd-trees */
void    orc_mo_build_local_map(orc_mo* mo, const int32_t* ids, size_t n);
size_t  orc_mo_map_size(const orc_mo* mo, int which /*0 corner, 1 surf*/);
void    orc_mo_get_map(const orc_mo* mo, int which, float* out);
/* downsampleCurrentScan + scan2MapOptimization on the current local map (kd-trees are
 * rebuilt inside, as the reference does at MO:1322-1323) */
void    orc_mo_register_scan(orc_mo* mo, const float* corner_raw, size_t nc_raw,
                             const float* surf_raw, size_t ns_raw, float pose[6], orc_result* res,
                             size_t* nc_ds, size_t* ns_ds);

/* ---- "next" row: FeatureExtraction, oracle_feature.cpp --------------------------
 * calculateSmoothness + markOccludedPoints + extractFeatures of
 * lidar_odometry/src/featureExtraction.cpp:87-245 on one deskewed, ring-ordered cloud
 * (imageProjection.cpp:624-647 layout: start/end_ring_index, point_col_ind, point_range).
 * Pinned choices for what the reference leaves unspecified: std::sort ties -> by index;
 * cloudNeighborPicked / cloudLabel / cloudSmoothness entries the reference never initialises
 * (i < 5, i >= n-5) read as 0 / {0, i}.  corner_out / surf_out have room for n rows.
 * label_out (optional, n): cloudLabel.  Returns 0. */
int orc_extract_features(const float* pts, size_t n, const float* point_range, const int32_t* point_col_ind,
                         const int32_t* start_ring_index, const int32_t* end_ring_index, int n_scan,
                         float edge_threshold, float surf_threshold, float surf_leaf,
                         float* corner_out, size_t* n_corner, float* surf_out, size_t* n_surf,
                         int32_t* label_out);

/* ---- "next" row: loop-closure ICP, oracle_icp.cpp + oracle_reg.cpp ---------------
 * pcl::IterativeClosestPoint as configured at MO:578-590, getFitnessScore MO:592, submaps MO:719-741,
 * candidate search MO:630-661, pose correction MO:600-609.  Header of oracle_icp.cpp lists what is
 * restated from PCL and the one deliberate deviation (double raw-moment Umeyama). */
enum { ORC_ICP_NOT_CONVERGED = 0, ORC_ICP_ITERATIONS = 1, ORC_ICP_TRANSFORM = 2, ORC_ICP_ABS_MSE = 3,
       ORC_ICP_REL_MSE = 4, ORC_ICP_NO_CORRESPONDENCES = 5, ORC_ICP_NO_INPUT = 6 };
typedef struct orc_icp_params {
    float max_corr_dist;
    int max_iterations;
    double transformation_epsilon;
    double euclidean_fitness_epsilon;
    int num_threads;
} orc_icp_params;
typedef struct orc_icp_result {
    int converged, iterations, state, n_correspondences;
    double fitness, mse;
    float final_transformation[16];       /* row-major 4x4 */
} orc_icp_result;
typedef struct orc_loop_result {
    int status;                           /* 0 constraint produced, 1 submap too small (MO:572), 2 ICP not converged,
                                             3 fitness above the gate (MO:592) */
    int n_source, n_target;
    orc_icp_result icp;
    float pose_from[6];                   /* corrected pose of key_cur {r,p,y,x,y,z}, MO:604-609 */
    float pose_to[6];                     /* stored pose of key_pre, MO:611 */
    float noise;                          /* (float)getFitnessScore, MO:613 */
} orc_loop_result;
void orc_icp_default_params(orc_icp_params* p);
void orc_umeyama_from_moments(const double* mom17, float* T4x4);
void orc_nn1(const float* tgt, size_t nt, const float* q, size_t nq, int32_t* idx, float* d2, int num_threads);
void orc_icp_align(const float* src, size_t ns, const float* tgt, size_t nt, const orc_icp_params* P,
                   orc_icp_result* res);
void orc_correct_pose(const float* correction4x4, const float pose[6], float out[6]);
/* loopFindNearKeyframes into slot 0 (source) or 1 (target); returns the row count */
size_t orc_mo_loop_find_near_keyframes(orc_mo* mo, int key, int search_num, int slot);
void   orc_mo_get_loop_cloud(const orc_mo* mo, int slot, float* out);
/* publishGlobalMap MO:493-508 / saveMapService MO:199-231 for an explicit id list; result in slot 0 */
size_t orc_mo_build_global_map(orc_mo* mo, const int32_t* ids, size_t n_ids, int which, float leaf);
/* detectLoopClosureDistance without the loopIndexContainer bookkeeping; returns 1 when a pair was found */
int    orc_mo_detect_loop_closure_distance(orc_mo* mo, double time_cur, float radius, float time_diff,
                                           int* key_cur, int* key_pre);
void   orc_mo_perform_loop_closure(orc_mo* mo, int key_cur, int key_pre, int search_num,
                                   const orc_icp_params* P, float fitness_gate, orc_loop_result* out);

/* ---- "next" row: LiDAR depth for visual features, oracle_depth.cpp ----------------
 * feature_tracker_node.cpp:273-375 (lidar_callback: stack + 0.2 m VoxelGrid) and
 * feature_tracker.h:150-283 (DepthRegister::get_depth from the camera-frame transform on). */
typedef struct orc_depth orc_depth;
orc_depth* orc_depth_create(void);
void   orc_depth_destroy(orc_depth* d);
size_t orc_depth_add_cloud(orc_depth* d, const float* cloud, size_t n, const float T_now[12], double stamp);
size_t orc_depth_cloud_size(const orc_depth* d);
void   orc_depth_get_cloud(const orc_depth* d, float* out);
size_t orc_get_depth(const float* depth_cloud, size_t m, const float Tinv[12], const float* feat_xyz,
                     size_t n, int num_bins, float* depth_out, float* feat3d_out, float* local_out);

/* ---- "next" row: deskew + range-image projection, oracle_projection.cpp -----------
 * imageProjection.cpp:495-647 (findRotation, deskewPoint, projectPointCloud, cloudExtraction). */
typedef struct orc_projection_params {
    int n_scan, horizon_scan, downsample_rate;
    int sensor;                      /* 0 velodyne, 1 ouster, 2 livox (utility.h SensorType) */
    float lidar_min_range, lidar_max_range;
    int deskew;                      /* deskewFlag != -1 && cloudInfo.imu_available */
    int imu_pointer_cur;             /* index of the last valid IMU sample */
    double time_scan_cur;
    const double* imu_time;
    const double* imu_rot_x;
    const double* imu_rot_y;
    const double* imu_rot_z;
} orc_projection_params;
void orc_find_rotation(double point_time, const double* imu_time, const double* rx, const double* ry,
                       const double* rz, int imu_pointer_cur, float rot[3]);
size_t orc_project_cloud(const float* pts, const uint16_t* ring, const float* rel_time, size_t n,
                         const orc_projection_params* P, float* extracted, float* point_range,
                         int32_t* point_col_ind, int32_t* start_ring_index, int32_t* end_ring_index);

#ifdef __cplusplus
}
#endif
#endif
