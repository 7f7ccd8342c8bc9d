"""lidar_visual_inertial_slam_b200 -- B200-native scan-to-map registration (liblvreg).

The product is the C-ABI shared library `liblvreg.so` (include/lvreg.h).  This module is the thin
ctypes binding used by the tests, bench.py and Python callers; it loads the in-tree library and
fails loudly if it is missing (there is no CPU fallback and nothing here touches oracle/).
"""
from .binding import (  # noqa: F401
    Lvreg, LvregError, Params, Result, MapInfo, Timings, lib, lib_path, default_params,
    CORNER, SURF, KNN_GRID_GATED, KNN_GRID_EXACT, KNN_BRUTE, KNN_GRID_STAGED,
    OK, ERR_INVALID, ERR_CUDA, ERR_NOT_ENOUGH_FEATURES, ERR_NO_KEYFRAMES, ERR_NO_MAP, ERR_CAPACITY,
    pose_to_affine, host_alloc_f32,
    RawCloud, ProjectionParams, make_raw_cloud, SENSOR_VELODYNE, SENSOR_OUSTER, SENSOR_LIVOX, LAYOUT_VELODYNE, LAYOUT_LIVOX,
    IcpParams, IcpResult, LoopResult, icp_default_params, correct_pose,
    ICP_NOT_CONVERGED, ICP_ITERATIONS, ICP_TRANSFORM, ICP_ABS_MSE, ICP_REL_MSE, ICP_NO_CORRESPONDENCES, ICP_NO_INPUT,
    LOOP_OK, LOOP_SUBMAP_TOO_SMALL, LOOP_NOT_CONVERGED, LOOP_FITNESS_TOO_HIGH,
)
