"""ctypes binding of include/lvreg.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "liblvreg.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOT_ENOUGH_FEATURES, ERR_NO_KEYFRAMES, ERR_NO_MAP, ERR_CAPACITY = range(7)
CORNER, SURF = 0, 1
KNN_GRID_GATED, KNN_GRID_EXACT, KNN_BRUTE, KNN_GRID_STAGED = 0, 1, 2, 3
MAX_ITERS = 32


class LvregError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("lvreg status %d: %s" % (status, msg))
        self.status = status


class Cloud(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("stride", C.c_uint32),
                ("intensity_offset", C.c_uint32), ("on_device", C.c_int32), ("reserved", C.c_int32)]


class CloudOut(C.Structure):
    _fields_ = [("data", C.c_void_p), ("capacity", C.c_size_t), ("stride", C.c_uint32),
                ("intensity_offset", C.c_uint32), ("on_device", C.c_int32), ("reserved", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("corner_leaf", C.c_float), ("surf_leaf", C.c_float),
                ("edge_min_valid", C.c_int32), ("surf_min_valid", C.c_int32),
                ("max_iters", C.c_int32), ("knn_gate_sq", C.c_float), ("line_eig_ratio", C.c_float),
                ("plane_tol", C.c_float), ("min_weight", C.c_float), ("min_matches", C.c_int32),
                ("degeneracy_eig", C.c_float), ("conv_deg", C.c_float), ("conv_cm", C.c_float),
                ("reference_quirks", C.c_int32), ("rotation_tolerance", C.c_float),
                ("z_tolerance", C.c_float), ("imu_rpy_weight", C.c_float), ("reserved", C.c_int32 * 7)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("degenerate", C.c_int32),
                ("n_corner_ds", C.c_int32), ("n_surf_ds", C.c_int32),
                ("n_corner_map", C.c_int32), ("n_surf_map", C.c_int32),
                ("n_sel", C.c_int32 * MAX_ITERS), ("pose_iter", (C.c_float * 6) * MAX_ITERS),
                ("cost", C.c_float * MAX_ITERS)]


class MapInfo(C.Structure):
    _fields_ = [("n_corner_in", C.c_uint64), ("n_surf_in", C.c_uint64),
                ("n_corner_ds", C.c_uint64), ("n_surf_ds", C.c_uint64),
                ("grid_dims", (C.c_int32 * 3) * 2), ("grid_cell", C.c_float * 2)]


class Timings(C.Structure):
    _fields_ = [("upload_ms", C.c_float), ("downsample_ms", C.c_float), ("map_build_ms", C.c_float),
                ("grid_build_ms", C.c_float), ("register_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("reserved", C.c_int32)]


class IcpParams(C.Structure):
    _fields_ = [("max_corr_dist", C.c_float), ("max_iterations", C.c_int32),
                ("transformation_epsilon", C.c_double), ("euclidean_fitness_epsilon", C.c_double),
                ("reserved", C.c_float * 4)]


class IcpResult(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int32), ("state", C.c_int32),
                ("n_correspondences", C.c_int32), ("fitness", C.c_double), ("mse", C.c_double),
                ("final_transformation", C.c_float * 16)]

    @property
    def T(self):
        return np.array(self.final_transformation[:], np.float32).reshape(4, 4)


class LoopResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_source", C.c_int32), ("n_target", C.c_int32), ("reserved", C.c_int32),
                ("icp", IcpResult), ("pose_from", C.c_float * 6), ("pose_to", C.c_float * 6),
                ("noise", C.c_float), ("reserved2", C.c_float)]


class RawCloud(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("stride", C.c_uint32), ("intensity_offset", C.c_uint32),
                ("ring_offset", C.c_uint32), ("time_offset", C.c_uint32), ("on_device", C.c_int32), ("reserved", C.c_int32)]


class ProjectionParams(C.Structure):
    _fields_ = [("n_scan", C.c_int32), ("horizon_scan", C.c_int32), ("downsample_rate", C.c_int32), ("sensor", C.c_int32),
                ("lidar_min_range", C.c_float), ("lidar_max_range", C.c_float), ("deskew", C.c_int32),
                ("imu_pointer_cur", C.c_int32), ("time_scan_cur", C.c_double), ("imu_time", C.c_void_p),
                ("imu_rot_x", C.c_void_p), ("imu_rot_y", C.c_void_p), ("imu_rot_z", C.c_void_p)]


SENSOR_VELODYNE, SENSOR_OUSTER, SENSOR_LIVOX = 0, 1, 2
# (intensity, ring, time) byte offsets of the reference's 32-byte point structs (imageProjection.cpp:4-29)
LAYOUT_VELODYNE = (16, 20, 24)
LAYOUT_LIVOX = (16, 24, 20)


def make_raw_cloud(xyzi, ring, rel_time, layout=LAYOUT_LIVOX):
    """pack [n,4] float32 + ring (uint16) + time (float32) into the reference's 32-byte AoS point"""
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    n = len(xyzi)
    buf = np.zeros((n, 32), np.uint8)
    f = buf.view(np.float32).reshape(n, 8)
    f[:, 0:3] = xyzi[:, 0:3]
    f[:, 3] = 1.0
    f[:, layout[0] // 4] = xyzi[:, 3]
    f[:, layout[2] // 4] = np.asarray(rel_time, np.float32)
    buf.view(np.uint16).reshape(n, 16)[:, layout[1] // 2] = np.asarray(ring, np.uint16)
    return buf


ICP_NOT_CONVERGED, ICP_ITERATIONS, ICP_TRANSFORM, ICP_ABS_MSE, ICP_REL_MSE, ICP_NO_CORRESPONDENCES, ICP_NO_INPUT = range(7)
LOOP_OK, LOOP_SUBMAP_TOO_SMALL, LOOP_NOT_CONVERGED, LOOP_FITNESS_TOO_HIGH = range(4)

_lib = None


def lib_path():
    return _LIBPATH


def lib():
    """Load liblvreg.so (in-tree).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise ImportError("liblvreg.so is missing (%s): run __graft_entry__.build() / "
                              "python -m lidar_visual_inertial_slam_b200.build" % _LIBPATH)
        L = C.CDLL(_LIBPATH)
        L.lvreg_last_error.restype = C.c_char_p
        L.lvreg_status_string.restype = C.c_char_p
        L.lvreg_host_alloc.restype = C.c_void_p
        L.lvreg_host_alloc.argtypes = [C.c_size_t]
        L.lvreg_host_free.argtypes = [C.c_void_p]
        L.lvreg_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.lvreg_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().lvreg_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def icp_default_params(**kw):
    p = IcpParams()
    lib().lvreg_icp_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def correct_pose(correction4x4, pose):
    T = np.ascontiguousarray(correction4x4, np.float32).reshape(16)
    pose = np.ascontiguousarray(pose, np.float32)
    out = np.zeros(6, np.float32)
    lib().lvreg_correct_pose(T.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p),
                             out.ctypes.data_as(C.c_void_p))
    return out


def pose_to_affine(pose):
    pose = np.ascontiguousarray(pose, np.float32)
    T = np.zeros(12, np.float32)
    lib().lvreg_pose_to_affine(pose.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p))
    return T


def host_alloc_f32(shape):
    """float32 array in page-locked memory (lvreg_host_alloc); keep the returned array alive."""
    n = int(np.prod(shape))
    ptr = lib().lvreg_host_alloc(n * 4)
    if not ptr:
        raise MemoryError("lvreg_host_alloc failed")
    buf = (C.c_float * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float32).reshape(shape)
    return arr


def _cloud(a):
    """numpy [n,4] (packed x,y,z,i) or [n,8] (pcl::PointXYZI layout) float32 -> lvreg_cloud"""
    if isinstance(a, Cloud):
        return a, None
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("cloud must be [n,4] or [n,8] float32")
    c = Cloud()
    c.data = a.ctypes.data if len(a) else None
    c.n = len(a)
    c.stride = a.shape[1] * 4
    c.intensity_offset = 12 if a.shape[1] == 4 else 16
    c.on_device = 0
    return c, a


def device_cloud(ptr, n, stride=16, intensity_offset=12):
    c = Cloud()
    c.data = ptr
    c.n = n
    c.stride = stride
    c.intensity_offset = intensity_offset
    c.on_device = 1
    return c


def to_pcl_layout(a):
    """[n,4] packed -> [n,8] pcl::PointXYZI rows {x,y,z,1 | intensity,0,0,0}"""
    a = np.asarray(a, np.float32)
    out = np.zeros((len(a), 8), np.float32)
    out[:, :3] = a[:, :3]
    out[:, 3] = 1.0
    out[:, 4] = a[:, 3]
    return out


class ScanInfo(C.Structure):
    _fields_ = [("start_ring_index", C.c_void_p), ("end_ring_index", C.c_void_p), ("n_scan", C.c_int32),
                ("reserved", C.c_int32), ("point_col_ind", C.c_void_p), ("point_range", C.c_void_p)]


class Lvreg:
    """One registration handle = one mapOptimization instance on one GPU / stream."""

    def __init__(self, params=None, device=0, stream=None):
        self.L = lib()
        self.params = params or default_params()
        h = C.c_void_p()
        st = self.L.lvreg_create(C.byref(self.params), int(device), C.c_void_p(stream or 0), C.byref(h))
        if st != OK:
            raise LvregError(st, "lvreg_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.lvreg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st, soft=()):
        if st != OK and st not in soft:
            raise LvregError(st, self.L.lvreg_last_error(self.h).decode())
        return st

    def reserve(self, map_points_corner=0, map_points_surf=0, scan_points_corner=0, scan_points_surf=0, max_grid_cells=0):
        self._ck(self.L.lvreg_reserve(self.h, C.c_size_t(map_points_corner), C.c_size_t(map_points_surf),
                                      C.c_size_t(scan_points_corner), C.c_size_t(scan_points_surf),
                                      C.c_size_t(max_grid_cells)))

    # ---- keyframes / map ----
    def add_keyframe(self, corner, surf, pose):
        c, _k1 = _cloud(corner)
        s, _k2 = _cloud(surf)
        pose = np.ascontiguousarray(pose, np.float32)
        kid = C.c_int32(-1)
        self._ck(self.L.lvreg_add_keyframe(self.h, C.byref(c), C.byref(s), pose.ctypes.data_as(C.c_void_p), C.byref(kid)))
        return kid.value

    def update_keyframe_poses(self, poses):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 6)
        self._ck(self.L.lvreg_update_keyframe_poses(self.h, poses.ctypes.data_as(C.c_void_p), C.c_size_t(len(poses))))

    def num_keyframes(self):
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_num_keyframes(self.h, C.byref(n)))
        return n.value

    def clear_keyframes(self):
        self._ck(self.L.lvreg_clear_keyframes(self.h))

    def build_local_map(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        info = MapInfo()
        self._ck(self.L.lvreg_build_local_map(self.h, ids.ctypes.data_as(C.c_void_p), C.c_size_t(len(ids)), C.byref(info)))
        return info

    def set_local_map(self, corner_ds, surf_ds):
        c, _k1 = _cloud(corner_ds)
        s, _k2 = _cloud(surf_ds)
        info = MapInfo()
        self._ck(self.L.lvreg_set_local_map(self.h, C.byref(c), C.byref(s), C.byref(info)))
        return info

    def _get_cloud(self, fn, which, pcl_layout=False):
        n = C.c_size_t(0)
        self._ck(fn(self.h, which, None, C.byref(n)))
        cols = 8 if pcl_layout else 4
        out = np.zeros((n.value, cols), np.float32)
        co = CloudOut()
        co.data = out.ctypes.data if n.value else None
        co.capacity = n.value
        co.stride = cols * 4
        co.intensity_offset = 16 if pcl_layout else 12
        self._ck(fn(self.h, which, C.byref(co), C.byref(n)))
        return out

    def get_local_map(self, which, pcl_layout=False):
        return self._get_cloud(self.L.lvreg_get_local_map, which, pcl_layout)

    # ---- scan ----
    def downsample_scan(self, corner_raw, surf_raw):
        c, _k1 = _cloud(corner_raw)
        s, _k2 = _cloud(surf_raw)
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        self._ck(self.L.lvreg_downsample_scan(self.h, C.byref(c), C.byref(s), C.byref(nc), C.byref(ns)))
        return nc.value, ns.value

    def set_scan_ds(self, corner_ds, surf_ds):
        c, _k1 = _cloud(corner_ds)
        s, _k2 = _cloud(surf_ds)
        self._ck(self.L.lvreg_set_scan_ds(self.h, C.byref(c), C.byref(s)))

    def get_scan_ds(self, which, pcl_layout=False):
        return self._get_cloud(self.L.lvreg_get_scan_ds, which, pcl_layout)

    # ---- registration ----
    def scan2map(self, pose):
        pose = np.ascontiguousarray(pose, np.float32).copy()
        res = Result()
        st = self._ck(self.L.lvreg_scan2map(self.h, pose.ctypes.data_as(C.c_void_p), C.byref(res)),
                      soft=(ERR_NOT_ENOUGH_FEATURES, ERR_NO_KEYFRAMES, ERR_NO_MAP))
        return pose, res, st

    def register_scan(self, corner_raw, surf_raw, ids, pose):
        c, _k1 = _cloud(corner_raw)
        s, _k2 = _cloud(surf_raw)
        pose = np.ascontiguousarray(pose, np.float32).copy()
        res = Result()
        if ids is None:
            idp, nid = None, 0
        else:
            ids = np.ascontiguousarray(ids, np.int32)
            idp, nid = ids.ctypes.data_as(C.c_void_p), len(ids)
        st = self._ck(self.L.lvreg_register_scan(self.h, C.byref(c), C.byref(s), idp, C.c_size_t(nid),
                                                 pose.ctypes.data_as(C.c_void_p), C.byref(res)),
                      soft=(ERR_NOT_ENOUGH_FEATURES, ERR_NO_KEYFRAMES, ERR_NO_MAP))
        return pose, res, st

    def transform_update(self, pose, imu_available=False, imu_roll=0.0, imu_pitch=0.0):
        pose = np.ascontiguousarray(pose, np.float32).copy()
        self._ck(self.L.lvreg_transform_update(self.h, pose.ctypes.data_as(C.c_void_p), int(imu_available),
                                               C.c_float(imu_roll), C.c_float(imu_pitch)))
        return pose

    def set_imu_prior(self, imu_available, imu_roll=0.0, imu_pitch=0.0):
        self._ck(self.L.lvreg_set_imu_prior(self.h, int(bool(imu_available)), C.c_float(imu_roll), C.c_float(imu_pitch)))

    def get_degenerate(self):
        v = C.c_int(0)
        self._ck(self.L.lvreg_get_degenerate(self.h, C.byref(v)))
        return v.value

    def reset_lm_state(self):
        self._ck(self.L.lvreg_reset_lm_state(self.h))

    # ---- loop closure ----
    def loop_find_near_keyframes(self, key, search_num, slot):
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_loop_find_near_keyframes(self.h, int(key), int(search_num), int(slot), C.byref(n)))
        return n.value

    def build_global_map(self, ids, which=3, leaf=1.0):
        ids = np.ascontiguousarray(ids, np.int32)
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_build_global_map(self.h, ids.ctypes.data_as(C.c_void_p), C.c_size_t(len(ids)), int(which),
                                               C.c_float(leaf), C.byref(n)))
        return self.icp_get_cloud(0)

    def icp_set_cloud(self, slot, cloud):
        c, _keep = _cloud(cloud)
        self._ck(self.L.lvreg_icp_set_cloud(self.h, int(slot), C.byref(c)))

    def icp_get_cloud(self, slot, pcl_layout=False):
        return self._get_cloud(self.L.lvreg_icp_get_cloud, slot, pcl_layout)

    def nn1(self, queries, max_dist=0.0):
        c, _keep = _cloud(queries)
        idx = np.zeros(c.n, np.int32)
        d2 = np.zeros(c.n, np.float32)
        self._ck(self.L.lvreg_nn1(self.h, C.byref(c), C.c_float(max_dist), idx.ctypes.data_as(C.c_void_p),
                                  d2.ctypes.data_as(C.c_void_p)))
        return idx, d2

    def icp_align(self, params=None):
        params = params or icp_default_params()
        res = IcpResult()
        self._ck(self.L.lvreg_icp_align(self.h, C.byref(params), C.byref(res)))
        return res

    def perform_loop_closure(self, key_cur, key_pre, search_num=25, params=None, fitness_gate=0.3):
        params = params or icp_default_params()
        out = LoopResult()
        self._ck(self.L.lvreg_perform_loop_closure(self.h, int(key_cur), int(key_pre), int(search_num),
                                                   C.byref(params), C.c_float(fitness_gate), C.byref(out)))
        return out

    # ---- deskew + range-image projection ----
    def project_cloud(self, raw, layout=LAYOUT_LIVOX, n_scan=4, horizon_scan=6000, downsample_rate=1,
                      sensor=SENSOR_LIVOX, lidar_min_range=0.5, lidar_max_range=1000.0, deskew=False,
                      time_scan_cur=0.0, imu_time=None, imu_rot=None):
        """raw: uint8 [n, stride] AoS (make_raw_cloud).  imu_rot: [k,3] integrated rotations.  -> n_extracted"""
        raw = np.ascontiguousarray(raw, np.uint8)
        rc = RawCloud()
        rc.data = raw.ctypes.data if len(raw) else None
        rc.n = len(raw)
        rc.stride = raw.shape[1] if raw.ndim == 2 else 32
        rc.intensity_offset, rc.ring_offset, rc.time_offset = layout
        rc.on_device = 0
        pp = ProjectionParams()
        pp.n_scan, pp.horizon_scan, pp.downsample_rate, pp.sensor = n_scan, horizon_scan, downsample_rate, sensor
        pp.lidar_min_range, pp.lidar_max_range = lidar_min_range, lidar_max_range
        pp.deskew = int(bool(deskew))
        pp.time_scan_cur = time_scan_cur
        keep = []
        if deskew:
            t = np.ascontiguousarray(imu_time, np.float64)
            r = np.asarray(imu_rot, np.float64)
            cols = [np.ascontiguousarray(r[:, k]) for k in range(3)]
            keep = [t] + cols
            pp.imu_pointer_cur = len(t) - 1
            pp.imu_time = t.ctypes.data
            pp.imu_rot_x, pp.imu_rot_y, pp.imu_rot_z = (c.ctypes.data for c in cols)
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_project_cloud(self.h, C.byref(rc), C.byref(pp), C.byref(n)))
        del keep
        self._proj_n_scan = n_scan
        return n.value

    def download_projection(self):
        """-> (extracted [m,4], point_range, point_col_ind, start_ring_index, end_ring_index)"""
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_download_projection(self.h, None, None, None, None, None, C.byref(n)))
        m, ns = n.value, self._proj_n_scan
        out = np.zeros((m, 4), np.float32)
        rg = np.zeros(m, np.float32)
        col = np.zeros(m, np.int32)
        sr = np.zeros(ns, np.int32)
        er = np.zeros(ns, np.int32)
        co = CloudOut()
        co.data = out.ctypes.data if m else None
        co.capacity = m
        co.stride = 16
        co.intensity_offset = 12
        self._ck(self.L.lvreg_download_projection(self.h, C.byref(co), rg.ctypes.data_as(C.c_void_p),
                                                  col.ctypes.data_as(C.c_void_p), sr.ctypes.data_as(C.c_void_p),
                                                  er.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out, rg, col, sr, er

    def extract_features_projected(self, edge_threshold=1.0, surf_threshold=0.1, surf_leaf=0.4):
        """FeatureExtraction on the device-resident result of the last project_cloud -> (n_corner, n_surf)"""
        cloud = Cloud()
        info = ScanInfo()
        self._ck(self.L.lvreg_get_projection(self.h, C.byref(cloud), C.byref(info)))
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        self._ck(self.L.lvreg_extract_features(self.h, C.byref(cloud), C.byref(info), C.c_float(edge_threshold),
                                               C.c_float(surf_threshold), C.c_float(surf_leaf), None, C.byref(nc),
                                               None, C.byref(ns), None))
        return nc.value, ns.value

    # ---- LiDAR depth for visual features ----
    def depth_clear(self):
        self._ck(self.L.lvreg_depth_clear(self.h))

    def depth_add_cloud(self, cloud, T_now, stamp):
        c, _keep = _cloud(cloud)
        T = np.ascontiguousarray(T_now, np.float32).reshape(12)
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_depth_add_cloud(self.h, C.byref(c), T.ctypes.data_as(C.c_void_p), C.c_double(stamp),
                                              C.byref(n)))
        return n.value

    def depth_set_cloud(self, cloud):
        c, _keep = _cloud(cloud)
        self._ck(self.L.lvreg_depth_set_cloud(self.h, C.byref(c)))

    def depth_get_cloud(self, which=0, pcl_layout=False):
        return self._get_cloud(self.L.lvreg_depth_get_cloud, which, pcl_layout)

    def get_depth(self, T_inv, features_xyz, num_bins=360):
        """-> (depth per feature (-1 = none), features_3d_sphere [n,4])"""
        T = np.ascontiguousarray(T_inv, np.float32).reshape(12)
        f = np.ascontiguousarray(features_xyz, np.float32).reshape(-1, 3)
        depth = np.zeros(len(f), np.float32)
        f3d = np.zeros((len(f), 4), np.float32)
        self._ck(self.L.lvreg_get_depth(self.h, T.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
                                        C.c_size_t(len(f)), int(num_bins), depth.ctypes.data_as(C.c_void_p),
                                        f3d.ctypes.data_as(C.c_void_p)))
        return depth, f3d

    # ---- stage level ----
    def transform_cloud(self, pts, pose):
        c, keep = _cloud(pts)
        pose = np.ascontiguousarray(pose, np.float32)
        out = np.zeros((c.n, 4), np.float32)
        co = CloudOut()
        co.data = out.ctypes.data if c.n else None
        co.capacity = c.n
        co.stride = 16
        co.intensity_offset = 12
        self._ck(self.L.lvreg_transform_cloud(self.h, C.byref(c), pose.ctypes.data_as(C.c_void_p), C.byref(co)))
        return out

    def voxelgrid(self, pts, leaf, pcl_layout_out=False):
        """returns (filtered cloud, voxel idx per output point, passthrough flag)"""
        c, keep = _cloud(pts)
        cols = 8 if pcl_layout_out else 4
        out = np.zeros((max(c.n, 1), cols), np.float32)
        keys = np.zeros(max(c.n, 1), np.uint32)
        co = CloudOut()
        co.data = out.ctypes.data
        co.capacity = c.n
        co.stride = cols * 4
        co.intensity_offset = 16 if pcl_layout_out else 12
        n_out = C.c_size_t(0)
        pt = C.c_int(0)
        self._ck(self.L.lvreg_voxelgrid(self.h, C.byref(c), C.c_float(leaf), C.byref(co), C.byref(n_out),
                                        keys.ctypes.data_as(C.c_void_p), C.byref(pt)))
        return out[:n_out.value].copy(), keys[:n_out.value].copy(), bool(pt.value)

    def voxel_keys(self, pts, leaf):
        c, keep = _cloud(pts)
        keys = np.zeros(max(c.n, 1), np.uint32)
        self._ck(self.L.lvreg_voxel_keys(self.h, C.byref(c), C.c_float(leaf), keys.ctypes.data_as(C.c_void_p)))
        return keys[:c.n].copy()

    def knn5(self, which, queries, variant=KNN_GRID_GATED):
        c, keep = _cloud(queries)
        idx = np.zeros((c.n, 5), np.int32)
        d2 = np.zeros((c.n, 5), np.float32)
        self._ck(self.L.lvreg_knn5(self.h, which, C.byref(c), variant, idx.ctypes.data_as(C.c_void_p),
                                   d2.ctypes.data_as(C.c_void_p)))
        return idx, d2

    def bench_knn5(self, which, queries, variant, repeats=10):
        c, keep = _cloud(queries)
        ms = C.c_float(0)
        self._ck(self.L.lvreg_bench_knn5(self.h, which, C.byref(c), variant, repeats, C.byref(ms)))
        return ms.value

    def bench_residuals(self, which, queries, pose=None, repeats=10):
        c, keep = _cloud(queries)
        pose = np.ascontiguousarray(np.zeros(6) if pose is None else pose, np.float32)
        ms = C.c_float(0)
        self._ck(self.L.lvreg_bench_residuals(self.h, which, C.byref(c), pose.ctypes.data_as(C.c_void_p), repeats, C.byref(ms)))
        return ms.value

    def extract_features(self, pts, point_range, point_col_ind, start_ring, end_ring, edge_threshold=1.0,
                         surf_threshold=0.1, surf_leaf=0.4):
        """FeatureExtraction (featureExtraction.cpp:87-245) -> (corner, surf, label)"""
        c, keep = _cloud(pts)
        rng = np.ascontiguousarray(point_range, np.float32)
        col = np.ascontiguousarray(point_col_ind, np.int32)
        sr = np.ascontiguousarray(start_ring, np.int32)
        er = np.ascontiguousarray(end_ring, np.int32)
        info = ScanInfo(sr.ctypes.data, er.ctypes.data, len(sr), 0, col.ctypes.data, rng.ctypes.data)
        n = c.n
        corner = np.zeros((max(len(sr) * 240, 1), 4), np.float32)
        surf = np.zeros((max(n, 1), 4), np.float32)
        label = np.zeros(max(n, 1), np.int32)
        co, so = CloudOut(), CloudOut()
        for o, a in ((co, corner), (so, surf)):
            o.data = a.ctypes.data
            o.capacity = len(a)
            o.stride = 16
            o.intensity_offset = 12
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        self._ck(self.L.lvreg_extract_features(self.h, C.byref(c), C.byref(info), C.c_float(edge_threshold),
                                               C.c_float(surf_threshold), C.c_float(surf_leaf), C.byref(co), C.byref(nc),
                                               C.byref(so), C.byref(ns), label.ctypes.data_as(C.c_void_p)))
        return corner[:nc.value].copy(), surf[:ns.value].copy(), label[:n].copy()

    def feature_clouds(self):
        """device-resident (corner, surf) descriptors of the last extract_features call"""
        c, s = Cloud(), Cloud()
        self._ck(self.L.lvreg_get_feature_clouds(self.h, C.byref(c), C.byref(s)))
        return c, s

    def bench_sort(self, n, key_bits=28, repeats=5):
        """(ms per whole sort, number of 8-bit passes) for n random (key, index) pairs"""
        ms, passes = C.c_float(0), C.c_int(0)
        self._ck(self.L.lvreg_bench_sort(self.h, C.c_size_t(n), int(key_bits), int(repeats), C.byref(ms), C.byref(passes)))
        return ms.value, passes.value

    def _residuals(self, fn, pts, pose):
        c, keep = _cloud(pts)
        pose = np.ascontiguousarray(pose, np.float32)
        coeff = np.zeros((c.n, 4), np.float32)
        flag = np.zeros(c.n, np.uint8)
        nn = np.zeros((c.n, 5), np.int32)
        self._ck(fn(self.h, C.byref(c), pose.ctypes.data_as(C.c_void_p), coeff.ctypes.data_as(C.c_void_p),
                    flag.ctypes.data_as(C.c_void_p), nn.ctypes.data_as(C.c_void_p)))
        return coeff, flag, nn

    def corner_residuals(self, pts, pose):
        return self._residuals(self.L.lvreg_corner_residuals, pts, pose)

    def surf_residuals(self, pts, pose):
        return self._residuals(self.L.lvreg_surf_residuals, pts, pose)

    def lm_step(self, ori, coeff, it, pose):
        ori = np.ascontiguousarray(ori, np.float32)
        coeff = np.ascontiguousarray(coeff, np.float32)
        pose = np.ascontiguousarray(pose, np.float32).copy()
        AtA = np.zeros((6, 6), np.float32)
        Atb = np.zeros(6, np.float32)
        x = np.zeros(6, np.float32)
        conv = C.c_int(0)
        self._ck(self.L.lvreg_lm_step(self.h, ori.ctypes.data_as(C.c_void_p), coeff.ctypes.data_as(C.c_void_p),
                                      C.c_size_t(len(ori)), int(it), pose.ctypes.data_as(C.c_void_p),
                                      AtA.ctypes.data_as(C.c_void_p), Atb.ctypes.data_as(C.c_void_p),
                                      x.ctypes.data_as(C.c_void_p), C.byref(conv)))
        return conv.value, pose, AtA, Atb, x

    # ---- measurement ----
    def timings(self):
        t = Timings()
        self._ck(self.L.lvreg_get_timings(self.h, C.byref(t)))
        return t

    def iteration_profile(self):
        """[iterations, 4] microseconds: tile work, barrier wait, grid reduction, solve (block 0)"""
        us = np.zeros((MAX_ITERS, 4), np.float32)
        n = C.c_int(0)
        self._ck(self.L.lvreg_get_iteration_profile(self.h, us.ctypes.data_as(C.c_void_p), C.byref(n)))
        return us[:n.value].copy()

    def debug_tile_times(self):
        n = C.c_size_t(0)
        self._ck(self.L.lvreg_debug_tile_times(self.h, None, C.c_size_t(0), C.byref(n)))
        out = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.L.lvreg_debug_tile_times(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(n.value), C.byref(n)))
        return out

    def debug_stage_stats(self):
        out = np.zeros(8, np.uint32)
        self._ck(self.L.lvreg_debug_stage_stats(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def enable_kernel_timing(self, on=True):
        self._ck(self.L.lvreg_enable_kernel_timing(self.h, C.c_int(1 if on else 0)))

    def bucket_kernel_ms(self):
        """(ms[2], points_in[2], voxels_out[2]) of the local-map bucket kernels of the last call (corner, surf)"""
        ms = np.zeros(2, np.float32)
        nin = np.zeros(2, np.uint32)
        nout = np.zeros(2, np.uint32)
        self._ck(self.L.lvreg_get_bucket_kernel_ms(self.h, ms.ctypes.data_as(C.c_void_p), nin.ctypes.data_as(C.c_void_p),
                                                   nout.ctypes.data_as(C.c_void_p)))
        return ms, nin, nout

    def launch_count(self):
        n = C.c_uint64(0)
        self._ck(self.L.lvreg_get_launch_count(self.h, C.byref(n)))
        return n.value
