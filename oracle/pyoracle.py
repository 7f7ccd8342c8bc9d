"""ctypes binding of the CPU oracle (oracle/liblvreg_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblvreg_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


class Params(C.Structure):
    _fields_ = [("corner_leaf", C.c_float), ("surf_leaf", C.c_float),
                ("edge_min_valid", C.c_int), ("surf_min_valid", C.c_int),
                ("max_iters", C.c_int), ("knn_gate_sq", C.c_float),
                ("line_eig_ratio", C.c_float), ("plane_tol", C.c_float),
                ("min_weight", C.c_float), ("min_matches", C.c_int),
                ("degeneracy_eig", C.c_float), ("conv_deg", C.c_float), ("conv_cm", C.c_float),
                ("reference_quirks", C.c_int), ("keyframe_search_radius", C.c_float),
                ("keyframe_density", C.c_float), ("rotation_tolerance", C.c_float),
                ("z_tolerance", C.c_float), ("num_threads", C.c_int)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("converged", C.c_int),
                ("degenerate", C.c_int), ("n_sel", C.c_int * 32),
                ("pose_iter", (C.c_float * 6) * 32)]


class LmState(C.Structure):
    _fields_ = [("is_degenerate", C.c_int), ("matP", C.c_float * 36)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_voxelgrid.restype = C.c_size_t
        _lib.orc_kdtree_build.restype = C.c_void_p
        _lib.orc_kdtree_radius.restype = C.c_size_t
        _lib.orc_mo_create.restype = C.c_void_p
        _lib.orc_mo_num_keyframes.restype = C.c_size_t
        _lib.orc_mo_extract_nearby.restype = C.c_size_t
        _lib.orc_mo_map_size.restype = C.c_size_t
        _lib.orc_mo_loop_find_near_keyframes.restype = C.c_size_t
        _lib.orc_mo_build_global_map.restype = C.c_size_t
        _lib.orc_project_cloud.restype = C.c_size_t
        _lib.orc_depth_create.restype = C.c_void_p
        _lib.orc_depth_add_cloud.restype = C.c_size_t
        _lib.orc_depth_cloud_size.restype = C.c_size_t
        _lib.orc_get_depth.restype = C.c_size_t
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


# ---- dense math -------------------------------------------------------------------------
def jacobi_eigen(A):
    A = _f32(A)
    n = A.shape[0]
    w = np.zeros(n, np.float32)
    v = np.zeros((n, n), np.float32)
    lib().orc_jacobi_eigen(_p(A), n, _p(w), _p(v))
    return w, v


def qr_solve(A, b):
    A = _f32(A)
    b = _f32(b).reshape(A.shape[0], -1)
    x = np.zeros_like(b)
    ok = lib().orc_qr_solve(_p(A), _p(b), A.shape[0], b.shape[1], _p(x))
    return ok, x


def lu_solve(A, B):
    A = _f32(A)
    B = _f32(B).reshape(A.shape[0], -1)
    X = np.zeros_like(B)
    ok = lib().orc_lu_solve(_p(A), _p(B), A.shape[0], B.shape[1], _p(X))
    return ok, X


def gemm(A, B):
    A = _f32(A)
    B = _f32(B)
    out = np.zeros((A.shape[0], B.shape[1]), np.float32)
    lib().orc_gemm(_p(A), _p(B), A.shape[0], A.shape[1], B.shape[1], _p(out))
    return out


def normal_equations(A, b):
    A = _f32(A)
    b = _f32(b)
    AtA = np.zeros((6, 6), np.float32)
    Atb = np.zeros(6, np.float32)
    lib().orc_normal_equations(_p(A), _p(b), A.shape[0], _p(AtA), _p(Atb))
    return AtA, Atb


def plane_fit(A5x3, b5=None):
    A = _f32(A5x3)
    b = _f32(np.full(5, -1.0) if b5 is None else b5)
    x = np.zeros(3, np.float32)
    lib().orc_colpiv_qr_solve_5x3(_p(A), _p(b), _p(x))
    return x


# ---- clouds -----------------------------------------------------------------------------
def pose_to_affine(pose):
    pose = _f32(pose)
    T = np.zeros(12, np.float32)
    lib().orc_pose_to_affine(_p(pose), _p(T))
    return T


def transform_cloud(pts, pose=None, T=None, num_threads=1):
    pts = _f32(pts)
    if T is None:
        T = pose_to_affine(pose)
    T = _f32(T)
    out = np.zeros_like(pts)
    lib().orc_transform_cloud(_p(pts), C.c_size_t(len(pts)), _p(T), _p(out), num_threads)
    return out


def voxelgrid(pts, leaf):
    """returns (out points, per-input keys, per-output keys, passthrough)"""
    pts = _f32(pts)
    n = len(pts)
    out = np.zeros((max(n, 1), 4), np.float32)
    keys = np.zeros(max(n, 1), np.uint32)
    okeys = np.zeros(max(n, 1), np.uint32)
    pt = C.c_int(0)
    m = lib().orc_voxelgrid(_p(pts), C.c_size_t(n), C.c_float(leaf), _p(out), _p(keys), _p(okeys),
                            C.byref(pt))
    return out[:m].copy(), keys[:n].copy(), okeys[:m].copy(), bool(pt.value)


# ---- kNN ----------------------------------------------------------------------------------
def knn5_brute(map_pts, queries, num_threads=8):
    map_pts = _f32(map_pts)
    queries = _f32(queries)
    nq = len(queries)
    idx = np.zeros((nq, 5), np.int32)
    d2 = np.zeros((nq, 5), np.float32)
    lib().orc_knn5_brute(_p(map_pts), C.c_size_t(len(map_pts)), _p(queries), C.c_size_t(nq),
                         _p(idx), _p(d2), num_threads)
    return idx, d2


class KdTree:
    def __init__(self, map_pts):
        self.map = _f32(map_pts)
        self.h = C.c_void_p(lib().orc_kdtree_build(_p(self.map), C.c_size_t(len(self.map))))

    def knn(self, queries, k=5, num_threads=8):
        queries = _f32(queries)
        nq = len(queries)
        idx = np.zeros((nq, k), np.int32)
        d2 = np.zeros((nq, k), np.float32)
        lib().orc_kdtree_knn(self.h, _p(queries), C.c_size_t(nq), k, _p(idx), _p(d2), num_threads)
        return idx, d2

    def radius(self, q, radius):
        q = _f32(q)
        cap = len(self.map)
        idx = np.zeros(cap, np.int32)
        d2 = np.zeros(cap, np.float32)
        n = lib().orc_kdtree_radius(self.h, _p(q), C.c_float(radius), _p(idx), _p(d2), C.c_size_t(cap))
        return idx[:n].copy(), d2[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kdtree_free(self.h)
            self.h = None


# ---- registration -----------------------------------------------------------------------
def _residuals(fn, map_pts, pts, pose, params, tree):
    map_pts = _f32(map_pts)
    pts = _f32(pts)
    pose = _f32(pose)
    tree = tree or KdTree(map_pts)
    n = len(pts)
    coeff = np.zeros((n, 4), np.float32)
    flag = np.zeros(n, np.uint8)
    nn = np.zeros((n, 5), np.int32)
    fn(_p(map_pts), C.c_size_t(len(map_pts)), tree.h, _p(pts), C.c_size_t(n), _p(pose),
       C.byref(params), _p(coeff), _p(flag), _p(nn))
    return coeff, flag, nn


def corner_residuals(map_pts, pts, pose, params=None, tree=None):
    return _residuals(lib().orc_corner_residuals, map_pts, pts, pose, params or default_params(), tree)


def surf_residuals(map_pts, pts, pose, params=None, tree=None):
    return _residuals(lib().orc_surf_residuals, map_pts, pts, pose, params or default_params(), tree)


def jacobian_rows(ori, coeff, pose):
    ori = _f32(ori)
    coeff = _f32(coeff)
    pose = _f32(pose)
    n = len(ori)
    A = np.zeros((n, 6), np.float32)
    b = np.zeros(n, np.float32)
    lib().orc_jacobian_rows(_p(ori), _p(coeff), C.c_size_t(n), _p(pose), _p(A), _p(b))
    return A, b


def lm_step(ori, coeff, it, pose, state=None, params=None):
    ori = _f32(ori)
    coeff = _f32(coeff)
    pose = _f32(pose).copy()
    state = state or LmState()
    params = params or default_params()
    AtA = np.zeros((6, 6), np.float32)
    Atb = np.zeros(6, np.float32)
    x = np.zeros(6, np.float32)
    conv = lib().orc_lm_step(_p(ori), _p(coeff), C.c_size_t(len(ori)), it, _p(pose),
                             C.byref(state), C.byref(params), _p(AtA), _p(Atb), _p(x))
    return conv, pose, AtA, Atb, x, state


def scan2map(corner_map, surf_map, corner, surf, pose, params=None, state=None):
    corner_map, surf_map, corner, surf = map(_f32, (corner_map, surf_map, corner, surf))
    pose = _f32(pose).copy()
    params = params or default_params()
    state = state or LmState()
    res = Result()
    lib().orc_scan2map(_p(corner_map), C.c_size_t(len(corner_map)), _p(surf_map),
                       C.c_size_t(len(surf_map)), _p(corner), C.c_size_t(len(corner)),
                       _p(surf), C.c_size_t(len(surf)), _p(pose), C.byref(state),
                       C.byref(params), C.byref(res))
    return pose, res, state


def transform_update(pose, imu_available=False, imu_roll=0.0, imu_pitch=0.0, imu_weight=0.01,
                     params=None):
    pose = _f32(pose).copy()
    params = params or default_params()
    lib().orc_transform_update(_p(pose), int(imu_available), C.c_float(imu_roll),
                               C.c_float(imu_pitch), C.c_float(imu_weight), C.byref(params))
    return pose


def extract_features(pts, point_range, point_col_ind, start_ring, end_ring, edge_threshold=1.0,
                     surf_threshold=0.1, surf_leaf=0.4):
    """FeatureExtraction (featureExtraction.cpp:87-245): returns (corner, surf, label)"""
    pts = _f32(pts)
    rng = _f32(point_range)
    col = np.ascontiguousarray(point_col_ind, np.int32)
    sr = np.ascontiguousarray(start_ring, np.int32)
    er = np.ascontiguousarray(end_ring, np.int32)
    n = len(pts)
    corner = np.zeros((max(n, 1), 4), np.float32)
    surf = np.zeros((max(n, 1), 4), np.float32)
    label = np.zeros(max(n, 1), np.int32)
    nc, ns = C.c_size_t(0), C.c_size_t(0)
    lib().orc_extract_features(_p(pts), C.c_size_t(n), _p(rng), _p(col), _p(sr), _p(er), len(sr),
                               C.c_float(edge_threshold), C.c_float(surf_threshold), C.c_float(surf_leaf),
                               _p(corner), C.byref(nc), _p(surf), C.byref(ns), _p(label))
    return corner[:nc.value].copy(), surf[:ns.value].copy(), label[:n].copy()


# ---- loop-closure ICP, oracle_icp.cpp ------------------------------------------
class IcpParams(C.Structure):
    _fields_ = [("max_corr_dist", C.c_float), ("max_iterations", C.c_int),
                ("transformation_epsilon", C.c_double), ("euclidean_fitness_epsilon", C.c_double),
                ("num_threads", C.c_int)]


class IcpResult(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_int), ("state", C.c_int),
                ("n_correspondences", C.c_int), ("fitness", C.c_double), ("mse", C.c_double),
                ("final_transformation", C.c_float * 16)]

    @property
    def T(self):
        return np.array(self.final_transformation[:], np.float32).reshape(4, 4)


class LoopResult(C.Structure):
    _fields_ = [("status", C.c_int), ("n_source", C.c_int), ("n_target", C.c_int), ("icp", IcpResult),
                ("pose_from", C.c_float * 6), ("pose_to", C.c_float * 6), ("noise", C.c_float)]


def icp_default_params(**kw):
    p = IcpParams()
    lib().orc_icp_default_params(C.byref(p))
    p.num_threads = 8
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def umeyama_from_moments(mom17):
    mom = np.ascontiguousarray(mom17, np.float64)
    T = np.zeros(16, np.float32)
    lib().orc_umeyama_from_moments(_p(mom), _p(T))
    return T.reshape(4, 4)


def nn1(tgt, queries, num_threads=8):
    tgt = _f32(tgt)
    queries = _f32(queries)
    idx = np.zeros(len(queries), np.int32)
    d2 = np.zeros(len(queries), np.float32)
    lib().orc_nn1(_p(tgt), C.c_size_t(len(tgt)), _p(queries), C.c_size_t(len(queries)), _p(idx), _p(d2),
                  C.c_int(num_threads))
    return idx, d2


def icp_align(src, tgt, params=None):
    src = _f32(src)
    tgt = _f32(tgt)
    params = params or icp_default_params()
    res = IcpResult()
    lib().orc_icp_align(_p(src), C.c_size_t(len(src)), _p(tgt), C.c_size_t(len(tgt)), C.byref(params),
                        C.byref(res))
    return res


def correct_pose(correction4x4, pose):
    T = _f32(correction4x4).reshape(16)
    pose = _f32(pose)
    out = np.zeros(6, np.float32)
    lib().orc_correct_pose(_p(T), _p(pose), _p(out))
    return out


# ---- deskew + range-image projection, oracle_projection.cpp ------------------
class ProjectionParams(C.Structure):
    _fields_ = [("n_scan", C.c_int), ("horizon_scan", C.c_int), ("downsample_rate", C.c_int), ("sensor", C.c_int),
                ("lidar_min_range", C.c_float), ("lidar_max_range", C.c_float), ("deskew", C.c_int),
                ("imu_pointer_cur", C.c_int), ("time_scan_cur", C.c_double), ("imu_time", C.c_void_p),
                ("imu_rot_x", C.c_void_p), ("imu_rot_y", C.c_void_p), ("imu_rot_z", C.c_void_p)]


def find_rotation(point_time, imu_time, imu_rot):
    t = np.ascontiguousarray(imu_time, np.float64)
    r = np.asarray(imu_rot, np.float64)
    cols = [np.ascontiguousarray(r[:, k]) for k in range(3)]
    rot = np.zeros(3, np.float32)
    lib().orc_find_rotation(C.c_double(point_time), _p(t), _p(cols[0]), _p(cols[1]), _p(cols[2]), C.c_int(len(t) - 1), _p(rot))
    return rot


def project_cloud(xyzi, ring, rel_time, n_scan=4, horizon_scan=6000, downsample_rate=1, sensor=2,
                  lidar_min_range=0.5, lidar_max_range=1000.0, deskew=False, time_scan_cur=0.0, imu_time=None,
                  imu_rot=None):
    """projectPointCloud + cloudExtraction -> (extracted, point_range, point_col_ind, start_ring, end_ring)"""
    pts = _f32(xyzi)
    ring = np.ascontiguousarray(ring, np.uint16)
    rel = _f32(rel_time)
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
    cells = n_scan * horizon_scan
    out = np.zeros((cells, 4), np.float32)
    rg = np.zeros(cells, np.float32)
    col = np.zeros(cells, np.int32)
    sr = np.zeros(n_scan, np.int32)
    er = np.zeros(n_scan, np.int32)
    m = lib().orc_project_cloud(_p(pts), _p(ring), _p(rel), C.c_size_t(len(pts)), C.byref(pp), _p(out), _p(rg), _p(col),
                                _p(sr), _p(er))
    del keep
    return out[:m].copy(), rg[:m].copy(), col[:m].copy(), sr, er


# ---- LiDAR depth for visual features, oracle_depth.cpp ------------------------
class DepthRegister:
    """lidar_callback's cloud queue + depthCloud (feature_tracker_node.cpp:273-375)"""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_depth_create())

    def add_cloud(self, cloud, T_now, stamp):
        cloud = _f32(cloud)
        T = _f32(T_now).reshape(12)
        return lib().orc_depth_add_cloud(self.h, _p(cloud), C.c_size_t(len(cloud)), _p(T), C.c_double(stamp))

    def cloud(self):
        n = lib().orc_depth_cloud_size(self.h)
        out = np.zeros((n, 4), np.float32)
        if n:
            lib().orc_depth_get_cloud(self.h, _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_depth_destroy(self.h)
            self.h = None


def get_depth(depth_cloud, T_inv, features_xyz, num_bins=360):
    """DepthRegister::get_depth -> (depth [n], features_3d_sphere [n,4], filtered local cloud [k,4])"""
    dc = _f32(depth_cloud)
    T = _f32(T_inv).reshape(12)
    f = _f32(features_xyz).reshape(-1, 3)
    depth = np.zeros(len(f), np.float32)
    f3d = np.zeros((len(f), 4), np.float32)
    local = np.zeros((max(len(dc), 1), 4), np.float32)
    k = lib().orc_get_depth(_p(dc), C.c_size_t(len(dc)), _p(T), _p(f), C.c_size_t(len(f)), C.c_int(num_bins),
                            _p(depth), _p(f3d), _p(local))
    return depth, f3d, local[:k].copy()


class MapOptimization:
    """mapOptimization-like object: keyframes, local map, per-scan registration."""

    def __init__(self, params=None):
        self.params = params or default_params()
        self.h = C.c_void_p(lib().orc_mo_create(C.byref(self.params)))

    def add_keyframe(self, corner, surf, pose, time):
        corner = _f32(corner)
        surf = _f32(surf)
        pose = _f32(pose)
        return lib().orc_mo_add_keyframe(self.h, _p(corner), C.c_size_t(len(corner)), _p(surf),
                                         C.c_size_t(len(surf)), _p(pose), C.c_double(time))

    def num_keyframes(self):
        return lib().orc_mo_num_keyframes(self.h)

    def extract_nearby(self, time_now):
        cap = 2 * self.num_keyframes() + 8
        ids = np.zeros(cap, np.int32)
        n = lib().orc_mo_extract_nearby(self.h, C.c_double(time_now), _p(ids), C.c_size_t(cap))
        return ids[:n].copy()

    def build_local_map(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        lib().orc_mo_build_local_map(self.h, _p(ids), C.c_size_t(len(ids)))

    def get_map(self, which):
        n = lib().orc_mo_map_size(self.h, which)
        out = np.zeros((n, 4), np.float32)
        lib().orc_mo_get_map(self.h, which, _p(out))
        return out

    def register_scan(self, corner_raw, surf_raw, pose):
        corner_raw = _f32(corner_raw)
        surf_raw = _f32(surf_raw)
        pose = _f32(pose).copy()
        res = Result()
        nc = C.c_size_t(0)
        ns = C.c_size_t(0)
        lib().orc_mo_register_scan(self.h, _p(corner_raw), C.c_size_t(len(corner_raw)),
                                   _p(surf_raw), C.c_size_t(len(surf_raw)), _p(pose),
                                   C.byref(res), C.byref(nc), C.byref(ns))
        return pose, res, nc.value, ns.value

    def loop_find_near_keyframes(self, key, search_num, slot=0):
        n = lib().orc_mo_loop_find_near_keyframes(self.h, C.c_int(key), C.c_int(search_num), C.c_int(slot))
        out = np.zeros((n, 4), np.float32)
        if n:
            lib().orc_mo_get_loop_cloud(self.h, C.c_int(slot), _p(out))
        return out

    def build_global_map(self, ids, which=3, leaf=1.0):
        ids = np.ascontiguousarray(ids, np.int32)
        n = lib().orc_mo_build_global_map(self.h, _p(ids), C.c_size_t(len(ids)), C.c_int(which), C.c_float(leaf))
        out = np.zeros((n, 4), np.float32)
        if n:
            lib().orc_mo_get_loop_cloud(self.h, C.c_int(0), _p(out))
        return out

    def detect_loop_closure_distance(self, time_cur, radius=15.0, time_diff=30.0):
        cur = C.c_int(-1)
        pre = C.c_int(-1)
        ok = lib().orc_mo_detect_loop_closure_distance(self.h, C.c_double(time_cur), C.c_float(radius),
                                                       C.c_float(time_diff), C.byref(cur), C.byref(pre))
        return (cur.value, pre.value) if ok else None

    def perform_loop_closure(self, key_cur, key_pre, search_num=25, params=None, fitness_gate=0.3):
        params = params or icp_default_params()
        out = LoopResult()
        lib().orc_mo_perform_loop_closure(self.h, C.c_int(key_cur), C.c_int(key_pre), C.c_int(search_num),
                                          C.byref(params), C.c_float(fitness_gate), C.byref(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mo_destroy(self.h)
            self.h = None
