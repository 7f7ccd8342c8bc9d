"""Wire / disk formats either side of the path: lidar_odometry/msg/CloudInfo.msg:1-32 as a ROS 2 CDR
message body, and the binary PCD files of the save-map service (MO:179-236).

The CDR encoder below is written independently of the library's (struct.pack from the XCDR1 rules: little endian,
4-byte encapsulation header, primitives aligned to their size from the start of the body, sequences and strings
length-prefixed), so equal bytes pin the layout, not just a round trip."""
import ctypes as C
import os
import struct
import tempfile

import numpy as np
import pytest


def _host():
    from lidar_visual_inertial_slam_b200 import harness as H
    L = H.lib()
    L.lvh_cloudinfo_serialize.restype = C.c_size_t
    L.lvh_cloudinfo_serialize.argtypes = [C.c_double, C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                          C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.lvh_cloudinfo_roundtrip.restype = C.c_size_t
    L.lvh_cloudinfo_roundtrip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    L.lvh_load_pcd_xyzi.restype = C.c_longlong
    L.lvh_load_pcd_xyzi.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    L.lvh_mo_save_map.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
    return L


class Cdr:
    def __init__(self):
        self.b = bytearray(b"\x00\x01\x00\x00")

    def align(self, a):
        self.b += b"\x00" * ((a - (len(self.b) - 4) % a) % a)

    def put(self, fmt, v):
        self.align(struct.calcsize(fmt))
        self.b += struct.pack("<" + fmt, v)

    def string(self, s):
        self.put("I", len(s) + 1)
        self.b += s.encode() + b"\x00"

    def seq(self, arr, fmt):
        self.put("I", len(arr))
        if len(arr):
            self.align(struct.calcsize(fmt))
            self.b += np.ascontiguousarray(arr).tobytes()

    def header(self, stamp, frame):
        sec = int(np.floor(stamp))
        self.put("i", sec)
        self.put("I", int(round((stamp - sec) * 1e9)))
        self.string(frame)

    def cloud(self, xyzi, stamp, frame):
        self.header(stamp, frame)
        self.put("I", 1)
        self.put("I", len(xyzi))
        self.put("I", 4)
        for name, off in (("x", 0), ("y", 4), ("z", 8), ("intensity", 16)):
            self.string(name)
            self.put("I", off)
            self.put("B", 7)
            self.put("I", 1)
        self.put("B", 0)
        self.put("I", 32)
        self.put("I", 32 * len(xyzi))
        pcl = np.zeros((len(xyzi), 8), np.float32)            # pcl::PointXYZI: {x, y, z, 1 | intensity, 0, 0, 0}
        pcl[:, :3] = xyzi[:, :3]
        pcl[:, 3] = 1.0
        pcl[:, 4] = xyzi[:, 3]
        self.put("I", pcl.nbytes)
        self.b += pcl.tobytes()
        self.put("B", 1)


def _message(rng, n_ring, n_pts, nd, nc, ns):
    m = dict(stamp=1700000123.250000001, frame="lidar_link",
             start=rng.integers(0, 1000, n_ring).astype(np.int32), end=rng.integers(0, 1000, n_ring).astype(np.int32),
             col=rng.integers(0, 6000, n_pts).astype(np.int32), rng=rng.uniform(1, 80, n_pts).astype(np.float32),
             si=np.array([1, 0, 7], np.int64), sf=rng.normal(0, 1, 9).astype(np.float32),
             deskewed=rng.normal(0, 10, (nd, 4)).astype(np.float32), corner=rng.normal(0, 10, (nc, 4)).astype(np.float32),
             surf=rng.normal(0, 10, (ns, 4)).astype(np.float32))
    w = Cdr()
    w.header(m["stamp"], m["frame"])
    w.seq(m["start"], "i"); w.seq(m["end"], "i"); w.seq(m["col"], "i"); w.seq(m["rng"], "f")
    w.put("q", int(m["si"][0])); w.put("q", int(m["si"][1]))
    for v in m["sf"]:
        w.put("f", float(v))
    w.put("q", int(m["si"][2]))
    for c in (m["deskewed"], m["corner"], m["surf"]):
        w.cloud(c, m["stamp"], m["frame"])
    return m, bytes(w.b)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("shape", [(4, 100, 100, 12, 60), (16, 0, 0, 0, 0), (1, 3, 3, 1, 2), (128, 5000, 5000, 700, 3000)])
def test_cloud_info_cdr_layout_and_round_trip(shape):
    L = _host()
    rng = np.random.default_rng(sum(shape))
    m, ref = _message(rng, *shape)
    size = L.lvh_cloudinfo_serialize(m["stamp"], m["frame"].encode(), _p(m["start"]), _p(m["end"]), len(m["start"]),
                                     _p(m["col"]), _p(m["rng"]), len(m["col"]), _p(m["si"]), _p(m["sf"]),
                                     _p(m["deskewed"]), len(m["deskewed"]), _p(m["corner"]), len(m["corner"]),
                                     _p(m["surf"]), len(m["surf"]), None, 0)
    assert size == len(ref)
    out = np.zeros(size, np.uint8)
    L.lvh_cloudinfo_serialize(m["stamp"], m["frame"].encode(), _p(m["start"]), _p(m["end"]), len(m["start"]),
                              _p(m["col"]), _p(m["rng"]), len(m["col"]), _p(m["si"]), _p(m["sf"]),
                              _p(m["deskewed"]), len(m["deskewed"]), _p(m["corner"]), len(m["corner"]),
                              _p(m["surf"]), len(m["surf"]), _p(out), size)
    assert out.tobytes() == ref                              # the library's bytes are the independently encoded ones
    # decode the reference bytes, check what came out, encode again
    src = np.frombuffer(ref, np.uint8).copy()
    back = np.zeros(size, np.uint8)
    summ = np.zeros(19, np.float64)
    n = L.lvh_cloudinfo_roundtrip(_p(src), len(src), _p(back), size, _p(summ))
    assert n == size and back.tobytes() == ref
    assert abs(summ[0] - m["stamp"]) < 1e-6 and summ[1] == len(m["start"]) and summ[2] == len(m["col"])
    assert list(summ[3:6]) == [1.0, 0.0, 7.0]
    assert np.array_equal(summ[6:15].astype(np.float32), m["sf"])
    assert list(summ[15:18]) == [len(m["deskewed"]), len(m["corner"]), len(m["surf"])]
    want = sum(float(c.astype(np.float64).sum()) for c in (m["deskewed"], m["corner"], m["surf"]))
    assert abs(summ[18] - want) <= 1e-6 * max(1.0, abs(want))


def test_cloud_info_rejects_truncated_and_foreign_messages():
    L = _host()
    rng = np.random.default_rng(3)
    m, ref = _message(rng, 4, 50, 50, 5, 20)
    src = np.frombuffer(ref, np.uint8).copy()
    for cut in (3, 10, len(src) // 2, len(src) - 1):
        assert L.lvh_cloudinfo_roundtrip(_p(src), cut, None, 0, None) == 0
    big_endian = src.copy()
    big_endian[1] = 0
    assert L.lvh_cloudinfo_roundtrip(_p(big_endian), len(big_endian), None, 0, None) == 0


@pytest.mark.gpu
def test_save_map_service_writes_pcl_compatible_pcd_files():
    """saveMapService on the mirror after a short replay: the five files, their PCD headers, and their contents
    against the oracle's transform + concatenate (+ VoxelGrid)"""
    from lidar_visual_inertial_slam_b200 import harness as H
    from oracle import pyoracle as O
    L = _host()
    gen = H.Generator(H.MID360, 0x5EED0003)
    mo = H.MapOptimizationMirror()
    kfs = []
    n_kf = 0
    for k in range(12):
        truth = gen.truth_pose(k, 0.6, 1.0)
        corner, surf = gen.scan(truth, 900 + k, 4)
        st, pose, res, tim, nkf = mo.handle_scan(corner, surf, k * 0.6, truth if k == 0 else gen.guess_pose(k, truth, 0.05, 0.01))
        if nkf > n_kf:
            kfs.append((O.voxelgrid(corner, 0.2)[0], O.voxelgrid(surf, 0.4)[0], pose.copy(), k * 0.6))
            n_kf = nkf
    assert n_kf >= 4

    def load(path):
        n = L.lvh_load_pcd_xyzi(path.encode(), None, 0)
        assert n >= 0
        out = np.zeros((n, 4), np.float32)
        L.lvh_load_pcd_xyzi(path.encode(), _p(out), n)
        return out

    for resolution in (0.0, 0.5):
        with tempfile.TemporaryDirectory() as d:
            assert L.lvh_mo_save_map(mo.mo, d.encode(), C.c_float(resolution)) == 1
            assert sorted(os.listdir(d)) == ["CornerMap.pcd", "GlobalMap.pcd", "SurfMap.pcd", "trajectory.pcd", "transformations.pcd"]
            head = open(os.path.join(d, "transformations.pcd"), "rb").read(400).decode("latin1")
            assert "FIELDS x y z intensity roll pitch yaw time" in head and "SIZE 4 4 4 4 4 4 4 8" in head
            assert ("POINTS %d" % n_kf) in head and "DATA binary" in head
            head = open(os.path.join(d, "GlobalMap.pcd"), "rb").read(300).decode("latin1")
            assert head.startswith("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
                                   "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH ")
            traj = load(os.path.join(d, "trajectory.pcd"))
            assert len(traj) == n_kf and np.array_equal(traj[:, 3], np.arange(n_kf, dtype=np.float32))
            for i, (_, _, pose, _) in enumerate(kfs):
                assert np.array_equal(traj[i, :3], pose[3:])
            gc = np.concatenate([O.transform_cloud(c, p) for c, _, p, _ in kfs])
            gs = np.concatenate([O.transform_cloud(s, p) for _, s, p, _ in kfs])
            assert np.array_equal(load(os.path.join(d, "GlobalMap.pcd")), np.concatenate([gc, gs]))
            want_c = gc if resolution == 0.0 else O.voxelgrid(gc, resolution)[0]
            want_s = gs if resolution == 0.0 else O.voxelgrid(gs, resolution)[0]
            assert np.array_equal(load(os.path.join(d, "CornerMap.pcd")), want_c)
            assert np.array_equal(load(os.path.join(d, "SurfMap.pcd")), want_s)
    mo.close()
