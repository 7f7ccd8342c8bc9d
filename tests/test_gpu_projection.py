"""GPU parity of the front end's per-point work: lvreg_project_cloud (projectPointCloud +
deskewPoint + cloudExtraction, imageProjection.cpp:495-647) against oracle/oracle_projection.cpp,
bit-exact: extracted cloud, point_range, point_col_ind, start / end ring indices; and the chained
raw scan -> features path against the oracle's projection + FeatureExtraction."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O                      # noqa: E402


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


@pytest.fixture(scope="module")
def h(lv):
    hd = lv.Lvreg()
    yield hd
    hd.close()


def raw_scan(rng, n, n_scan, spread=30.0, bad_rings=True):
    """an unordered (Livox-like) point stream: rings interleaved, some points out of range / ring"""
    d = rng.normal(0, 1, (n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.uniform(0.2, spread, n)
    xyz = d * r[:, None]
    xyzi = np.concatenate([xyz, rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    ring = rng.integers(0, n_scan + (2 if bad_rings else 0), n).astype(np.uint16)
    rel = np.sort(rng.uniform(0, 0.1, n)).astype(np.float32)
    return xyzi, ring, rel


def imu_track(rng, t0, k=40):
    t = t0 - 0.01 + np.arange(k) * 0.005 + rng.uniform(0, 1e-4, k)
    w = rng.normal(0, 0.6, (k, 3))
    rot = np.zeros((k, 3))
    for i in range(1, k):
        rot[i] = rot[i - 1] + w[i] * (t[i] - t[i - 1])
    return t, rot


def compare(h, lv, xyzi, ring, rel, layout, **kw):
    raw = lv.make_raw_cloud(xyzi, ring, rel, layout)
    n = h.project_cloud(raw, layout=layout, **kw)
    g = h.download_projection()
    o = O.project_cloud(xyzi, ring, rel, **kw)
    assert n == len(o[0])
    names = ("extracted", "point_range", "point_col_ind", "start_ring_index", "end_ring_index")
    for a, b, name in zip(g, o, names):
        assert np.array_equal(a, b), name
    return g


@pytest.mark.parametrize("n,n_scan,horizon", [(20000, 4, 6000), (50000, 6, 4000), (3000, 4, 500)])
def test_livox_projection_bit_exact(h, lv, n, n_scan, horizon):
    rng = np.random.default_rng(n)
    xyzi, ring, rel = raw_scan(rng, n, n_scan)
    t, rot = imu_track(rng, 100.0)
    # without and with deskew; horizon 500 overflows the per-ring column counter (points dropped)
    compare(h, lv, xyzi, ring, rel, lv.LAYOUT_LIVOX, n_scan=n_scan, horizon_scan=horizon, sensor=2, lidar_min_range=1.0,
            lidar_max_range=25.0)
    g = compare(h, lv, xyzi, ring, rel, lv.LAYOUT_LIVOX, n_scan=n_scan, horizon_scan=horizon, sensor=2, lidar_min_range=1.0,
                lidar_max_range=25.0, deskew=True, time_scan_cur=100.0, imu_time=t, imu_rot=rot)
    assert len(g[0]) > 0
    compare(h, lv, xyzi, ring, rel, lv.LAYOUT_LIVOX, n_scan=n_scan, horizon_scan=horizon, sensor=2, downsample_rate=2,
            deskew=True, time_scan_cur=100.0, imu_time=t, imu_rot=rot)


@pytest.mark.parametrize("sensor", [0, 1])
def test_spinning_lidar_projection_bit_exact(h, lv, sensor):
    rng = np.random.default_rng(50 + sensor)
    n_scan, horizon = 16, 1800
    # ring-major spinning scan with duplicates landing in the same cell (first one must win)
    az = np.tile(np.linspace(-np.pi, np.pi, 2200, endpoint=False), n_scan)
    ring = np.repeat(np.arange(n_scan), 2200).astype(np.uint16)
    el = np.deg2rad(-15 + 2.0 * ring)
    r = rng.uniform(2, 60, len(az))
    xyzi = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el), rng.uniform(0, 255, len(az))], 1).astype(np.float32)
    order = rng.permutation(len(az))
    xyzi, ring = xyzi[order], ring[order]
    rel = np.sort(rng.uniform(0, 0.1, len(az))).astype(np.float32)
    t, rot = imu_track(rng, 5.0)
    g = compare(h, lv, xyzi, ring, rel, lv.LAYOUT_VELODYNE, n_scan=n_scan, horizon_scan=horizon, sensor=sensor, deskew=True,
                time_scan_cur=5.0, imu_time=t, imu_rot=rot, lidar_min_range=1.0, lidar_max_range=1000.0)
    assert len(g[0]) > n_scan * horizon // 2


def test_projection_edge_cases(h, lv):
    rng = np.random.default_rng(9)
    # empty scan
    e = np.zeros((0, 4), np.float32)
    compare(h, lv, e, np.zeros(0, np.uint16), np.zeros(0, np.float32), lv.LAYOUT_LIVOX, n_scan=4, horizon_scan=100, sensor=2)
    # everything out of range
    xyzi, ring, rel = raw_scan(rng, 500, 4)
    compare(h, lv, xyzi, ring, rel, lv.LAYOUT_LIVOX, n_scan=4, horizon_scan=100, sensor=2, lidar_min_range=500.0)
    # point times before the first / after the last IMU sample, and exactly on a sample
    t = 10.0 + np.arange(6) * 0.02
    rot = rng.normal(0, 0.05, (6, 3)).cumsum(0)
    rel = np.array([0.0, 0.02, 0.03, 0.1, 0.5, 0.0399999], np.float32)
    xyzi = rng.uniform(-10, 10, (6, 4)).astype(np.float32)
    for tsc in (9.9, 10.0, 10.05):
        compare(h, lv, xyzi, np.zeros(6, np.uint16), rel, lv.LAYOUT_LIVOX, n_scan=1, horizon_scan=10, sensor=2, deskew=True,
                time_scan_cur=tsc, imu_time=t, imu_rot=rot, lidar_min_range=0.0)
    for pt in (9.0, 10.0, 10.01, 10.02, 10.1, 11.0):
        assert O.find_rotation(pt, t, rot).shape == (3,)


def test_raw_scan_to_features_on_device(h, lv):
    """project_cloud -> extract_features with no host copy of the deskewed cloud in between"""
    from tests.synth import ring_scan
    rng = np.random.default_rng(3)
    pts, rg, col, sr, er = ring_scan(rng, 16, 900)
    # rebuild a raw stream from the ring-ordered cloud: ring = intensity channel of ring_scan's points
    ring = np.zeros(len(pts), np.uint16)
    for r in range(16):
        a, b = sr[r] - 4, er[r] + 6
        ring[a:b] = r
    rel = np.linspace(0, 0.1, len(pts)).astype(np.float32)
    raw = lv.make_raw_cloud(pts, ring, rel, lv.LAYOUT_VELODYNE)
    kw = dict(n_scan=16, horizon_scan=900, sensor=0, lidar_min_range=0.1, lidar_max_range=1000.0)
    n = h.project_cloud(raw, layout=lv.LAYOUT_VELODYNE, **kw)
    nc, ns = h.extract_features_projected(edge_threshold=0.5)
    oe, org, ocol, osr, oer = O.project_cloud(pts, ring, rel, **kw)
    oc, os_, ol = O.extract_features(oe, org, ocol, osr, oer, edge_threshold=0.5)
    assert n == len(oe)
    assert (nc, ns) == (len(oc), len(os_))
    assert nc > 10 and ns > 100


def test_projection_golden_vectors(h, lv):
    from tests.test_golden_cpu import _golden_projection
    z, kw, want = _golden_projection()
    raw = lv.make_raw_cloud(z["xyzi"], z["ring"], z["rel_time"], lv.LAYOUT_LIVOX)
    h.project_cloud(raw, layout=lv.LAYOUT_LIVOX, **kw)
    for a, b in zip(h.download_projection(), want):
        assert np.array_equal(a, b)


def test_front_end_mirrors_raw_scan_to_pose(lv):
    """C++ mirrors of ImageProjection + FeatureExtraction + mapOptimization on one handle: a raw ring scan goes
    to a registered pose with nothing but the raw points uploaded; checked against the oracle chain."""
    from lidar_visual_inertial_slam_b200 import harness as H
    from tests.synth import ring_scan, rot_rpy
    rng = np.random.default_rng(13)
    pts, rg, col, sr, er = ring_scan(rng, 16, 900)
    ring = np.zeros(len(pts), np.uint16)
    for r in range(16):
        ring[sr[r] - 4:er[r] + 6] = r
    rel = np.linspace(0, 0.1, len(pts)).astype(np.float32)
    kw = dict(n_scan=16, horizon_scan=900, sensor=0, lidar_min_range=0.1, lidar_max_range=1000.0)
    # oracle chain: projection -> features; the features double as the keyframe (identity pose) = the map
    oe, org, ocol, osr, oer = O.project_cloud(pts, ring, rel, **kw)
    oc, os_, ol = O.extract_features(oe, org, ocol, osr, oer, edge_threshold=0.5)
    mo = H.MapOptimizationMirror()
    st, pose0, res0, tim, nkf = mo.handle_scan(oc, os_, 0.0, np.zeros(6, np.float32))
    assert nkf == 1
    omo = O.MapOptimization()
    omo.add_keyframe(O.voxelgrid(oc, 0.2)[0], O.voxelgrid(os_, 0.4)[0], np.zeros(6, np.float32), 0.0)
    omo.build_local_map([0])
    # the same scan seen from a slightly different pose: transform the raw points into that sensor frame
    true = np.array([0.01, -0.008, 0.03, 0.12, -0.08, 0.02], np.float32)
    R = rot_rpy(*true[:3])
    moved = pts.copy()
    moved[:, :3] = ((pts[:, :3].astype(np.float64) - true[3:].astype(np.float64)) @ R).astype(np.float32)
    raw = lv.make_raw_cloud(moved, ring, rel, lv.LAYOUT_LIVOX)
    guess = true + np.array([0.004, -0.003, 0.01, 0.03, -0.02, 0.01], np.float32)
    st, pose, res, (n_ext, n_c, n_s) = mo.raw_scan_to_pose(raw, 16, 900, 0, guess, ids=[0], min_range=0.1, edge_threshold=0.5)
    me, mrg, mcol, msr, mer = O.project_cloud(moved, ring, rel, **kw)
    mc, ms, _ = O.extract_features(me, mrg, mcol, msr, mer, edge_threshold=0.5)
    assert (n_ext, n_c, n_s) == (len(me), len(mc), len(ms))
    opose, ores, nc, ns = omo.register_scan(mc, ms, guess)
    assert st == lv.OK and ores.status == 0
    assert res.iterations == ores.iterations and res.converged == ores.converged
    assert np.abs(pose[:3] - opose[:3]).max() <= 1e-5 and np.abs(pose[3:] - opose[3:]).max() <= 1e-4
    assert np.abs(pose[3:] - true[3:]).max() < 0.05
    mo.close()
