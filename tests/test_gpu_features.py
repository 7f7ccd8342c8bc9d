"""GPU parity of the "next" row 8f-1: FeatureExtraction (featureExtraction.cpp:87-245) on the device
against the oracle restatement -- labels, corner cloud (order included) and the per-ring
VoxelGrid of the surface cloud, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O          # noqa: E402
from tests.synth import ring_scan          # noqa: E402


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


@pytest.mark.parametrize("n_scan,horizon,seed", [(16, 1800, 1), (4, 6000, 2), (64, 512, 3)])
def test_extract_features_bit_exact(lv, n_scan, horizon, seed):
    rng = np.random.default_rng(seed)
    pts, rg, col, sr, er = ring_scan(rng, n_scan, horizon)
    rc, rs, rl = O.extract_features(pts, rg, col, sr, er)
    assert len(rc) > 20 and len(rs) > 200            # the scene has both kinds of features
    h = lv.Lvreg()
    c, s, l = h.extract_features(pts, rg, col, sr, er)
    assert np.array_equal(l, rl)
    assert np.array_equal(c, rc)
    assert np.array_equal(s, rs)
    # thresholds / leaf are parameters
    rc2, rs2, rl2 = O.extract_features(pts, rg, col, sr, er, edge_threshold=0.05, surf_threshold=0.01, surf_leaf=0.2)
    c2, s2, l2 = h.extract_features(pts, rg, col, sr, er, edge_threshold=0.05, surf_threshold=0.01, surf_leaf=0.2)
    assert np.array_equal(l2, rl2) and np.array_equal(c2, rc2) and np.array_equal(s2, rs2)
    assert len(rc2) > len(rc)
    h.close()


def test_golden_features(lv):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", "features.npz"))
    h = lv.Lvreg()
    c, s, l = h.extract_features(z["pts"], z["point_range"], z["point_col_ind"], z["start_ring_index"],
                                 z["end_ring_index"], edge_threshold=float(z["edge_threshold"]))
    assert np.array_equal(c, z["corner"]) and np.array_equal(s, z["surf"]) and np.array_equal(l, z["label"])
    h.close()


def test_extract_features_edge_cases(lv):
    h = lv.Lvreg()
    # empty cloud
    c, s, l = h.extract_features(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32),
                                 np.array([4], np.int32), np.array([-6], np.int32))
    assert len(c) == 0 and len(s) == 0
    # rings too short to hold a sector, and one empty ring
    rng = np.random.default_rng(5)
    pts, rg, col, sr, er = ring_scan(rng, 3, 40, drop=0.0)
    sr = np.r_[sr, [sr[-1] + 45]].astype(np.int32)         # a fourth ring with no points: start = count+4, end = count-6
    er = np.r_[er, [er[-1] - 0]].astype(np.int32)
    er[-1] = len(pts) - 1 - 5
    sr[-1] = len(pts) - 1 + 5
    rc, rs, rl = O.extract_features(pts, rg, col, sr, er)
    c, s, l = h.extract_features(pts, rg, col, sr, er)
    assert np.array_equal(l, rl) and np.array_equal(c, rc) and np.array_equal(s, rs)
    h.close()


def test_features_feed_registration_on_device(lv):
    """raw ring-ordered scan -> features -> downsample -> scan2map without a host round trip of the features"""
    rng = np.random.default_rng(9)
    pts, rg, col, sr, er = ring_scan(rng, 32, 1024)
    rc, rs, _ = O.extract_features(pts, rg, col, sr, er, edge_threshold=0.05)
    h = lv.Lvreg()
    c, s, _ = h.extract_features(pts, rg, col, sr, er, edge_threshold=0.05)
    assert np.array_equal(c, rc) and np.array_equal(s, rs)
    # the scan is its own map (keyframe at the origin); start from a perturbed pose
    h.add_keyframe(O.voxelgrid(rc, 0.2)[0], O.voxelgrid(rs, 0.4)[0], np.zeros(6, np.float32))
    h.build_local_map([0])
    dc, dsf = h.feature_clouds()
    assert dc.on_device == 1 and dc.n == len(rc) and dsf.n == len(rs)
    guess = np.array([0.01, -0.01, 0.02, 0.05, -0.04, 0.02], np.float32)
    pose, res, st = h.register_scan(dc, dsf, None, guess)
    mo = O.MapOptimization()
    mo.add_keyframe(O.voxelgrid(rc, 0.2)[0], O.voxelgrid(rs, 0.4)[0], np.zeros(6, np.float32), 0.0)
    mo.build_local_map([0])
    rpose, rres, _, _ = mo.register_scan(rc, rs, guess)
    assert st == (lv.OK if rres.status == 0 else lv.ERR_NOT_ENOUGH_FEATURES)
    if rres.status == 0:
        assert res.iterations == rres.iterations
        assert np.abs(pose[:3] - rpose[:3]).max() <= 1e-5 and np.abs(pose[3:] - rpose[3:]).max() <= 1e-4
    h.close()


def test_extract_features_rejects_malformed_ring_indices(lv):
    """start / end ring indices come from the caller (CloudInfo): a processed ring must lie inside the cloud and
    the rings must be disjoint and ascending; anything else is LVREG_ERR_INVALID, not device memory corruption"""
    rng = np.random.default_rng(6)
    pts, rg, col, sr, er = ring_scan(rng, 4, 400)
    h = lv.Lvreg()
    ref = h.extract_features(pts, rg, col, sr, er)
    bad = []
    s2 = sr.copy(); s2[0] = -3; bad.append((s2, er))                        # start before the cloud
    e2 = er.copy(); e2[-1] = len(pts) + 7; bad.append((sr, e2))             # end past the cloud
    s3 = sr.copy(); s3[2] = sr[1]; bad.append((s3, er))                     # ring 2 overlaps ring 1
    s4, e4 = sr[::-1].copy(), er[::-1].copy(); bad.append((s4, e4))         # descending rings
    s5 = sr.copy(); s5[1] = -(2 ** 31) + 2; bad.append((s5, er))            # overflow bait
    for s_, e_ in bad:
        with pytest.raises(lv.LvregError) as ei:
            h.extract_features(pts, rg, col, s_, e_)
        assert ei.value.status == lv.ERR_INVALID
    # the handle is still usable and gives the same answer
    again = h.extract_features(pts, rg, col, sr, er)
    for a, b in zip(ref, again):
        assert np.array_equal(a, b)
    h.close()
