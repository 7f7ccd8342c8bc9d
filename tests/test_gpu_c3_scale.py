"""GPU parity at BASELINE C3 scale, against the CPU restatement (oracle/), on the data set bench.py times:
198 resident keyframes (~13 M points) -> ~0.97 M-point local map, 128-beam scans with ~71 k DS features.

  MO:931-970   extractCloud + VoxelGrid      DS corner / surf maps array_equal the oracle's
  MO:1019/1111 nearestKSearch(5)             neighbour sets of ALL DS queries at the initial guess, every search
                                             variant of the library, against the oracle's exact kd-tree
  MO:1315-1343 scan2MapOptimization          poses <= 1e-4 m / 1e-5 rad, equal iteration counts and n_sel per
                                             iteration, for 3 scans, on every registration kernel
  determinism                                50 repeats of one registration are byte-identical
  C4 corner    Nq = 1e5, M = 4e6             gated / staged / exact vs the kd-tree
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
POS_TOL, ROT_TOL = 1e-4, 1e-5
N_SCANS = 3


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


@pytest.fixture(scope="module")
def c3(lv):
    """the bench data set + the oracle's results on it (computed once per module)"""
    import bench
    h = lv.Lvreg()
    ds = bench.make_dataset("c3", bench.SEED, lambda p, leaf: h.voxelgrid(p, leaf)[0], lambda m: None)
    h.close()
    threads = min(os.cpu_count() or 8, 32)
    mo = O.MapOptimization(O.default_params(num_threads=threads))
    for i in range(len(ds["kf_pose"])):
        mo.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i], float(i))
    ids = np.arange(len(ds["kf_pose"]), dtype=np.int32)
    mo.build_local_map(ids)
    maps = [mo.get_map(0), mo.get_map(1)]
    oracle_runs = []
    for j in range(N_SCANS):
        c, s = ds["scans"][j]
        pose, res, nc, ns = mo.register_scan(c, s, ds["guess"][j])
        oracle_runs.append((pose, res, nc, ns))
    return dict(ds=ds, ids=ids, maps=maps, oracle=oracle_runs, threads=threads)


def _handle(lv, c3, monkeypatch=None, variant=None):
    if monkeypatch is not None and variant is not None:
        monkeypatch.setenv("LVREG_REG", variant)
    h = lv.Lvreg()
    ds = c3["ds"]
    for i in range(len(ds["kf_pose"])):
        h.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i])
    return h


def test_c3_local_map_equals_oracle(lv, c3):
    h = _handle(lv, c3)
    info = h.build_local_map(c3["ids"])
    assert info.n_corner_in + info.n_surf_in > 12_000_000
    for which in (lv.CORNER, lv.SURF):
        got = h.get_local_map(which)
        assert got.shape == c3["maps"][which].shape
        assert np.array_equal(got, c3["maps"][which])
    assert len(c3["maps"][1]) > 700_000
    h.close()


def test_c3_neighbour_sets_of_all_queries(lv, c3):
    """every DS feature of a scan, transformed by the initial guess: 5-NN index sets and distances of the
    gated, staged and exact searches equal the oracle's kd-tree wherever the reference looks at them
    (5th distance inside the gate, MO:1025 / MO:1121); exact equals it everywhere."""
    h = _handle(lv, c3)
    h.build_local_map(c3["ids"])
    ds = c3["ds"]
    c, s = ds["scans"][0]
    nc, ns = h.downsample_scan(c, s)
    assert (nc, ns) == (c3["oracle"][0][2], c3["oracle"][0][3])
    for which in (lv.CORNER, lv.SURF):
        q = O.transform_cloud(h.get_scan_ds(which), ds["guess"][0], num_threads=c3["threads"])
        tree = O.KdTree(c3["maps"][which])
        oidx, od2 = tree.knn(q, 5, num_threads=c3["threads"])
        eidx, ed2 = h.knn5(which, q, lv.KNN_GRID_EXACT)
        assert np.array_equal(eidx, oidx) and np.array_equal(ed2, od2)
        inside = od2[:, 4] < 1.0
        assert inside.sum() > 1000
        for variant in (lv.KNN_GRID_GATED, lv.KNN_GRID_STAGED):
            gidx, gd2 = h.knn5(which, q, variant)
            assert np.array_equal(gidx[inside], oidx[inside]) and np.array_equal(gd2[inside], od2[inside])
            # outside the gate: exactly the neighbours closer than the gate, in order, then "none"
            k_in = (od2 < 1.0).sum(1)
            assert np.array_equal((gidx >= 0).sum(1), k_in)
            m = od2 < 1.0
            assert np.array_equal(gidx[m], oidx[m])
    h.close()


@pytest.mark.parametrize("variant", ["warm", "staged", "tpq"])
def test_c3_registration_vs_oracle(lv, c3, variant, monkeypatch):
    h = _handle(lv, c3, monkeypatch, variant)
    ds = c3["ds"]
    for j in range(N_SCANS):
        c, s = ds["scans"][j]
        pose, res, st = h.register_scan(c, s, c3["ids"], ds["guess"][j])
        opose, ores, nc, ns = c3["oracle"][j]
        assert st == lv.OK
        assert (res.n_corner_ds, res.n_surf_ds) == (nc, ns)
        assert res.iterations == ores.iterations and res.converged == ores.converged
        assert list(res.n_sel[:res.iterations]) == list(ores.n_sel[:ores.iterations])
        assert np.abs(pose[:3] - opose[:3]).max() <= ROT_TOL
        assert np.abs(pose[3:] - opose[3:]).max() <= POS_TOL
        # every intermediate pose as well (the loop has no room to drift and recover)
        gp = np.array(res.pose_iter)[:res.iterations]
        op = np.array(ores.pose_iter)[:ores.iterations]
        assert np.abs(gp[:, :3] - op[:, :3]).max() <= ROT_TOL and np.abs(gp[:, 3:] - op[:, 3:]).max() <= POS_TOL
    h.close()


def test_c3_registration_is_bit_reproducible(lv, c3):
    """static tile -> warp assignment and fixed-order fp64 sums: 50 runs of one registration give the same bytes"""
    h = _handle(lv, c3)
    ds = c3["ds"]
    c, s = ds["scans"][1]
    h.build_local_map(c3["ids"])
    h.downsample_scan(c, s)
    ref = None
    for _ in range(50):
        h.reset_lm_state()
        pose, res, st = h.scan2map(ds["guess"][1])
        assert st == lv.OK
        blob = (pose.tobytes(), bytes(res))
        if ref is None:
            ref = blob
        assert blob == ref
    h.close()


def test_c4_corner_large_map(lv):
    """C4's largest corner: 1e5 queries against a 4e6-point map, gated / staged / exact vs the kd-tree"""
    rng = np.random.default_rng(44)
    m = 4_000_000
    # points on a bundle of planes inside a 160 x 160 x 24 m box: a few points per search cell
    pts = np.empty((m, 4), np.float32)
    pts[:, 0] = rng.uniform(-80, 80, m)
    pts[:, 1] = rng.uniform(-80, 80, m)
    pts[:, 2] = np.round(rng.uniform(-12, 12, m) / 1.5) * 1.5 + rng.normal(0, 0.02, m)
    pts[:, 3] = rng.uniform(0, 255, m)
    nq = 100_000
    q = pts[rng.integers(0, m, nq)].copy()
    q[:, :3] += rng.normal(0, 0.05, (nq, 3)).astype(np.float32)
    q[: nq // 10, :3] = rng.uniform(-90, 90, (nq // 10, 3)).astype(np.float32)     # 10 % anywhere (outside the gate)
    h = lv.Lvreg()
    h.set_local_map(pts[:1000], pts)
    tree = O.KdTree(pts)
    oidx, od2 = tree.knn(q, 5, num_threads=min(os.cpu_count() or 8, 32))
    eidx, ed2 = h.knn5(lv.SURF, q, lv.KNN_GRID_EXACT)
    assert np.array_equal(eidx, oidx) and np.array_equal(ed2, od2)
    inside = od2[:, 4] < 1.0
    assert inside.sum() > nq // 2
    for variant in (lv.KNN_GRID_GATED, lv.KNN_GRID_STAGED):
        gidx, gd2 = h.knn5(lv.SURF, q, variant)
        assert np.array_equal(gidx[inside], oidx[inside]) and np.array_equal(gd2[inside], od2[inside])
    h.close()
