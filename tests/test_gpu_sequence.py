"""GPU: (1) the committed golden vectors through the C ABI, (2) the C++ mapOptimization mirror
replaying a synthetic sequence against the oracle's mapOptimization-like object (same keyframe
selection, same maps, poses within tolerance), (3) size-independent properties at BASELINE's
full C3 sizes, where the oracle would take too long for exhaustive comparison."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
POS_TOL, ROT_TOL = 1e-4, 1e-5


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


def test_golden_voxelgrid(lv):
    z = np.load(os.path.join(G, "voxelgrid.npz"))
    h = lv.Lvreg()
    out, okeys, _ = h.voxelgrid(z["pts"], float(z["leaf"]))
    assert np.array_equal(out, z["out"]) and np.array_equal(okeys, z["out_keys"])
    assert np.array_equal(h.voxel_keys(z["pts"], float(z["leaf"])), z["keys"])
    h.close()


def test_golden_registration(lv):
    z = np.load(os.path.join(G, "registration.npz"))
    h = lv.Lvreg()
    h.set_local_map(z["corner_map"], z["surf_map"])
    for variant in (lv.KNN_GRID_EXACT, lv.KNN_BRUTE):
        idx, d2 = h.knn5(lv.SURF, z["surf_queries"], variant)
        assert np.array_equal(idx, z["knn_idx"]) and np.array_equal(d2, z["knn_d2"])
    c, f, nn = h.corner_residuals(z["corner_ds"], z["guess"])
    assert np.array_equal(f, z["corner_flag"]) and np.array_equal(c, z["corner_coeff"])
    c, f, nn = h.surf_residuals(z["surf_ds"], z["guess"])
    assert np.array_equal(f, z["surf_flag"]) and np.array_equal(c, z["surf_coeff"])
    ori = np.concatenate([z["corner_ds"][z["corner_flag"] == 1], z["surf_ds"][z["surf_flag"] == 1]])
    coef = np.concatenate([z["corner_coeff"][z["corner_flag"] == 1], z["surf_coeff"][z["surf_flag"] == 1]])
    conv, pose1, AtA, Atb, x = h.lm_step(ori, coef, 0, z["guess"])
    assert np.array_equal(AtA, z["lm_AtA"]) and np.array_equal(Atb, z["lm_Atb"]) and np.array_equal(x, z["lm_x"])
    assert np.array_equal(pose1, z["lm_pose"]) and conv == int(z["lm_conv"])
    h.reset_lm_state()
    h.set_scan_ds(z["corner_ds"], z["surf_ds"])
    pose, res, st = h.scan2map(z["guess"])
    assert st == lv.OK and res.iterations == int(z["iterations"]) and res.converged == int(z["converged"])
    assert np.abs(pose[:3] - z["final_pose"][:3]).max() <= ROT_TOL
    assert np.abs(pose[3:] - z["final_pose"][3:]).max() <= POS_TOL
    assert res.n_sel[0] == int(z["n_sel"][0])
    h.close()


@pytest.mark.parametrize("variant", ["grouped", "tpq", "staged", "warm"])
def test_all_registration_kernels_agree_with_oracle(lv, variant, monkeypatch):
    monkeypatch.setenv("LVREG_REG", variant)
    z = np.load(os.path.join(G, "registration.npz"))
    h = lv.Lvreg()
    h.set_local_map(z["corner_map"], z["surf_map"])
    h.set_scan_ds(z["corner_ds"], z["surf_ds"])
    pose, res, st = h.scan2map(z["guess"])
    assert st == lv.OK and res.iterations == int(z["iterations"])
    assert list(res.n_sel[:res.iterations]) == list(z["n_sel"])
    assert np.abs(pose[:3] - z["final_pose"][:3]).max() <= ROT_TOL
    assert np.abs(pose[3:] - z["final_pose"][3:]).max() <= POS_TOL
    prof = h.iteration_profile()
    assert prof.shape == (res.iterations, 4) and (prof >= 0).all()
    h.close()


def test_mirror_sequence_replay_vs_oracle(lv):
    """C2-style: 14 MID360-like scans through the C++ mirror (keyframe rule, extractNearby,
    map rebuild on selection change, per-scan registration) vs the oracle doing the same."""
    from lidar_visual_inertial_slam_b200 import harness as H
    gen = H.Generator(H.MID360, 0x5EED0000)
    mo = H.MapOptimizationMirror()
    omo = O.MapOptimization()
    opose = None
    period = 0.4
    n_kf = 0
    for k in range(14):
        truth = gen.truth_pose(k, period, 1.0)
        corner, surf = gen.scan(truth, 100 + k, 4)
        guess = truth if k == 0 else gen.guess_pose(k, truth, 0.08, 0.02)
        st, pose, res, tim, nkf = mo.handle_scan(corner, surf, k * period, guess)
        # ---- the oracle runs the same per-scan flow (MO:316-326) ----
        t = k * period
        if omo.num_keyframes() > 0:
            ids = omo.extract_nearby(t)
            assert np.array_equal(ids, mo.selection()), "keyframe selection differs at scan %d" % k
            omo.build_local_map(ids)
        opose, ores, nc, ns = omo.register_scan(corner, surf, guess)
        if k == 0:
            assert st == lv.ERR_NO_KEYFRAMES and ores.status == 2
            assert np.array_equal(pose, guess)
        else:
            assert st == lv.OK and ores.status == 0
            assert (res.n_corner_ds, res.n_surf_ds) == (nc, ns)
            assert res.iterations == ores.iterations and res.converged == ores.converged
            assert np.abs(pose[:3] - opose[:3]).max() <= ROT_TOL and np.abs(pose[3:] - opose[3:]).max() <= POS_TOL
            assert np.abs(pose[3:] - truth[3:]).max() < 0.1
            assert tim.register_ms > 0
        # saveFrame rule: both sides add a keyframe when the pose moved > 1 m / 0.2 rad or > 1 s passed
        cds = O.voxelgrid(corner, 0.2)[0]
        sds = O.voxelgrid(surf, 0.4)[0]
        if nkf > n_kf:
            omo.add_keyframe(cds, sds, pose, t)      # same pose on both sides keeps the replays comparable
            n_kf = nkf
    assert n_kf >= 5
    mo.close()


def test_mirror_loop_closure_vs_oracle(lv):
    """A trajectory that returns to its start after > 30 s: the mirror's performLoopClosure (candidate
    search on the host, submaps + ICP + pose correction on the device) against the oracle."""
    from lidar_visual_inertial_slam_b200 import harness as H
    gen = H.Generator(H.MID360, 0x5EED0007)
    mo = H.MapOptimizationMirror()
    omo = O.MapOptimization()
    period, speed = 1.2, 2.5
    n_kf = 0
    t = 0.0
    for k in range(32):
        truth = gen.truth_pose(k, period, speed)
        corner, surf = gen.scan(truth, 500 + k, 4)
        guess = truth if k == 0 else gen.guess_pose(k, truth, 0.05, 0.01)
        t = k * period
        st, pose, res, tim, nkf = mo.handle_scan(corner, surf, t, guess)
        if nkf > n_kf:
            omo.add_keyframe(O.voxelgrid(corner, 0.2)[0], O.voxelgrid(surf, 0.4)[0], pose, t)
            n_kf = nkf
    assert n_kf >= 25
    queued, cur, pre, g = mo.perform_loop_closure()
    pair = omo.detect_loop_closure_distance(t)
    assert pair == (cur, pre) and cur == n_kf - 1 and pre < 8
    o = omo.perform_loop_closure(cur, pre, 25)
    assert g.status == o.status
    assert queued == (o.status == 0)
    assert (g.n_source, g.n_target) == (o.n_source, o.n_target)
    if o.n_source >= 300 and o.n_target >= 1000:
        assert (g.icp.iterations, g.icp.state, g.icp.converged) == (o.icp.iterations, o.icp.state, o.icp.converged)
        assert np.abs(g.icp.T[:3, 3] - o.icp.T[:3, 3]).max() <= POS_TOL
        assert np.abs(g.icp.T[:3, :3] - o.icp.T[:3, :3]).max() <= ROT_TOL
    if o.status == 0:
        assert np.abs(np.array(g.pose_from[:]) - np.array(o.pose_from[:])).max() <= POS_TOL
    # the same pair is not closed twice (loopIndexContainer, MO:636-638)
    if queued:
        assert mo.perform_loop_closure()[0] is False
    mo.close()


def test_whole_sequence_replay_binary_path(lv):
    from lidar_visual_inertial_slam_b200 import harness as H
    r = H.replay(H.MID360, 0x5EED0001, 25, device=0, period=0.2, gen_threads=4)
    assert r["registered"] == 24 and r["converged"] >= 22 and r["keyframes"] >= 4
    assert r["max_pos_err"] < 0.15 and r["max_rot_err"] < 0.01


def test_full_size_properties_c3(lv):
    """BASELINE C3 sizes (~180k-point scan, 1.5M-point map input): properties that hold regardless of size."""
    from lidar_visual_inertial_slam_b200 import harness as H
    gen = H.Generator(H.BEAM128, 0x5EED0000)
    h = lv.Lvreg()
    pose0 = np.zeros(6, np.float32)
    corner, surf = gen.scan(pose0, 7, 8)
    assert len(surf) > 150000
    ds, keys, _ = h.voxelgrid(surf, 0.4)
    assert (np.diff(keys.astype(np.int64)) > 0).all()                # one point per voxel, ascending idx
    ds2, keys2, _ = h.voxelgrid(ds, 0.4)
    assert len(ds2) == len(ds) and np.array_equal(keys2, keys)       # idempotent: centroids stay in their voxel
    assert np.array_equal(ds2, ds)
    pk = h.voxel_keys(surf, 0.4)
    assert len(np.unique(pk)) == len(ds)                             # as many outputs as distinct keys
    assert np.array_equal(np.unique(pk), keys)
    # centroid of everything is preserved by a weighted recombination (checksum of checksums)
    cnt = np.bincount(np.searchsorted(keys, pk), minlength=len(keys))
    assert np.allclose((ds[:, :3] * cnt[:, None]).sum(0) / len(surf), surf[:, :3].astype(np.float64).mean(0), atol=2e-3)
    # transform linearity / round trip at full size
    p = np.array([0.02, -0.01, 0.4, 3, -2, 0.5], np.float32)
    w = h.transform_cloud(surf, p)
    assert np.array_equal(w, O.transform_cloud(surf, p, num_threads=8))
    # kNN at full size: sorted, self-match at distance 0, exact == brute on a sample
    h.set_local_map(ds[:1000], ds)
    idx, d2 = h.knn5(lv.SURF, ds[:20000], lv.KNN_GRID_EXACT)
    assert (np.diff(d2, axis=1) >= 0).all() and (d2[:, 0] == 0).all()
    assert np.array_equal(idx[:, 0], np.arange(20000))
    bidx, bd2 = h.knn5(lv.SURF, ds[:2000], lv.KNN_BRUTE)
    assert np.array_equal(idx[:2000], bidx) and np.array_equal(d2[:2000], bd2)
    h.close()
