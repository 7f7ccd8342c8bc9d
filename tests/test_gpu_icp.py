"""GPU parity of the loop-closure row: lvreg_nn1 / lvreg_loop_find_near_keyframes /
lvreg_icp_align / lvreg_perform_loop_closure against oracle/oracle_icp.cpp on identical inputs.
Neighbour indices, squared distances and the submaps are compared bit-exactly; the ICP result
(iteration count, final transformation, fitness) within the north-star pose tolerance -- the
moments are summed in double in a different order on the two sides (1e-13 relative)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O                      # noqa: E402
from tests.synth import room_world, scan_from_world, rot_rpy   # noqa: E402

POS_TOL = 1e-4
ROT_TOL = 1e-5


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


@pytest.fixture(scope="module")
def h(lv):
    hd = lv.Lvreg()
    yield hd
    hd.close()


@pytest.fixture(scope="module")
def world():
    rng = np.random.default_rng(7)
    cw, sw = room_world(rng)
    tgt = O.voxelgrid(np.concatenate([cw, sw]), 0.4)[0]
    return dict(cw=cw, sw=sw, tgt=tgt)


def displaced(rng, tgt, n, pose, noise=0.01):
    R = rot_rpy(*pose[:3])
    t = np.asarray(pose[3:], np.float64)
    sel = tgt[rng.choice(len(tgt), n, replace=False)].copy()
    src = sel.copy()
    src[:, :3] = ((sel[:, :3].astype(np.float64) - t) @ R + rng.normal(0, noise, (n, 3))).astype(np.float32)
    return src


# ---- exact 1-NN ----------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,spread", [(1, 0.05), (5000, 0.05), (5000, 3.0), (2000, 40.0)])
def test_nn1_bit_exact(h, world, nq, spread):
    rng = np.random.default_rng(nq + int(spread * 10))
    tgt = world["tgt"]
    q = tgt[rng.integers(0, len(tgt), nq)].copy()
    q[:, :3] += rng.normal(0, spread, (nq, 3)).astype(np.float32)       # spread 40: far outside the target box
    h.icp_set_cloud(1, tgt)
    gi, gd = h.nn1(q)
    oi, od = O.nn1(tgt, q)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)


def test_nn1_ties_and_duplicates(h):
    # lattice target with duplicated points: equal distances everywhere -> the lower index must win
    g = np.arange(-4, 5, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    base = np.stack([X.ravel(), Y.ravel(), Z.ravel(), np.zeros(X.size, np.float32)], 1)
    tgt = np.concatenate([base, base[::3]]).astype(np.float32)
    rng = np.random.default_rng(3)
    q = np.concatenate([base[rng.integers(0, len(base), 500)] + np.float32(0.5), base[:200]]).astype(np.float32)
    h.icp_set_cloud(1, tgt)
    gi, gd = h.nn1(q)
    oi, od = O.nn1(tgt, q)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)


def test_nn1_max_dist_gate(h, world):
    rng = np.random.default_rng(11)
    tgt = world["tgt"]
    q = rng.uniform(-60, 60, (3000, 4)).astype(np.float32)
    h.icp_set_cloud(1, tgt)
    gi, gd = h.nn1(q, max_dist=5.0)
    oi, od = O.nn1(tgt, q)
    near = od <= 25.0
    assert np.array_equal(gi[near], oi[near])
    assert np.array_equal(gd[near], od[near])
    # beyond the gate the search may stop early: whatever it reports is farther than the gate
    far = ~near
    assert np.all((gi[far] == -1) | (gd[far] > 25.0))


# ---- ICP -----------------------------------------------------------------------------------------
def check_icp(g, o):
    assert g.converged == o.converged
    assert g.state == o.state
    assert g.iterations == o.iterations
    assert g.n_correspondences == o.n_correspondences
    assert np.max(np.abs(g.T[:3, 3] - o.T[:3, 3])) <= POS_TOL
    assert np.max(np.abs(g.T[:3, :3] - o.T[:3, :3])) <= ROT_TOL
    assert abs(g.fitness - o.fitness) <= 1e-6 * max(1.0, abs(o.fitness))


@pytest.mark.parametrize("pose", [
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    [0.02, -0.015, 0.05, 0.4, -0.3, 0.1],
    [-0.03, 0.02, -0.12, -1.2, 0.9, -0.2],
    [0.0, 0.0, 0.35, 2.5, -2.0, 0.3],
])
def test_icp_align_matches_oracle(lv, h, world, pose):
    rng = np.random.default_rng(int(abs(pose[3]) * 100) + 5)
    tgt = world["tgt"]
    src = displaced(rng, tgt, 4000, np.array(pose))
    h.icp_set_cloud(0, src)
    h.icp_set_cloud(1, tgt)
    g = h.icp_align()
    o = O.icp_align(src, tgt)
    check_icp(g, o)
    assert g.converged == 1
    # and it actually recovers the displacement
    R = rot_rpy(*pose[:3])
    assert np.max(np.abs(g.T[:3, :3] - R)) < 5e-3
    assert np.max(np.abs(g.T[:3, 3] - np.array(pose[3:]))) < 5e-2


def test_icp_iteration_cap_and_gate(lv, h, world):
    rng = np.random.default_rng(21)
    tgt = world["tgt"]
    src = displaced(rng, tgt, 3000, np.array([0.05, 0.0, 0.3, 2.0, 1.0, 0.0]))
    h.icp_set_cloud(0, src)
    h.icp_set_cloud(1, tgt)
    for kw in (dict(max_iterations=3), dict(max_iterations=1), dict(max_corr_dist=0.5), dict(transformation_epsilon=1e-3)):
        g = h.icp_align(lv.icp_default_params(**kw))
        o = O.icp_align(src, tgt, O.icp_default_params(**kw))
        check_icp(g, o)
    g = h.icp_align(lv.icp_default_params(max_iterations=3))
    assert g.state == lv.ICP_ITERATIONS and g.converged == 1 and g.iterations == 3


def test_icp_no_correspondences_and_empty(lv, h, world):
    tgt = world["tgt"]
    src = tgt[:500].copy()
    src[:, 0] += 500.0                                  # farther than max_corr_dist from everything
    h.icp_set_cloud(0, src)
    h.icp_set_cloud(1, tgt)
    g = h.icp_align()
    o = O.icp_align(src, tgt)
    assert g.converged == o.converged == 0
    assert g.state == o.state == lv.ICP_NO_CORRESPONDENCES
    assert g.iterations == o.iterations == 0
    assert abs(g.fitness - o.fitness) <= 1e-6 * o.fitness
    h.icp_set_cloud(0, np.zeros((0, 4), np.float32))
    g = h.icp_align()
    assert g.converged == 0 and g.state == lv.ICP_NO_INPUT


def test_icp_golden_vectors(lv, h):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "icp.npz"))
    h.icp_set_cloud(0, z["src"])
    h.icp_set_cloud(1, z["tgt"])
    gi, gd = h.nn1(z["src"])
    assert np.array_equal(gi, z["nn_idx"]) and np.array_equal(gd, z["nn_d2"])
    g = h.icp_align()
    assert (g.iterations, g.state, g.converged, g.n_correspondences) == \
        (int(z["iterations"]), int(z["state"]), int(z["converged"]), int(z["n_corr"]))
    assert np.max(np.abs(g.T[:3, 3] - z["final_T"][:3, 3])) <= POS_TOL
    assert np.max(np.abs(g.T[:3, :3] - z["final_T"][:3, :3])) <= ROT_TOL
    assert abs(g.fitness - float(z["fitness"])) <= 1e-9
    assert np.array_equal(lv.correct_pose(z["final_T"], z["stored_pose"]), z["corrected_pose"])


# ---- submaps + performLoopClosure ----------------------------------------------------------------
@pytest.fixture(scope="module")
def loop_sequence(lv, world):
    """a there-and-back trajectory: the last keyframe revisits the first ones with a drifted pose"""
    rng = np.random.default_rng(31)
    hd = lv.Lvreg()
    mo = O.MapOptimization()
    xs = list(np.linspace(-6, 6, 13)) + list(np.linspace(6, -6, 13))
    truth = []
    for k, x in enumerate(xs):
        pose = np.array([0.0, 0.0, 0.05 * np.sin(k), x, 0.3 * np.cos(k), 0.0], np.float32)
        c, s = scan_from_world(rng, world["cw"], world["sw"], pose, 500, 2500)
        stored = pose.copy()
        if k == len(xs) - 1:
            stored += np.array([0.01, -0.01, 0.03, 0.35, -0.25, 0.08], np.float32)     # accumulated drift
        hd.add_keyframe(c, s, stored)
        mo.add_keyframe(c, s, stored, float(2 * k))
        truth.append(pose)
    yield hd, mo, truth
    hd.close()


def test_loop_find_near_keyframes_bit_exact(loop_sequence):
    hd, mo, _ = loop_sequence
    for key, num, slot in ((25, 0, 0), (2, 25, 1), (0, 3, 1), (12, 2, 0)):
        n = hd.loop_find_near_keyframes(key, num, slot)
        ref = mo.loop_find_near_keyframes(key, num, slot)
        got = hd.icp_get_cloud(slot)
        assert n == len(ref)
        assert np.array_equal(got, ref)


def test_global_map_bit_exact(loop_sequence):
    hd, mo, _ = loop_sequence
    for ids, which, leaf in (([0, 5, 3, 25, 12], 3, 1.0), (list(range(26)), 3, 0.4), ([7, 7, 2], 1, 0.2), ([9], 2, 0.4)):
        got = hd.build_global_map(ids, which, leaf)
        ref = mo.build_global_map(ids, which, leaf)
        assert len(got) > 0 and np.array_equal(got, ref)


def test_perform_loop_closure_matches_oracle(lv, loop_sequence):
    hd, mo, truth = loop_sequence
    pair = mo.detect_loop_closure_distance(time_cur=50.0)
    assert pair is not None
    cur, pre = pair
    assert cur == 25 and pre in (0, 1, 2)
    g = hd.perform_loop_closure(cur, pre, 25)
    o = mo.perform_loop_closure(cur, pre, 25)
    assert g.status == o.status == lv.LOOP_OK
    assert g.n_source == o.n_source and g.n_target == o.n_target
    check_icp(g.icp, o.icp)
    gp, op = np.array(g.pose_from[:]), np.array(o.pose_from[:])
    assert np.max(np.abs(gp[3:] - op[3:])) <= POS_TOL
    assert np.max(np.abs(gp[:3] - op[:3])) <= ROT_TOL
    assert np.array_equal(np.array(g.pose_to[:]), np.array(o.pose_to[:]))
    assert abs(g.noise - o.noise) <= 1e-6
    # the correction moves the pose towards the truth (point-to-point ICP on voxel-sampled walls
    # stops early under PCL's 1 mm / iteration criterion: it does not remove the whole drift)
    stored = np.array([0.01, -0.01, 0.03, 0.35, -0.25, 0.08]) + truth[cur]
    assert np.linalg.norm(gp[3:] - truth[cur][3:]) < np.linalg.norm(stored[3:] - truth[cur][3:])


def test_perform_loop_closure_gates(lv, loop_sequence):
    hd, mo, _ = loop_sequence
    g = hd.perform_loop_closure(25, 0, 25, fitness_gate=1e-9)
    o = mo.perform_loop_closure(25, 0, 25, fitness_gate=1e-9)
    assert g.status == o.status == lv.LOOP_FITNESS_TOO_HIGH
    hs = lv.Lvreg()
    ms = O.MapOptimization()
    tiny = np.zeros((50, 4), np.float32)
    tiny[:, :3] = np.random.default_rng(2).uniform(-5, 5, (50, 3))
    for k in range(3):
        hs.add_keyframe(tiny, tiny, np.zeros(6, np.float32))
        ms.add_keyframe(tiny, tiny, np.zeros(6, np.float32), float(k))
    g = hs.perform_loop_closure(2, 0, 25)
    o = ms.perform_loop_closure(2, 0, 25)
    assert g.status == o.status == lv.LOOP_SUBMAP_TOO_SMALL
    assert g.n_source == o.n_source and g.n_target == o.n_target
    hs.close()


def test_next_row_apis_reject_bad_input(lv):
    """error behaviour of the loop-closure / global-map / depth / projection entry points"""
    hd = lv.Lvreg()
    with pytest.raises(lv.LvregError) as e:
        hd.icp_align()
    assert e.value.status == lv.ERR_NO_MAP
    with pytest.raises(lv.LvregError) as e:
        hd.perform_loop_closure(0, 0)
    assert e.value.status == lv.ERR_NO_KEYFRAMES
    with pytest.raises(lv.LvregError) as e:
        hd.build_global_map([0])
    assert e.value.status == lv.ERR_NO_KEYFRAMES
    pts = np.random.default_rng(0).uniform(-5, 5, (400, 4)).astype(np.float32)
    hd.add_keyframe(pts, pts, np.zeros(6, np.float32))
    with pytest.raises(lv.LvregError) as e:
        hd.build_global_map([3])
    assert e.value.status == lv.ERR_INVALID
    with pytest.raises(lv.LvregError) as e:
        hd.perform_loop_closure(0, 5)
    assert e.value.status == lv.ERR_INVALID
    with pytest.raises(lv.LvregError) as e:
        hd.icp_set_cloud(1, pts)
        hd.icp_set_cloud(0, pts)
        hd.icp_align(lv.icp_default_params(max_iterations=0))
    assert e.value.status == lv.ERR_INVALID
    # no depth cloud yet: every feature comes back without depth
    d, f3 = hd.get_depth(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32), np.array([[0.1, 0.2, 1.0]], np.float32))
    assert d[0] == -1 and f3[0, 3] == -1
    raw = lv.make_raw_cloud(pts, np.zeros(len(pts), np.uint16), np.zeros(len(pts), np.float32))
    for bad in (dict(n_scan=0), dict(n_scan=300), dict(horizon_scan=0), dict(downsample_rate=0), dict(sensor=7),
                dict(deskew=True, imu_time=np.zeros(0), imu_rot=np.zeros((0, 3)))):
        with pytest.raises(lv.LvregError) as e:
            hd.project_cloud(raw, **bad)
        assert e.value.status == lv.ERR_INVALID, bad
    # a handle keeps working after rejected calls
    assert hd.project_cloud(raw, n_scan=1, horizon_scan=1000, lidar_min_range=0.0) > 0
    hd.close()
