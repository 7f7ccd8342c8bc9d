"""GPU parity: liblvreg (through the C ABI) against the CPU oracle on identical seeded inputs.
Bit-exact for integer / index work (voxel keys, DS-map order, neighbour sets) and for the fp32
stages whose operation order is pinned (transform, centroids, d2, fits, normal equations);
the fused multi-iteration loop is compared within the north-star tolerance 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O                      # noqa: E402
from tests.synth import room_world, scan_from_world, corridor_world   # noqa: E402

POS_TOL = 1e-4      # metres  (BASELINE.json north_star)
ROT_TOL = 1e-5      # radians


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
def room():
    rng = np.random.default_rng(100)
    cw, sw = room_world(rng)
    cm = O.voxelgrid(cw, 0.2)[0]
    sm = O.voxelgrid(sw, 0.4)[0]
    truth = np.array([0.02, -0.03, 0.3, 1.0, -2.0, 0.2], np.float32)
    c, s = scan_from_world(rng, cw, sw, truth, 1500, 6000)
    cds = O.voxelgrid(c, 0.2)[0]
    sds = O.voxelgrid(s, 0.4)[0]
    guess = truth + np.array([0.02, -0.02, 0.03, 0.1, -0.08, 0.05], np.float32)
    return dict(cw=cw, sw=sw, cm=cm, sm=sm, c=c, s=s, cds=cds, sds=sds, truth=truth, guess=guess)


# ---- transform ---------------------------------------------------------------------------------
def test_pose_to_affine_host_bit_exact(lv):
    rng = np.random.default_rng(1)
    for _ in range(200):
        pose = rng.uniform(-3, 3, 6).astype(np.float32)
        assert np.array_equal(lv.pose_to_affine(pose), O.pose_to_affine(pose))


@pytest.mark.parametrize("n", [0, 1, 777, 100000])
def test_transform_cloud_bit_exact(h, n):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-80, 80, (n, 4)).astype(np.float32)
    pose = np.array([0.05, -0.1, 2.5, 10, -20, 1.5], np.float32)
    assert np.array_equal(h.transform_cloud(pts, pose), O.transform_cloud(pts, pose))


# ---- VoxelGrid ---------------------------------------------------------------------------------
@pytest.mark.parametrize("leaf,n,scale", [(0.2, 5000, 15), (0.4, 200000, 40), (0.4, 1500000, 90), (2.0, 300, 60)])
def test_voxelgrid_bit_exact(h, leaf, n, scale):
    rng = np.random.default_rng(n)
    pts = np.concatenate([rng.uniform(-scale, scale, (n, 3)) * [1, 1, 0.1], rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    ref, ref_keys, ref_okeys, _ = O.voxelgrid(pts, leaf)
    out, okeys, passthrough = h.voxelgrid(pts, leaf)
    assert not passthrough
    assert np.array_equal(okeys, ref_okeys)                 # voxel keys, ascending order
    assert np.array_equal(out, ref)                          # centroids, bit for bit
    assert np.array_equal(h.voxel_keys(pts, leaf), ref_keys)  # per-point keys


@pytest.mark.parametrize("n", [2, 3, 31, 1000, 2047, 2048, 2049, 4097])
def test_voxelgrid_single_block_path_and_its_threshold(h, n):
    """clouds of up to 2048 points take the one-block kernel, larger ones the sort pipeline: both bit-exact"""
    rng = np.random.default_rng(1000 + n)
    pts = np.concatenate([rng.uniform(-6, 6, (n, 3)) * [1, 1, 0.2], rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    pts[n // 2:, :3] = pts[: n - n // 2, :3] + np.float32(0.01)        # many shared voxels, order-dependent centroids
    for leaf in (0.2, 2.0):
        ref, ref_keys, ref_okeys, _ = O.voxelgrid(pts, leaf)
        out, okeys, passthrough = h.voxelgrid(pts, leaf)
        assert not passthrough
        assert np.array_equal(okeys, ref_okeys) and np.array_equal(out, ref)
        assert np.array_equal(h.voxel_keys(pts, leaf), ref_keys)


def test_voxelgrid_clustered_many_points_per_voxel(h):
    rng = np.random.default_rng(5)
    centers = rng.uniform(-10, 10, (50, 3))
    pts = (centers[rng.integers(0, 50, 300000)] + rng.normal(0, 0.15, (300000, 3)))
    pts = np.concatenate([pts, rng.uniform(0, 255, (300000, 1))], 1).astype(np.float32)
    ref = O.voxelgrid(pts, 0.2)[0]
    out, _, _ = h.voxelgrid(pts, 0.2)
    assert np.array_equal(out, ref)


def test_voxelgrid_pcl_layout_roundtrip(h, lv):
    from lidar_visual_inertial_slam_b200.binding import to_pcl_layout
    rng = np.random.default_rng(6)
    pts = rng.uniform(-20, 20, (20000, 4)).astype(np.float32)
    ref = O.voxelgrid(pts, 0.4)[0]
    out, _, _ = h.voxelgrid(to_pcl_layout(pts), 0.4, pcl_layout_out=True)
    assert out.shape[1] == 8
    assert np.array_equal(out[:, :3], ref[:, :3]) and np.array_equal(out[:, 4], ref[:, 3])
    assert (out[:, 3] == 1.0).all() and not out[:, 5:].any()


def test_voxelgrid_edge_cases(h):
    out, keys, _ = h.voxelgrid(np.zeros((0, 4), np.float32), 0.2)
    assert len(out) == 0
    p = np.array([[1.5, -2.5, 3.0, 9.0]], np.float32)
    out, keys, _ = h.voxelgrid(p, 0.4)
    assert np.array_equal(out, p) and keys[0] == 0
    # identical points collapse to one voxel
    out, _, _ = h.voxelgrid(np.tile(p, (1000, 1)), 0.2)
    assert np.array_equal(out, O.voxelgrid(np.tile(p, (1000, 1)), 0.2)[0]) and len(out) == 1
    # PCL's overflow rule: input returned unchanged
    big = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2], [1, 1, 1, 3]], np.float32)
    out, _, passthrough = h.voxelgrid(big, 0.2)
    assert passthrough and np.array_equal(out, big)


def test_reserve_presizes_without_changing_results(lv, room):
    """lvreg_reserve: same results, and it rejects sizes out of range"""
    a = lv.Lvreg()
    b = lv.Lvreg()
    b.reserve(map_points_corner=50000, map_points_surf=200000, scan_points_corner=4000, scan_points_surf=20000,
              max_grid_cells=1 << 20)
    for hd in (a, b):
        hd.add_keyframe(room["cw"], room["sw"], np.zeros(6, np.float32))
    pa, ra, sa = a.register_scan(room["c"], room["s"], [0], room["guess"])
    pb, rb, sb = b.register_scan(room["c"], room["s"], [0], room["guess"])
    assert sa == sb == lv.OK and np.array_equal(pa, pb) and ra.iterations == rb.iterations
    assert np.array_equal(a.get_local_map(lv.SURF), b.get_local_map(lv.SURF))
    with pytest.raises(lv.LvregError):
        b.reserve(max_grid_cells=1 << 30)
    a.close()
    b.close()


# ---- 5-NN ------------------------------------------------------------------------------------
def _set_map(h, m):
    h.set_local_map(m, m)


@pytest.mark.parametrize("variant", ["gated", "exact", "brute"])
def test_knn5_bit_exact_vs_oracle(h, lv, room, variant):
    rng = np.random.default_rng(7)
    mp = room["sm"]
    _set_map(h, mp)
    near = mp[rng.choice(len(mp), 4000)] + np.r_[rng.normal(0, 0.05, 3), 0].astype(np.float32)
    near = (mp[rng.choice(len(mp), 4000)][:, :3] + rng.normal(0, 0.05, (4000, 3))).astype(np.float32)
    far = rng.uniform(-15, 15, (600, 3)).astype(np.float32)
    q = np.concatenate([np.concatenate([near, far]), np.zeros((4600, 1), np.float32)], 1)
    ridx, rd2 = O.knn5_brute(mp, q)
    v = dict(gated=lv.KNN_GRID_GATED, exact=lv.KNN_GRID_EXACT, brute=lv.KNN_BRUTE)[variant]
    idx, d2 = h.knn5(lv.SURF, q, v)
    if variant == "gated":
        inside = rd2[:, 4] < 1.0
        assert inside.sum() > 3000
        assert np.array_equal(idx[inside], ridx[inside]) and np.array_equal(d2[inside], rd2[inside])
        assert (d2[~inside, 4] >= 1.0).all()          # unresolved rows are rejected by the gate anyway
    else:
        assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)


@pytest.mark.parametrize("variant", ["exact", "brute"])
def test_knn5_exact_ties_and_outside_queries(h, lv, variant):
    g = np.arange(-6, 7, dtype=np.float32) * 0.5
    mp = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    mp = np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1)
    _set_map(h, mp)
    q = np.array([[0, 0, 0, 0], [0.25, 0.25, 0.25, 0], [1, -1.5, 0.5, 0], [40, 3, -2, 0], [-9, -9, -9, 0],
                  [3.01, 3.01, 3.01, 0], [0, 0, 55, 0]], np.float32)
    ridx, rd2 = O.knn5_brute(mp, q)
    v = dict(exact=lv.KNN_GRID_EXACT, brute=lv.KNN_BRUTE)[variant]
    idx, d2 = h.knn5(lv.SURF, q, v)
    assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)


@pytest.mark.parametrize("m", [0, 1, 3, 5])
def test_knn5_tiny_maps(h, lv, m):
    rng = np.random.default_rng(m)
    mp = rng.uniform(-1, 1, (m, 4)).astype(np.float32)
    _set_map(h, mp)
    q = rng.uniform(-1, 1, (40, 4)).astype(np.float32)
    ridx, rd2 = O.knn5_brute(mp, q)
    for v in (lv.KNN_GRID_EXACT, lv.KNN_BRUTE):
        idx, d2 = h.knn5(lv.SURF, q, v)
        assert np.array_equal(idx, ridx) and np.array_equal(d2, rd2)


def test_knn5_large_map_exact(h, lv):
    rng = np.random.default_rng(8)
    m = 400000
    mp = np.concatenate([rng.uniform(-150, 150, (m, 2)), rng.uniform(-3, 12, (m, 1)), np.zeros((m, 1))], 1).astype(np.float32)
    _set_map(h, mp)
    q = mp[rng.choice(m, 20000)].copy()
    q[:, :3] += rng.normal(0, 0.3, (20000, 3)).astype(np.float32)
    tidx, td2 = O.KdTree(mp).knn(q, 5)
    idx, d2 = h.knn5(lv.SURF, q, lv.KNN_GRID_EXACT)
    assert np.array_equal(idx, tidx) and np.array_equal(d2, td2)


# ---- residuals ---------------------------------------------------------------------------------
def _check_residuals(coeff, flag, nn, rcoeff, rflag, rnn):
    mism = np.flatnonzero(flag != rflag)
    assert len(mism) <= max(1, len(flag) // 1000), "accept-flag mismatches: %s" % mism[:20]
    both = (flag == 1) & (rflag == 1)
    assert both.sum() > 0
    assert np.array_equal(nn[both], rnn[both])                      # neighbour sets bit-exact
    assert np.allclose(coeff[both], rcoeff[both], rtol=1e-4, atol=1e-5)
    return float((coeff[both] == rcoeff[both]).all(axis=1).mean())


def test_corner_residuals_vs_oracle(h, room):
    h.set_local_map(room["cm"], room["sm"])
    coeff, flag, nn = h.corner_residuals(room["cds"], room["guess"])
    rcoeff, rflag, rnn = O.corner_residuals(room["cm"], room["cds"], room["guess"])
    assert rflag.sum() > 300
    exact = _check_residuals(coeff, flag, nn, rcoeff, rflag, rnn)
    assert exact > 0.999, "corner coefficients bit-exact fraction %.5f" % exact


def test_surf_residuals_vs_oracle(h, room):
    h.set_local_map(room["cm"], room["sm"])
    coeff, flag, nn = h.surf_residuals(room["sds"], room["guess"])
    rcoeff, rflag, rnn = O.surf_residuals(room["sm"], room["sds"], room["guess"])
    assert rflag.sum() > 2000
    exact = _check_residuals(coeff, flag, nn, rcoeff, rflag, rnn)
    assert exact > 0.999, "surf coefficients bit-exact fraction %.5f" % exact


# ---- LM step -----------------------------------------------------------------------------------
def test_lm_step_vs_oracle(h, room):
    rcoeff_c, rflag_c, _ = O.corner_residuals(room["cm"], room["cds"], room["guess"])
    rcoeff_s, rflag_s, _ = O.surf_residuals(room["sm"], room["sds"], room["guess"])
    ori = np.concatenate([room["cds"][rflag_c == 1], room["sds"][rflag_s == 1]])
    coeff = np.concatenate([rcoeff_c[rflag_c == 1], rcoeff_s[rflag_s == 1]])
    rconv, rpose, rAtA, rAtb, rx, _ = O.lm_step(ori, coeff, 0, room["guess"])
    h.reset_lm_state()
    conv, pose, AtA, Atb, x = h.lm_step(ori, coeff, 0, room["guess"])
    assert np.array_equal(AtA, rAtA) and np.array_equal(Atb, rAtb)   # fp64 accumulation, fp32 result
    assert np.array_equal(x, rx) and np.array_equal(pose, rpose) and conv == rconv
    # fewer than 50 rows: no update (MO:1209-1212)
    conv, pose, _, _, _ = h.lm_step(ori[:49], coeff[:49], 0, room["guess"])
    assert conv == 0 and np.array_equal(pose, room["guess"])


# ---- full registration -----------------------------------------------------------------------------
def _pose_close(a, b):
    return np.abs(a[:3] - b[:3]).max() <= ROT_TOL and np.abs(a[3:] - b[3:]).max() <= POS_TOL


def test_scan2map_vs_oracle(h, lv, room):
    rpose, rres, _ = O.scan2map(room["cm"], room["sm"], room["cds"], room["sds"], room["guess"])
    h.reset_lm_state()
    h.set_local_map(room["cm"], room["sm"])
    h.set_scan_ds(room["cds"], room["sds"])
    pose, res, st = h.scan2map(room["guess"])
    assert st == lv.OK
    assert res.iterations == rres.iterations and res.converged == rres.converged == 1
    assert res.degenerate == rres.degenerate == 0
    assert res.n_sel[0] == rres.n_sel[0]                     # iteration 0 sees the identical problem
    for i in range(res.iterations):
        assert abs(res.n_sel[i] - rres.n_sel[i]) <= 3
        assert _pose_close(np.array(res.pose_iter[i]), np.array(rres.pose_iter[i]))
    assert _pose_close(pose, rpose)
    assert np.abs(pose - room["truth"])[3:].max() < 0.01       # and it actually registers


def test_scan2map_deterministic(h, room):
    h.set_local_map(room["cm"], room["sm"])
    h.set_scan_ds(room["cds"], room["sds"])
    outs = []
    for _ in range(3):
        h.reset_lm_state()
        pose, res, _ = h.scan2map(room["guess"])
        outs.append((pose.tobytes(), bytes(res)))
    assert outs[0] == outs[1] == outs[2]


def test_scan2map_soft_failures(h, lv, room):
    h.set_local_map(room["cm"], room["sm"])
    # Nc <= edgeFeatureMinValidNum (MO:1320): warn, pose untouched
    h.set_scan_ds(room["cds"][:10], room["sds"])
    pose, res, st = h.scan2map(room["guess"])
    assert st == lv.ERR_NOT_ENOUGH_FEATURES and np.array_equal(pose, room["guess"])
    h.set_scan_ds(room["cds"], room["sds"][:100])
    pose, res, st = h.scan2map(room["guess"])
    assert st == lv.ERR_NOT_ENOUGH_FEATURES and np.array_equal(pose, room["guess"])
    # fewer than 50 correspondences: 20 iterations without an update (MO:1209-1212)
    far = room["guess"].copy()
    far[3] += 500.0
    h.set_scan_ds(room["cds"], room["sds"])
    pose, res, st = h.scan2map(far)
    rpose, rres, _ = O.scan2map(room["cm"], room["sm"], room["cds"], room["sds"], far)
    assert st == lv.OK and np.array_equal(pose, far) and np.array_equal(rpose, far)
    assert res.iterations == rres.iterations == 20 and res.converged == rres.converged == 0
    # no map at all
    fresh = lv.Lvreg()
    fresh.set_scan_ds(room["cds"], room["sds"])
    pose, res, st = fresh.scan2map(room["guess"])
    assert st == lv.ERR_NO_KEYFRAMES and np.array_equal(pose, room["guess"])
    fresh.close()


@pytest.mark.parametrize("quirks", [1, 0])
def test_degenerate_corridor_matches_oracle(lv, quirks):
    rng = np.random.default_rng(300)
    cw, sw = corridor_world(rng, noise=0.003)
    cm = O.voxelgrid(cw, 0.2)[0]
    sm = O.voxelgrid(sw, 0.4)[0]
    truth = np.array([0.0, 0.0, 0.02, 0.5, 0.1, 0.0], np.float32)
    c, s = scan_from_world(rng, cw, sw, truth, 1500, 6000)
    cds = O.voxelgrid(c, 0.2)[0]
    sds = O.voxelgrid(s, 0.4)[0]
    guess = truth + np.array([0.01, -0.01, 0.02, 0.3, 0.05, -0.04], np.float32)
    # lambda_min of JtJ is ~1e2 here (x is unobservable up to plane-fit noise); the threshold is
    # raised so that exactly that direction is classified degenerate
    rpose, rres, _ = O.scan2map(cm, sm, cds, sds, guess, params=O.default_params(reference_quirks=quirks, degeneracy_eig=1000.0))
    assert rres.degenerate == 1
    hd = lv.Lvreg(lv.default_params(reference_quirks=quirks, degeneracy_eig=1000.0))
    hd.set_local_map(cm, sm)
    hd.set_scan_ds(cds, sds)
    pose, res, st = hd.scan2map(guess)
    assert st == lv.OK and res.degenerate == 1 and hd.get_degenerate() == 1
    assert res.iterations == rres.iterations and res.converged == rres.converged
    if quirks:
        assert res.iterations == 2       # one projected update, then the zero matP "converges" 
    assert np.abs(pose[:3] - rpose[:3]).max() <= 5e-5 and np.abs(pose[3:] - rpose[3:]).max() <= 5e-4
    hd.close()


# ---- keyframes -> local map -> per-scan registration (C2-style) ---------------------------------
def test_local_map_build_and_register_scan_vs_oracle(lv, room):
    from lidar_visual_inertial_slam_b200.binding import to_pcl_layout
    rng = np.random.default_rng(400)
    hd = lv.Lvreg()
    mo = O.MapOptimization()
    poses = []
    for k in range(6):
        pose = np.array([0.01 * k, -0.01 * k, 0.1 * k, 0.8 * k, -0.3 * k, 0.05 * k], np.float32)
        c, s = scan_from_world(rng, room["cw"], room["sw"], pose, 1500, 6000)
        cds = O.voxelgrid(c, 0.2)[0]
        sds = O.voxelgrid(s, 0.4)[0]
        assert hd.add_keyframe(to_pcl_layout(cds), to_pcl_layout(sds), pose) == k
        mo.add_keyframe(cds, sds, pose, float(k))
        poses.append(pose)
    ids = mo.extract_nearby(6.0)
    assert len(ids) >= 6                               # recent keyframes are appended again (duplicates)
    mo.build_local_map(ids)
    info = hd.build_local_map(ids)
    for which in (lv.CORNER, lv.SURF):
        assert np.array_equal(hd.get_local_map(which), mo.get_map(which))     # bit-exact DS map, same order
    assert info.n_corner_ds == len(mo.get_map(0)) and info.n_surf_ds == len(mo.get_map(1))

    truth = np.array([0.03, -0.02, 0.65, 5.0, -1.9, 0.3], np.float32)
    c, s = scan_from_world(rng, room["cw"], room["sw"], truth, 1500, 6000)
    guess = truth + np.array([-0.02, 0.01, 0.02, -0.1, 0.06, 0.03], np.float32)
    rpose, rres, rnc, rns = mo.register_scan(c, s, guess)
    pose, res, st = hd.register_scan(to_pcl_layout(c), to_pcl_layout(s), ids, guess)
    assert st == lv.OK and (res.n_corner_ds, res.n_surf_ds) == (rnc, rns)
    assert np.array_equal(hd.get_scan_ds(lv.CORNER), O.voxelgrid(c, 0.2)[0])
    assert res.iterations == rres.iterations and res.converged == rres.converged
    assert _pose_close(pose, rpose)
    t = hd.timings()
    assert t.kernel_launches >= 4 and t.register_ms > 0 and t.map_build_ms > 0      # a handful of fused launches
    # after a pose correction (loop closure) the map must be rebuilt
    hd.update_keyframe_poses(np.array(poses) + np.float32(0.001))
    _, _, st = hd.scan2map(guess)
    assert st == lv.ERR_NO_MAP
    hd.close()


def test_transform_update(lv, h):
    p = np.array([0.3, -0.2, 1.0, 1, 2, 3], np.float32)
    assert np.array_equal(h.transform_update(p), O.transform_update(p))
    a = h.transform_update(p, True, 0.5, 0.1)
    b = O.transform_update(p, True, 0.5, 0.1, 0.01)
    assert np.allclose(a, b, atol=1e-6)


def test_transform_update_order_inside_scan2map(lv, room):
    """MO:1345-1372: the IMU blend comes first, the clamp once after it.  With a tight rotation tolerance the two
    orders differ (clamp -> blend -> clamp pulls the clamped angle towards the IMU value; blend -> clamp clamps the
    blended angle): the registration call must give the reference's."""
    prm = lv.default_params()
    prm.rotation_tolerance = 0.015
    prm.z_tolerance = 0.15
    prm.imu_rpy_weight = 0.5
    hd = lv.Lvreg(prm)
    hd.set_local_map(room["cm"], room["sm"])
    hd.set_scan_ds(room["cds"], room["sds"])
    imu_roll, imu_pitch = -0.02, 0.03          # pulls the (out-of-tolerance) LM angles back inside the tolerance
    hd.set_imu_prior(True, imu_roll, imu_pitch)
    pose, res, st = hd.scan2map(room["guess"])
    assert st == lv.OK
    raw = np.array(res.pose_iter)[res.iterations - 1].astype(np.float32)       # transformTobeMapped before transformUpdate
    op = O.default_params(rotation_tolerance=0.015, z_tolerance=0.15)
    want = O.transform_update(raw, True, imu_roll, imu_pitch, 0.5, op)
    assert np.allclose(pose, want, atol=1e-6)
    # the other order would have given something else
    wrong = O.transform_update(O.transform_update(raw, False, 0, 0, 0.5, op), True, imu_roll, imu_pitch, 0.5, op)
    assert np.abs(raw[:2]).max() > 0.015 and np.abs(wrong - want).max() > 1e-3
    # without a prior: clamps only
    hd.set_imu_prior(False)
    hd.reset_lm_state()
    pose2, _, _ = hd.scan2map(room["guess"])
    assert np.allclose(pose2, O.transform_update(raw, False, 0, 0, 0.5, op), atol=1e-6)
    hd.close()


# ---- ragged / degenerate inputs ---------------------------------------------------------------------
def test_ragged_inputs_through_the_scan_path(lv, room):
    hd = lv.Lvreg()
    empty = np.zeros((0, 4), np.float32)
    # keyframes with an empty corner / surf cloud, and a selection that repeats an id
    hd.add_keyframe(room["cm"], empty, np.zeros(6, np.float32))
    hd.add_keyframe(empty, room["sm"], np.zeros(6, np.float32))
    hd.add_keyframe(room["cm"], room["sm"], np.array([0, 0, 0, 0.05, 0, 0], np.float32))
    mo = O.MapOptimization()
    mo.add_keyframe(room["cm"], empty, np.zeros(6, np.float32), 0.0)
    mo.add_keyframe(empty, room["sm"], np.zeros(6, np.float32), 1.0)
    mo.add_keyframe(room["cm"], room["sm"], np.array([0, 0, 0, 0.05, 0, 0], np.float32), 2.0)
    ids = np.array([2, 0, 1, 2], np.int32)            # duplicates are concatenated twice (MO:919-926)
    info = hd.build_local_map(ids)
    mo.build_local_map(ids)
    assert info.n_corner_in == 3 * len(room["cm"]) and info.n_surf_in == 3 * len(room["sm"])
    for which in (lv.CORNER, lv.SURF):
        assert np.array_equal(hd.get_local_map(which), mo.get_map(which))
    # empty incoming scan: "not enough features", pose untouched, nothing crashes
    pose, res, st = hd.register_scan(empty, empty, None, room["guess"])
    assert st == lv.ERR_NOT_ENOUGH_FEATURES and np.array_equal(pose, room["guess"])
    assert (res.n_corner_ds, res.n_surf_ds) == (0, 0)
    # an empty selection gives empty maps; registration then finds no correspondences
    hd.build_local_map(np.zeros(0, np.int32))
    assert len(hd.get_local_map(lv.SURF)) == 0
    pose, res, st = hd.register_scan(room["c"], room["s"], None, room["guess"])
    assert st == lv.OK and res.converged == 0 and res.iterations == 20 and np.array_equal(pose, room["guess"])
    # invalid ids are rejected
    with pytest.raises(lv.LvregError):
        hd.build_local_map(np.array([7], np.int32))
    hd.close()


def test_far_from_origin_coordinates(lv, room):
    """maps a few km from the origin: voxel keys / neighbour sets stay bit-exact (fp32 rounding of large coordinates)"""
    off = np.array([4321.5, -2876.25, 37.0, 0], np.float32)
    cm = (room["cm"] + off).astype(np.float32)
    sm = (room["sm"] + off).astype(np.float32)
    hd = lv.Lvreg()
    out, okeys, _ = hd.voxelgrid(sm, 0.4)
    ref, _, rkeys, _ = O.voxelgrid(sm, 0.4)
    assert np.array_equal(out, ref) and np.array_equal(okeys, rkeys)
    hd.set_local_map(cm, sm)
    guess = room["guess"].copy()
    guess[3:] += off[:3]
    coeff, flag, nn = hd.surf_residuals(room["sds"], guess)
    rcoeff, rflag, rnn = O.surf_residuals(sm, room["sds"], guess)
    assert np.array_equal(flag, rflag) and rflag.sum() > 1000
    assert np.array_equal(nn[rflag == 1], rnn[rflag == 1]) and np.array_equal(coeff, rcoeff)
    hd.close()


def test_two_handles_are_independent(lv, room):
    a, b = lv.Lvreg(), lv.Lvreg()
    a.set_local_map(room["cm"], room["sm"])
    b.set_local_map(room["cm"][:100], room["sm"][:100])
    a.set_scan_ds(room["cds"], room["sds"])
    b.set_scan_ds(room["cds"], room["sds"])
    pa, ra, sa = a.scan2map(room["guess"])
    pb, rb, sb = b.scan2map(room["guess"])
    pa2, ra2, _ = a.scan2map(room["guess"])
    assert ra.converged == 1 and np.array_equal(pa, pa2)             # b's work did not disturb a
    rpb, rrb, _ = O.scan2map(room["cm"][:100], room["sm"][:100], room["cds"], room["sds"], room["guess"])
    assert rb.iterations == rrb.iterations and rb.converged == rrb.converged
    assert np.abs(pb - rpb).max() <= 1e-4 and not np.array_equal(pa, pb)
    a.close()
    b.close()


def test_hashed_cell_directory_sparse_kilometre_map(lv):
    """A sparse map spread over > 2^26 search cells (two blocks of structure 1.6 km apart): the directory becomes
    a hash of the occupied cells, the cell keeps its 1 m edge (no coarsening), and every search variant and the
    registration kernels still equal the oracle."""
    rng = np.random.default_rng(77)
    cw, sw = room_world(rng, n_surf=20000, n_corner=4000)
    off = np.array([1600.0, 1500.0, 40.0, 0.0], np.float32)
    surf_map = np.concatenate([sw, sw + off]).astype(np.float32)
    corner_map = np.concatenate([cw, cw + off]).astype(np.float32)
    h = lv.Lvreg()
    info = h.set_local_map(corner_map, surf_map)
    dims = np.array(info.grid_dims[1][:], np.int64)
    assert dims.prod() > (1 << 26)                          # a dense directory would need > 2^26 cells ...
    assert abs(info.grid_cell[1] - 1.0078125) < 1e-6       # ... and the cell was NOT coarsened
    q = np.concatenate([surf_map[rng.integers(0, len(surf_map), 3000)],
                        rng.uniform(-20, 1650, (300, 4)).astype(np.float32)]).astype(np.float32)
    q[:3000, :3] += rng.normal(0, 0.05, (3000, 3)).astype(np.float32)
    tree = O.KdTree(surf_map)
    oidx, od2 = tree.knn(q, 5)
    eidx, ed2 = h.knn5(lv.SURF, q, lv.KNN_GRID_EXACT)
    assert np.array_equal(eidx, oidx) and np.array_equal(ed2, od2)
    inside = od2[:, 4] < 1.0
    assert inside.sum() > 2000
    for variant in (lv.KNN_GRID_GATED, lv.KNN_GRID_STAGED):
        gidx, gd2 = h.knn5(lv.SURF, q, variant)
        assert np.array_equal(gidx[inside], oidx[inside]) and np.array_equal(gd2[inside], od2[inside])
    # registration in the far block, every kernel variant
    truth = np.array([0.01, -0.02, 0.2, 1600.5, 1499.0, 40.1], np.float32)
    c, s = scan_from_world(rng, cw + off, sw + off, truth, 1200, 4000)
    guess = truth + np.array([0.01, -0.01, 0.02, 0.08, -0.05, 0.03], np.float32)
    cds, sds = O.voxelgrid(c, 0.2)[0], O.voxelgrid(s, 0.4)[0]
    opose, ores = O.scan2map(corner_map, surf_map, cds, sds, guess)[:2]
    h.close()
    import os
    for variant in ("warm", "tpq", "grouped", "staged"):
        os.environ["LVREG_REG"] = variant
        try:
            h = lv.Lvreg()
            h.set_local_map(corner_map, surf_map)
            h.set_scan_ds(cds, sds)
            pose, res, st = h.scan2map(guess)
            assert st == lv.OK and res.iterations == ores.iterations
            assert np.abs(pose[:3] - opose[:3]).max() <= 1e-5 and np.abs(pose[3:] - opose[3:]).max() <= 1e-4
            h.close()
        finally:
            os.environ.pop("LVREG_REG", None)


def test_failed_call_leaves_a_consistent_handle(lv, room):
    """a call that fails half way (bad keyframe id, malformed second cloud) must not leave counts and data of
    different scans / maps behind: the next registration reports 'no map' / 'not enough features'"""
    hd = lv.Lvreg()
    hd.add_keyframe(room["cm"], room["sm"], np.zeros(6, np.float32))
    hd.build_local_map([0])
    pose, res, st = hd.register_scan(room["c"], room["s"], None, room["guess"])
    assert st == lv.OK
    with pytest.raises(lv.LvregError):
        hd.register_scan(room["c"], room["s"], np.array([5], np.int32), room["guess"])      # id out of range
    _, _, st = hd.scan2map(room["guess"])
    assert st in (lv.ERR_NO_MAP, lv.ERR_NOT_ENOUGH_FEATURES)
    # malformed second cloud in set_scan_ds: the first upload already happened, the counts must not survive
    hd.build_local_map([0])
    hd.set_scan_ds(room["cds"], room["sds"])
    from lidar_visual_inertial_slam_b200.binding import Cloud, _cloud
    import ctypes as C
    good, keep = _cloud(room["cds"])
    bad = Cloud()
    bad.data = None
    bad.n = 10
    bad.stride = 16
    bad.intensity_offset = 12
    assert hd.L.lvreg_set_scan_ds(hd.h, C.byref(good), C.byref(bad)) == lv.ERR_INVALID
    _, _, st = hd.scan2map(room["guess"])
    assert st == lv.ERR_NOT_ENOUGH_FEATURES
    # and the handle still works
    pose2, res2, st = hd.register_scan(room["c"], room["s"], np.array([0], np.int32), room["guess"])
    assert st == lv.OK and np.array_equal(pose2, pose)
    hd.close()


def test_repeated_sessions_do_not_grow_device_memory(lv, room):
    """lvreg_clear_keyframes hands the keyframe arena back: replaying session after session on one handle must
    not allocate more and more device memory (keyframe clouds of changing sizes used to abandon their blocks)"""
    import torch
    hd = lv.Lvreg()
    rng = np.random.default_rng(8)

    def session(scale):
        for k in range(24):
            nc = int(len(room["cm"]) * scale * rng.uniform(0.5, 1.0))
            ns = int(len(room["sm"]) * scale * rng.uniform(0.5, 1.0))
            hd.add_keyframe(room["cm"][:nc], room["sm"][:ns], np.array([0, 0, 0, 0.01 * k, 0, 0], np.float32))
        hd.build_local_map(np.arange(24, dtype=np.int32))
        hd.clear_keyframes()

    session(1.0)
    session(1.0)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for i in range(12):
        session(0.3 + 0.7 * ((i * 7) % 10) / 10.0)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (64 << 20), "device memory grew by %.1f MB over 12 sessions" % ((free0 - free1) / 1048576.0)
    hd.close()
