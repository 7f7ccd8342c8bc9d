"""Pins of the CPU oracle against the third-party code the reference calls (where that code is
available in this image: the cv2 wheel = OpenCV 4.13 built without Eigen -> Jacobi cv::eigen) and
against independent numpy / scipy computations.  CPU only."""
import numpy as np
import pytest

from oracle import pyoracle as O

cv2 = pytest.importorskip("cv2")


def _spd(rng, n, scale):
    B = rng.standard_normal((n + 2, n)).astype(np.float32) * np.float32(scale)
    A = (B.T @ B).astype(np.float32)
    return (np.triu(A) + np.triu(A, 1).T).astype(np.float32)


@pytest.mark.parametrize("n", [3, 6])
def test_jacobi_eigen_bit_exact_vs_cv2(n):
    # cv::eigen at MO:1050 (3x3) and MO:1268 (6x6)
    rng = np.random.default_rng(10 + n)
    for _ in range(800):
        A = _spd(rng, n, rng.choice([0.01, 1.0, 30.0]))
        _, w, v = cv2.eigen(A)
        w2, v2 = O.jacobi_eigen(A)
        assert np.array_equal(w.ravel(), w2)
        assert np.array_equal(v, v2)


def test_jacobi_eigen_line_covariances_bit_exact_vs_cv2():
    rng = np.random.default_rng(3)
    for _ in range(800):
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        pts = (np.outer(rng.uniform(-1, 1, 5), d) + rng.normal(0, 0.02, (5, 3))).astype(np.float32)
        c = pts - pts.mean(0)
        A = (c.T @ c / 5).astype(np.float32)
        A = np.triu(A) + np.triu(A, 1).T
        _, w, v = cv2.eigen(A)
        w2, v2 = O.jacobi_eigen(A)
        assert np.array_equal(w.ravel(), w2) and np.array_equal(v, v2)


def test_qr_solve_bit_exact_vs_cv2():
    # cv::solve(matAtA, matAtB, matX, DECOMP_QR) MO:1260
    rng = np.random.default_rng(4)
    for _ in range(800):
        J = rng.standard_normal((200, 6)).astype(np.float32) * np.array([5, 5, 5, 1, 1, 1], np.float32)
        A = (J.T @ J).astype(np.float32)
        b = rng.standard_normal((6, 1)).astype(np.float32)
        _, x = cv2.solve(A, b, flags=cv2.DECOMP_QR)
        ok, x2 = O.qr_solve(A, b)
        assert ok == 1 and np.array_equal(x, x2)


def test_qr_solve_singular_behaves_like_cv2():
    # exactly singular -> NaNs propagate (0/0 in the reflector), tiny pivot -> "false" and x = 0
    b = np.ones((6, 1), np.float32)
    A = np.zeros((6, 6), np.float32)
    A[0, 0] = 1
    okc, x = cv2.solve(A, b, flags=cv2.DECOMP_QR)
    ok, x2 = O.qr_solve(A, b)
    assert bool(okc) == bool(ok) and np.array_equal(x, x2, equal_nan=True)
    A = np.diag([1, 1, 1, 1, 1, 1e-7]).astype(np.float32)
    okc, x = cv2.solve(A, b, flags=cv2.DECOMP_QR)
    ok, x2 = O.qr_solve(A, b)
    assert (not okc) and ok == 0 and not x.any() and not x2.any()


def test_lu_solve_bit_exact_vs_cv2():
    # matV.inv() * matV2 (MO:1283) lowers to cv::solve(matV, matV2, DECOMP_LU)
    rng = np.random.default_rng(5)
    for _ in range(800):
        A = rng.standard_normal((6, 6)).astype(np.float32)
        B = rng.standard_normal((6, 6)).astype(np.float32)
        _, x = cv2.solve(A, B, flags=cv2.DECOMP_LU)
        ok, x2 = O.lu_solve(A, B)
        assert ok == 1 and np.array_equal(x, x2)


@pytest.mark.parametrize("n", [57, 300, 5000, 12000, 120000])
def test_normal_equations_bit_exact_vs_cv2_gemm(n):
    # cv::transpose(matA, matAt); matAtA = matAt*matA; matAtB = matAt*matB  MO:1257-1259
    rng = np.random.default_rng(n)
    J = (rng.standard_normal((n, 6)) * np.array([5, 5, 5, 1, 1, 1])).astype(np.float32)
    b = rng.standard_normal((n, 1)).astype(np.float32)
    At = np.ascontiguousarray(J.T)
    AtA, Atb = O.normal_equations(J, b)
    assert np.array_equal(cv2.gemm(At, J, 1, None, 0), AtA)
    assert np.array_equal(cv2.gemm(At, b, 1, None, 0).ravel(), Atb)


def test_small_gemm_bit_exact_vs_cv2():
    rng = np.random.default_rng(6)
    for _ in range(300):
        P = rng.standard_normal((6, 6)).astype(np.float32)
        x = rng.standard_normal((6, 1)).astype(np.float32)
        assert np.array_equal(cv2.gemm(P, x, 1, None, 0), O.gemm(P, x))


def test_plane_fit_matches_lstsq():
    # Eigen colPivHouseholderQr().solve MO:1128: Eigen is absent here -> tolerance check only
    rng = np.random.default_rng(7)
    for _ in range(500):
        n = rng.standard_normal(3)
        n /= np.linalg.norm(n)
        c = rng.uniform(-30, 30, 3) + n * rng.uniform(2, 20)
        basis = np.linalg.svd(n[None])[2][1:]
        pts = c + rng.uniform(-0.6, 0.6, (5, 2)) @ basis + rng.normal(0, 0.01, (5, 3))
        A = pts.astype(np.float32)
        x = O.plane_fit(A)
        ref = np.linalg.lstsq(A.astype(np.float64), -np.ones(5), rcond=None)[0]
        assert np.allclose(x, ref, rtol=2e-3, atol=1e-5)


def test_plane_fit_rank_deficient_gives_basic_solution():
    # all five points identical -> rank 1: Eigen returns a basic solution with zeros, no NaN
    A = np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (5, 1))
    x = O.plane_fit(A)
    assert np.isfinite(x).all()
    assert np.count_nonzero(x) == 1
    assert np.allclose(A @ x, -1, atol=1e-5)


def test_pose_to_affine_matches_rotation_convention():
    # pcl::getTransformation: R = Rz(yaw) Ry(pitch) Rx(roll)
    from tests.synth import rot_rpy
    rng = np.random.default_rng(8)
    for _ in range(100):
        pose = rng.uniform(-1, 1, 6).astype(np.float32)
        T = O.pose_to_affine(pose).reshape(3, 4)
        assert np.allclose(T[:, :3], rot_rpy(*pose[:3].astype(np.float64)), atol=3e-7)
        assert np.array_equal(T[:, 3], pose[3:])


def test_transform_is_unfused_fp32():
    rng = np.random.default_rng(9)
    pts = rng.uniform(-50, 50, (1000, 4)).astype(np.float32)
    pose = np.array([0.1, -0.2, 0.7, 3, -4, 5], np.float32)
    T = O.pose_to_affine(pose)
    out = O.transform_cloud(pts, T=T)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for r in range(3):
        ref = ((T[4 * r] * x + T[4 * r + 1] * y) + T[4 * r + 2] * z) + T[4 * r + 3]   # fp32 numpy, one rounding per op
        assert np.array_equal(out[:, r], ref)
    assert np.array_equal(out[:, 3], pts[:, 3])


def _numpy_voxel_keys(pts, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    mn = pts[:, :3].min(0)
    mx = pts[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(pts[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int32)
    return (ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]).astype(np.uint32)


@pytest.mark.parametrize("leaf,n", [(0.2, 5000), (0.4, 20000), (2.0, 300)])
def test_voxelgrid_keys_and_centroids_vs_numpy(leaf, n):
    rng = np.random.default_rng(int(leaf * 10) + n)
    pts = np.concatenate([rng.uniform(-20, 20, (n, 3)) * [1, 1, 0.1], rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    out, keys, okeys, passthrough = O.voxelgrid(pts, leaf)
    assert not passthrough
    ref_keys = _numpy_voxel_keys(pts, leaf)
    assert np.array_equal(keys, ref_keys)
    uk = np.unique(ref_keys)
    assert np.array_equal(okeys, uk)            # ascending voxel idx, one output per voxel
    # centroids: sequential fp32 sums in input order within each voxel
    order = np.argsort(ref_keys, kind="stable")
    sk = ref_keys[order]
    starts = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1]])
    ends = np.r_[starts[1:], len(sk)]
    for v in rng.choice(len(starts), size=min(200, len(starts)), replace=False):
        acc = np.zeros(4, np.float32)
        for j in order[starts[v]:ends[v]]:
            acc = acc + pts[j]
        assert np.array_equal(out[v], acc / np.float32(ends[v] - starts[v]))


def test_voxelgrid_overflow_passthrough():
    # dx*dy*dz > INT32_MAX -> PCL warns and returns the input unchanged
    pts = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2], [1, 1, 1, 3]], np.float32)
    out, _, _, passthrough = O.voxelgrid(pts, 0.2)
    assert passthrough and np.array_equal(out, pts)


def test_voxelgrid_empty_and_single():
    out, _, _, _ = O.voxelgrid(np.zeros((0, 4), np.float32), 0.2)
    assert len(out) == 0
    p = np.array([[1.5, -2.5, 3.0, 9.0]], np.float32)
    out, keys, _, _ = O.voxelgrid(p, 0.4)
    assert np.array_equal(out, p) and keys[0] == 0


def test_kdtree_matches_brute_and_scipy():
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(11)
    m = 30000
    mp = np.concatenate([rng.uniform(-20, 20, (m, 3)) * [1, 1, 0.2], np.zeros((m, 1))], 1).astype(np.float32)
    q = np.concatenate([rng.uniform(-22, 22, (3000, 3)) * [1, 1, 0.2], np.zeros((3000, 1))], 1).astype(np.float32)
    tree = O.KdTree(mp)
    idx, d2 = tree.knn(q, 5)
    bidx, bd2 = O.knn5_brute(mp, q)
    assert np.array_equal(idx, bidx) and np.array_equal(d2, bd2)
    # numpy fp32 distances, (d2, index) order
    for i in rng.choice(len(q), 50, replace=False):
        d = mp[:, :3] - q[i, :3]
        dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        order = np.lexsort((np.arange(m), dd))[:5]
        assert np.array_equal(order, idx[i]) and np.array_equal(dd[order], d2[i])
    # scipy (float64 distances): same sets whenever the 5th/6th gap is not a rounding tie
    _, sidx = cKDTree(mp[:, :3].astype(np.float64)).query(q[:, :3].astype(np.float64), k=6)
    same = sum(set(sidx[i, :5]) == set(idx[i]) for i in range(len(q)))
    assert same >= len(q) - 3


def test_kdtree_exact_ties_resolved_by_index():
    # lattice map: many exactly equal distances
    g = np.arange(-5, 6, dtype=np.float32)
    mp = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    mp = np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1)
    q = np.array([[0, 0, 0, 0], [0.5, 0.5, 0.5, 0], [2, -3, 1, 0]], np.float32)
    idx, d2 = O.KdTree(mp).knn(q, 5)
    bidx, bd2 = O.knn5_brute(mp, q)
    assert np.array_equal(idx, bidx) and np.array_equal(d2, bd2)
    assert (np.diff(d2, axis=1) >= 0).all()
    for r in range(len(q)):
        for j in range(4):
            if d2[r, j] == d2[r, j + 1]:
                assert idx[r, j] < idx[r, j + 1]


def test_knn_small_maps():
    mp = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0]], np.float32)
    q = np.array([[0.1, 0, 0, 0]], np.float32)
    idx, d2 = O.KdTree(mp).knn(q, 5)
    assert list(idx[0]) == [0, 1, 2, -1, -1] and np.isinf(d2[0, 3:]).all()
    idx, d2 = O.knn5_brute(mp, q)
    assert list(idx[0]) == [0, 1, 2, -1, -1]


def test_radius_search_sorted_strict():
    rng = np.random.default_rng(12)
    mp = np.concatenate([rng.uniform(-60, 60, (500, 3)), np.zeros((500, 1))], 1).astype(np.float32)
    idx, d2 = O.KdTree(mp).radius(mp[-1], 50.0)
    d = mp[:, :3] - mp[-1, :3]
    dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    ref = np.flatnonzero(dd < np.float32(2500.0))
    ref = ref[np.lexsort((ref, dd[ref]))]
    assert np.array_equal(idx, ref) and idx[0] == 499


# ---- loop-closure ICP ---------------------------------------------------------------
def _moments(s, t, d2=None):
    s = np.asarray(s, np.float64)
    t = np.asarray(t, np.float64)
    mom = np.zeros(17)
    mom[0] = len(s)
    mom[1] = 0.0 if d2 is None else float(np.sum(d2))
    mom[2:5], mom[5:8], mom[8:17] = s.sum(0), t.sum(0), (t.T @ s).reshape(9)
    return mom


def _umeyama_numpy(s, t):
    # Eigen::umeyama(src, dst, with_scaling=false), in float64
    ms, mt = s.mean(0), t.mean(0)
    sig = (t - mt).T @ (s - ms) / len(s)
    U, S, Vt = np.linalg.svd(sig)
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    return R, mt - R @ ms


def test_umeyama_matches_numpy_svd():
    from tests.synth import rot_rpy
    rng = np.random.default_rng(5)
    for k in range(300):
        n = int(rng.integers(3, 400))
        s = rng.normal(0, rng.choice([0.1, 5.0, 60.0]), (n, 3)) + rng.uniform(-100, 100, 3)
        R0 = rot_rpy(*rng.uniform(-3, 3, 3))
        t = s @ R0.T + rng.uniform(-20, 20, 3) + rng.normal(0, 0.01, (n, 3))
        if k % 5 == 0:
            t[:, 2] = t[:, 2].mean()                  # coplanar target: rank-2 covariance
        T = O.umeyama_from_moments(_moments(s, t))
        R, tr = _umeyama_numpy(s, t)
        assert np.abs(T[:3, :3] - R).max() < 2e-6
        assert np.abs(T[:3, 3] - tr).max() < 2e-5 * max(1.0, np.abs(tr).max())
        assert abs(np.linalg.det(T[:3, :3].astype(np.float64)) - 1.0) < 1e-5
        assert np.array_equal(T[3], [0, 0, 0, 1])


def test_umeyama_reflection_case_gives_a_rotation():
    # a mirrored copy: the best orthogonal map is a reflection, Umeyama must return det = +1
    rng = np.random.default_rng(6)
    s = rng.normal(0, 1, (200, 3))
    t = s * np.array([1, 1, -1])
    T = O.umeyama_from_moments(_moments(s, t))
    R, tr = _umeyama_numpy(s, t)
    assert np.linalg.det(T[:3, :3].astype(np.float64)) > 0.999
    assert np.abs(T[:3, :3] - R).max() < 2e-6


def test_nn1_matches_numpy_brute_force():
    rng = np.random.default_rng(8)
    tgt = rng.uniform(-20, 20, (3000, 4)).astype(np.float32)
    q = rng.uniform(-25, 25, (400, 4)).astype(np.float32)
    idx, d2 = O.nn1(tgt, q)
    dx = q[:, None, 0] - tgt[None, :, 0]
    dy = q[:, None, 1] - tgt[None, :, 1]
    dz = q[:, None, 2] - tgt[None, :, 2]
    D = (dx * dx + dy * dy).astype(np.float32) + dz * dz            # the fp32 operation order of the path
    assert np.array_equal(idx, D.argmin(1).astype(np.int32))
    assert np.array_equal(d2, D.min(1))


def test_icp_align_recovers_a_known_motion_and_reports_pcl_states():
    from tests.synth import rot_rpy, room_world
    rng = np.random.default_rng(9)
    cw, sw = room_world(rng, n_surf=9000, n_corner=2400, half=6.0, height=4.0)
    tgt = O.voxelgrid(np.concatenate([cw, sw]), 0.4)[0]
    pose = np.array([0.01, -0.02, 0.08, 0.3, -0.25, 0.1])
    R = rot_rpy(*pose[:3])
    sel = tgt[rng.choice(len(tgt), 1500, replace=False)]
    src = sel.copy()
    src[:, :3] = ((sel[:, :3].astype(np.float64) - pose[3:]) @ R).astype(np.float32)
    res = O.icp_align(src, tgt)
    assert res.converged == 1 and res.state in (2, 3, 4)        # TRANSFORM / ABS_MSE / REL_MSE
    assert np.abs(res.T[:3, :3] - R).max() < 1e-3 and np.abs(res.T[:3, 3] - pose[3:]).max() < 1e-2
    assert res.fitness < 1e-4
    # iteration cap -> PCL reports convergence with state ITERATIONS
    capped = O.icp_align(src, tgt, O.icp_default_params(max_iterations=2))
    assert capped.converged == 1 and capped.state == 1 and capped.iterations == 2
    # nothing within the correspondence distance -> not converged, zero iterations, identity
    far = src.copy()
    far[:, 0] += 1000
    none = O.icp_align(far, tgt)
    assert none.converged == 0 and none.state == 5 and none.iterations == 0
    assert np.array_equal(none.T, np.eye(4, dtype=np.float32))


def test_correct_pose_composes_like_pcl():
    from tests.synth import rot_rpy
    rng = np.random.default_rng(12)
    for _ in range(100):
        pose = np.concatenate([rng.uniform(-1.2, 1.2, 3), rng.uniform(-50, 50, 3)]).astype(np.float32)
        corr = np.eye(4)
        corr[:3, :3] = rot_rpy(*rng.uniform(-0.2, 0.2, 3))
        corr[:3, 3] = rng.uniform(-2, 2, 3)
        out = O.correct_pose(corr.astype(np.float32), pose)
        W = np.eye(4)
        W[:3, :3] = rot_rpy(*pose[:3])
        W[:3, 3] = pose[3:]
        Cm = corr @ W
        assert np.abs(out[3:] - Cm[:3, 3]).max() < 1e-4
        assert np.abs(rot_rpy(*out[:3]) - Cm[:3, :3]).max() < 1e-5


# ---- LiDAR depth for visual features ----------------------------------------------
def test_get_depth_against_a_numpy_restatement():
    """independent numpy version of feature_tracker.h:150-283 on a small cloud (identity transform)"""
    rng = np.random.default_rng(17)
    # a wall at x = 8 m seen through the camera cone, plus clutter behind the camera
    n = 6000
    wall = np.stack([np.full(n, 8.0) + rng.normal(0, 0.02, n), rng.uniform(-9, 9, n), rng.uniform(-6, 6, n),
                     rng.uniform(0, 255, n)], 1).astype(np.float32)
    back = np.stack([rng.uniform(-9, -1, 500), rng.uniform(-9, 9, 500), rng.uniform(-3, 3, 500), np.zeros(500)], 1).astype(np.float32)
    dc = np.concatenate([wall, back])
    f = np.ones((80, 3), np.float32)
    f[:, 0] = rng.uniform(-0.8, 0.8, 80)
    f[:, 1] = rng.uniform(-0.5, 0.5, 80)
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    depth, f3, local = O.get_depth(dc, ident, f)
    # range image
    x, y, z = dc[:, 0], dc[:, 1], dc[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        keep = (x >= 0) & (np.abs(y / x) <= 10) & (np.abs(z / x) <= 10)
    row = np.round((np.arctan2(z.astype(np.float64), np.sqrt(x * x + y * y).astype(np.float64)).astype(np.float32).astype(np.float64)
                    * 180.0 / np.pi + 90.0).astype(np.float32) / np.float32(0.5)).astype(int)
    col = np.round((np.arctan2(x.astype(np.float64), y.astype(np.float64)).astype(np.float32).astype(np.float64)
                    * 180.0 / np.pi).astype(np.float32) / np.float32(0.5)).astype(int)
    keep &= (row >= 0) & (row < 360) & (col >= 0) & (col < 360)
    dist = np.sqrt((x * x + y * y).astype(np.float32) + z * z)
    best = {}
    for i in np.flatnonzero(keep):
        b = row[i] * 360 + col[i]
        if b not in best or dist[i] < dist[best[b]]:
            best[b] = i
    order = [best[b] for b in sorted(best)]
    assert np.array_equal(local, dc[order])
    # depth of a feature that looks at the wall: close to 8 m along the optical axis
    got = depth[depth > 0]
    assert len(got) > 40
    assert np.all(np.abs(got - 8.0) < 0.15)
    # features_3d: unit vectors (no depth, intensity -1) or scaled by s with intensity = x
    for i in range(len(f)):
        v = f3[i, :3].astype(np.float64)
        if depth[i] > 0:
            assert f3[i, 3] == f3[i, 0] == depth[i]
        elif f3[i, 3] == -1:
            assert abs(np.linalg.norm(v) - 1.0) < 1e-6


# ---- deskew + range-image projection ----------------------------------------------
def test_find_rotation_matches_linear_interpolation():
    rng = np.random.default_rng(21)
    t = 50.0 + np.cumsum(rng.uniform(0.004, 0.006, 30))
    rot = np.cumsum(rng.normal(0, 0.01, (30, 3)), 0)
    for pt in np.concatenate([rng.uniform(t[0], t[-1], 200), t[[3, 7, 29]]]):
        got = O.find_rotation(pt, t, rot)
        want = np.array([np.interp(pt, t, rot[:, k]) for k in range(3)])
        assert np.abs(got - want).max() < 1e-6
    # before the first sample: the first sample; after the last: the last (imageProjection.cpp:507-511)
    assert np.array_equal(O.find_rotation(t[0] - 1.0, t, rot), rot[0].astype(np.float32))
    assert np.array_equal(O.find_rotation(t[-1] + 1.0, t, rot), rot[-1].astype(np.float32))


def test_project_cloud_against_a_python_loop():
    """projectPointCloud + cloudExtraction (imageProjection.cpp:571-647) restated as a plain loop, Livox mode,
    no deskew: column counters advance for every point that passes the range / ring / rate filters"""
    rng = np.random.default_rng(22)
    n, NS, H = 3000, 4, 400
    xyzi = np.concatenate([rng.normal(0, 6, (n, 3)), rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    ring = rng.integers(0, NS + 2, n).astype(np.uint16)
    rel = np.zeros(n, np.float32)
    ext, rg, col, sr, er = O.project_cloud(xyzi, ring, rel, n_scan=NS, horizon_scan=H, downsample_rate=2, sensor=2,
                                           lidar_min_range=2.0, lidar_max_range=15.0)
    counters = [0] * NS
    cells = {}
    for i in range(n):
        x, y, z = xyzi[i, :3]
        r = np.sqrt(np.float32(np.float32(x * x + y * y) + z * z))
        if r < 2.0 or r > 15.0 or ring[i] >= NS or ring[i] % 2 != 0:
            continue
        c = counters[ring[i]]
        counters[ring[i]] += 1
        if c >= H or (int(ring[i]), c) in cells:
            continue
        cells[(int(ring[i]), c)] = (i, r)
    order = sorted(cells)
    assert np.array_equal(ext, xyzi[[cells[k][0] for k in order]])
    assert np.array_equal(col, np.array([k[1] for k in order], np.int32))
    assert np.array_equal(rg, np.array([cells[k][1] for k in order], np.float32))
    count = 0
    for r_ in range(NS):
        assert sr[r_] == count - 1 + 5
        count += sum(1 for k in order if k[0] == r_)
        assert er[r_] == count - 1 - 5
    assert any(v > H for v in counters)          # the overflow branch was exercised


def test_deskew_is_identity_for_a_constant_rotation_and_undoes_a_known_spin():
    from tests.synth import rot_rpy
    rng = np.random.default_rng(23)
    n = 500
    xyzi = np.concatenate([rng.normal(0, 8, (n, 3)), np.zeros((n, 1))], 1).astype(np.float32)
    ring = np.zeros(n, np.uint16)
    rel = np.linspace(0, 0.1, n).astype(np.float32)
    t = np.linspace(-0.01, 0.12, 27)
    const = np.tile([0.1, -0.05, 0.3], (27, 1))
    ext, *_ = O.project_cloud(xyzi, ring, rel, n_scan=1, horizon_scan=n, sensor=2, lidar_min_range=0.0, deskew=True,
                              time_scan_cur=0.0, imu_time=t, imu_rot=const)
    assert np.abs(ext[:, :3] - xyzi[:, :3]).max() < 2e-5           # R_start^-1 R_start = identity up to rounding
    # pure yaw spin at 1 rad/s: a point seen at time s is rotated back by yaw(s) - yaw(0)
    spin = np.stack([np.zeros(27), np.zeros(27), t - t[0]], 1)
    ext, *_ = O.project_cloud(xyzi, ring, rel, n_scan=1, horizon_scan=n, sensor=2, lidar_min_range=0.0, deskew=True,
                              time_scan_cur=0.0, imu_time=t, imu_rot=spin)
    for i in (0, 100, 499):
        want = rot_rpy(0, 0, float(rel[i]) - float(rel[0])) @ xyzi[i, :3].astype(np.float64)
        assert np.abs(ext[i, :3] - want).max() < 1e-4
