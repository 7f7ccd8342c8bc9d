"""GPU parity of the visual-side depth association: lvreg_depth_add_cloud /
lvreg_get_depth against oracle/oracle_depth.cpp.  Everything is compared bit-exactly: the stacked
depth cloud, the range-image survivors (order included), depths and the published 3-D features."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O                      # noqa: E402
from tests.synth import room_world, rot_rpy           # noqa: E402


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


@pytest.fixture(scope="module")
def h(lv):
    hd = lv.Lvreg()
    yield hd
    hd.close()


def affine(pose):
    return O.pose_to_affine(np.asarray(pose, np.float32))


def inverse_affine(T12):
    T = np.eye(4)
    T[:3] = np.asarray(T12, np.float64).reshape(3, 4)
    return np.linalg.inv(T)[:3].astype(np.float32).reshape(12)


def sensor_cloud(rng, world, pose, n):
    R = rot_rpy(*pose[:3])
    sel = world[rng.choice(len(world), n, replace=False)]
    loc = (sel[:, :3].astype(np.float64) - np.asarray(pose[3:], np.float64)) @ R
    return np.concatenate([loc, sel[:, 3:4]], 1).astype(np.float32)


def features(rng, n):
    f = np.ones((n, 3), np.float32)
    f[:, 0] = rng.uniform(-0.9, 0.9, n)
    f[:, 1] = rng.uniform(-0.6, 0.6, n)
    return f


@pytest.fixture(scope="module")
def world():
    rng = np.random.default_rng(41)
    cw, sw = room_world(rng, n_surf=200000, n_corner=20000)
    return np.concatenate([cw, sw]).astype(np.float32)


def test_depth_stack_bit_exact(h, world):
    rng = np.random.default_rng(1)
    h.depth_clear()
    od = O.DepthRegister()
    sizes = []
    for k in range(9):
        pose = np.array([0.01 * k, -0.02, 0.1 * k, 0.3 * k - 2.0, 0.1 * k, 0.05], np.float32)
        cloud = sensor_cloud(rng, world, pose, 30000)
        T = affine(pose)
        stamp = 0.9 * k                                   # 9 clouds over 7.2 s: the first ones expire
        n = h.depth_add_cloud(cloud, T, stamp)
        m = od.add_cloud(cloud, T, stamp)
        assert n == m
        got = h.depth_get_cloud(0)
        assert np.array_equal(got, od.cloud())
        sizes.append(n)
    assert sizes[-1] > 1000
    # an empty scan still ages the queue
    n = h.depth_add_cloud(np.zeros((0, 4), np.float32), affine(np.zeros(6)), 20.0)
    m = od.add_cloud(np.zeros((0, 4), np.float32), affine(np.zeros(6)), 20.0)
    assert n == m == 0


@pytest.mark.parametrize("n_feat", [1, 150, 1000])
def test_get_depth_bit_exact(h, world, n_feat):
    rng = np.random.default_rng(n_feat)
    pose = np.array([0.02, -0.01, 0.4, 1.0, -0.5, 0.1], np.float32)
    Tinv = inverse_affine(affine(pose))
    dc = world[rng.choice(len(world), 120000, replace=False)]
    f = features(rng, n_feat)
    h.depth_set_cloud(dc)
    gd, g3 = h.get_depth(Tinv, f)
    od, o3, olocal = O.get_depth(dc, Tinv, f)
    glocal = h.depth_get_cloud(1)
    assert np.array_equal(glocal, olocal)
    assert np.array_equal(gd, od)
    assert np.array_equal(g3, o3)
    if n_feat >= 150:
        assert (gd > 0).sum() > n_feat // 4               # a good share of the features gets a depth
        # depth = distance along the optical axis of a point on the room's walls
        assert gd[gd > 0].min() > 3.0 and gd.max() < 40.0


def test_get_depth_edge_cases(h, world):
    rng = np.random.default_rng(5)
    f = features(rng, 20)
    ident = affine(np.zeros(6))
    # empty depth cloud, and fewer than 10 survivors: no depth at all
    for dc in (np.zeros((0, 4), np.float32), np.array([[5, 0.1 * i, 0.2 * i, 1] for i in range(6)], np.float32)):
        h.depth_set_cloud(dc)
        gd, g3 = h.get_depth(ident, f)
        od, o3, ol = O.get_depth(dc, ident, f)
        assert np.all(gd == -1) and np.array_equal(gd, od) and np.array_equal(g3, o3)
        assert np.array_equal(h.depth_get_cloud(1), ol)
    # points behind the camera, on the x = 0 plane and exactly on the view boundary
    dc = np.array([[-1, 0, 0, 1], [0, 1, 1, 2], [0, 0, 0, 3], [1, 10, 0, 4], [1, 10.001, 0, 5], [2, 0, -20, 6],
                   [3, 1, 1, 7]] + [[6, 0.05 * i, 0.03 * i, 8] for i in range(40)], np.float32)
    h.depth_set_cloud(dc)
    gd, g3 = h.get_depth(ident, f)
    od, o3, ol = O.get_depth(dc, ident, f)
    assert np.array_equal(h.depth_get_cloud(1), ol) and np.array_equal(gd, od) and np.array_equal(g3, o3)
    # other range-image resolutions
    dc = world[rng.choice(len(world), 50000, replace=False)]
    h.depth_set_cloud(dc)
    for nb in (90, 720):
        gd, g3 = h.get_depth(ident, f, num_bins=nb)
        od, o3, ol = O.get_depth(dc, ident, f, num_bins=nb)
        assert np.array_equal(h.depth_get_cloud(1), ol) and np.array_equal(gd, od) and np.array_equal(g3, o3)


def test_depth_golden_vectors(h):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth.npz"))
    h.depth_clear()
    for k in range(3):
        h.depth_add_cloud(z["cloud%d" % k], z["T_now"][k], float(z["stamps"][k]))
    assert np.array_equal(h.depth_get_cloud(0), z["depth_cloud"])
    h.depth_set_cloud(z["dense"])
    d, f3 = h.get_depth(z["T_inv"], z["features"])
    assert np.array_equal(d, z["depth"]) and np.array_equal(f3, z["features_3d"])
    assert np.array_equal(h.depth_get_cloud(1), z["local"])
