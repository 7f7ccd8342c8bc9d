"""extractCloud + pcl::VoxelGrid (MO:931-973) of a LARGE local map through the sample-sort path
(csrc/voxelgrid_bucket.cuh: cached world-frame keyframe clouds kept in voxel order, buckets sorted in shared
memory) against the CPU restatement: bit-exact maps, in PCL's output order, for

  * the default path (and the counters say the sample-sort path really ran),
  * a voxel with thousands of points (one bucket close to its capacity),
  * a forced bucket overflow (LVREG_VG_BUCKET_CAP) -> the job is redone by the device-wide sort,
  * the device-wide sort on the same (voxel-ordered) cache (LVREG_VG_BUCKET=0),
  * corrected keyframe poses (correctPoses, MO:1607-1640: the cache is dropped and rebuilt),
  * a growing map (keyframes added between builds),
  * more keyframes than threads of a bucket block (segment table built in several rounds), some of them without
    corner points, one of them a single point.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as O   # noqa: E402


@pytest.fixture(scope="module")
def lv():
    import lidar_visual_inertial_slam_b200 as lvmod
    return lvmod


def _room_cloud(rng, n, half=14.0, height=5.0, noise=0.02):
    """points on the six faces of a room + a few pillars, sensor frame"""
    face = rng.integers(0, 7, n)
    u = rng.uniform(-half, half, n).astype(np.float32)
    v = rng.uniform(-half, half, n).astype(np.float32)
    w = rng.uniform(0.0, height, n).astype(np.float32)
    x = np.where(face == 0, -half, np.where(face == 1, half, u))
    y = np.where(face == 2, -half, np.where(face == 3, half, v))
    z = np.where(face == 4, 0.0, np.where(face == 5, height, w))
    pil = face == 6
    x = np.where(pil, np.round(u / 4.0) * 4.0, x)
    y = np.where(pil, np.round(v / 4.0) * 4.0, y)
    pts = np.stack([x, y, z - 1.5, rng.uniform(0, 100, n)], 1).astype(np.float32)
    pts[:, :3] += rng.normal(0, noise, (n, 3)).astype(np.float32)
    return pts


def _keyframes(seed, n_kf=14, n_corner=24000, n_surf=42000, dense_voxel=0):
    rng = np.random.default_rng(seed)
    kfs = []
    for i in range(n_kf):
        pose = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(-3.1, 3.1),
                         rng.uniform(-6, 6), rng.uniform(-6, 6), rng.uniform(-0.3, 0.3)], np.float32)
        c = _room_cloud(rng, n_corner + int(rng.integers(0, 999)))
        s = _room_cloud(rng, n_surf + int(rng.integers(0, 999)))
        if dense_voxel and i == 3:
            blob = np.tile(np.array([[1.03, 2.01, 0.11, 5.0]], np.float32), (dense_voxel, 1))
            blob[:, :3] += rng.uniform(0, 0.05, (dense_voxel, 3)).astype(np.float32)
            s = np.concatenate([s[:1000], blob, s[1000:]])
        kfs.append((c, s, pose))
    return kfs


def _oracle_maps(kfs, n_use=None, poses=None):
    mo = O.MapOptimization(O.default_params(num_threads=8))
    for i, (c, s, pose) in enumerate(kfs[:n_use]):
        mo.add_keyframe(c, s, pose if poses is None else poses[i], float(i))
    mo.build_local_map(np.arange(len(kfs[:n_use]), dtype=np.int32))
    return mo.get_map(0), mo.get_map(1)


def _device_maps(lv, h, n):
    info = h.build_local_map(np.arange(n, dtype=np.int32))
    return info, h.get_local_map(lv.CORNER), h.get_local_map(lv.SURF)


def _fill(lv, kfs):
    h = lv.Lvreg()
    for c, s, pose in kfs:
        h.add_keyframe(c, s, pose)
    return h


def test_bucket_path_equals_oracle(lv):
    kfs = _keyframes(1)
    oc, os_ = _oracle_maps(kfs)
    h = _fill(lv, kfs)
    info, gc, gs = _device_maps(lv, h, len(kfs))
    assert info.n_corner_in > 262144 and info.n_surf_in > 262144
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    st = h.debug_stage_stats()
    assert st[7] == 2 and st[6] == 0, st          # both classes through the sample-sort path, no fallback
    # a second build re-uses the ordered cache and gives the same maps
    _, gc2, gs2 = _device_maps(lv, h, len(kfs))
    assert np.array_equal(gc2, oc) and np.array_equal(gs2, os_)
    assert h.debug_stage_stats()[7] == 4
    h.close()


def test_bucket_path_with_a_crowded_voxel(lv):
    kfs = _keyframes(2, dense_voxel=1200)
    oc, os_ = _oracle_maps(kfs)
    h = _fill(lv, kfs)
    _, gc, gs = _device_maps(lv, h, len(kfs))
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    assert h.debug_stage_stats()[6] == 0
    h.close()


def test_bucket_overflow_falls_back_to_the_sort(lv, monkeypatch):
    kfs = _keyframes(3, dense_voxel=9000)          # one voxel alone exceeds the 4096-point bucket
    oc, os_ = _oracle_maps(kfs)
    h = _fill(lv, kfs)
    _, gc, gs = _device_maps(lv, h, len(kfs))
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    st = h.debug_stage_stats()
    assert st[6] == 1 and st[7] == 1, st           # the surf job was redone, the corner job was not
    h.close()
    monkeypatch.setenv("LVREG_VG_BUCKET_CAP", "1500")
    h = _fill(lv, kfs)
    _, gc, gs = _device_maps(lv, h, len(kfs))
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    assert h.debug_stage_stats()[6] == 2
    h.close()


def test_sort_path_on_the_ordered_cache_and_without_it(lv, monkeypatch):
    kfs = _keyframes(4)
    oc, os_ = _oracle_maps(kfs)
    monkeypatch.setenv("LVREG_VG_BUCKET", "0")
    h = _fill(lv, kfs)
    _, gc, gs = _device_maps(lv, h, len(kfs))
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    assert h.debug_stage_stats()[7] == 0
    h.close()


def test_corrected_poses_and_growing_map(lv):
    kfs = _keyframes(5, n_kf=16)
    h = lv.Lvreg()
    for c, s, pose in kfs[:12]:
        h.add_keyframe(c, s, pose)
    oc, os_ = _oracle_maps(kfs, 12)
    _, gc, gs = _device_maps(lv, h, 12)
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    for c, s, pose in kfs[12:]:
        h.add_keyframe(c, s, pose)
    oc, os_ = _oracle_maps(kfs, 16)
    _, gc, gs = _device_maps(lv, h, 16)
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    # loop-closure correction: every pose moves, the cached clouds are stale
    rng = np.random.default_rng(99)
    poses = np.stack([k[2] for k in kfs]).astype(np.float32)
    poses[:, 2] += rng.uniform(-0.02, 0.02, len(kfs)).astype(np.float32)
    poses[:, 3:] += rng.uniform(-0.2, 0.2, (len(kfs), 3)).astype(np.float32)
    h.update_keyframe_poses(poses)
    oc, os_ = _oracle_maps(kfs, 16, poses)
    _, gc, gs = _device_maps(lv, h, 16)
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    assert h.debug_stage_stats()[6] == 0
    h.close()


def test_many_small_keyframes(lv):
    rng = np.random.default_rng(6)
    kfs = []
    for i in range(700):
        pose = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), rng.uniform(-3.1, 3.1),
                         rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-0.2, 0.2)], np.float32)
        nc = 0 if i % 7 == 3 else 420 + int(rng.integers(0, 60))
        ns = 1 if i == 11 else 800 + int(rng.integers(0, 200))
        c = _room_cloud(rng, nc) if nc else np.zeros((0, 4), np.float32)
        kfs.append((c, _room_cloud(rng, ns), pose))
    oc, os_ = _oracle_maps(kfs)
    h = _fill(lv, kfs)
    info, gc, gs = _device_maps(lv, h, len(kfs))
    assert info.n_corner_in > 262144 and info.n_surf_in > 262144
    assert np.array_equal(gc, oc) and np.array_equal(gs, os_)
    st = h.debug_stage_stats()
    assert st[7] == 2 and st[6] == 0, st
    h.close()
