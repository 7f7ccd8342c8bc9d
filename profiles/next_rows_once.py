# one invocation of each "next" row (8f-1..4) at benchmark size, for an ncu launch list
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lidar_visual_inertial_slam_b200 as lv
from lidar_visual_inertial_slam_b200 import harness as H
from benchmarks.feature_bench import fast_ring_scan

rng = np.random.default_rng(11)
h = lv.Lvreg()
# front end: raw scan -> projection -> features
pts, rg, col, sr, er = fast_ring_scan(rng, 128, 2048)
ring = pts[:, 3].astype(np.uint16)
rel = np.linspace(0, 0.1, len(pts)).astype(np.float32)
t = 10.0 - 0.01 + np.arange(60) * 0.002
rot = np.cumsum(rng.normal(0, 0.002, (60, 3)), 0)
raw = lv.make_raw_cloud(pts, ring, rel, lv.LAYOUT_VELODYNE)
for _ in range(2):
    h.project_cloud(raw, layout=lv.LAYOUT_VELODYNE, n_scan=128, horizon_scan=2048, sensor=0, deskew=True,
                    time_scan_cur=10.0, imu_time=t, imu_rot=rot)
    h.extract_features_projected()
# loop closure on a there-and-back drive
gen = H.Generator(H.BEAM128, 0x5EED0042)
n_kf = 40
xs = np.concatenate([np.arange(n_kf // 2) * 2.0, (n_kf // 2 - 1 - np.arange(n_kf - n_kf // 2)) * 2.0 + 0.5])
for k, x in enumerate(xs):
    pose = np.array([0.0, 0.0, 0.02 * np.sin(k), x, 0.2 * np.cos(0.3 * k), 0.0], np.float32)
    c, s = gen.scan(pose, 900 + k, 8)
    if k == n_kf - 1:
        pose = pose + np.array([0.004, -0.003, 0.01, 0.25, -0.2, 0.05], np.float32)
    h.add_keyframe(h.voxelgrid(c, 0.2)[0], h.voxelgrid(s, 0.4)[0], pose)
for _ in range(2):
    r = h.perform_loop_closure(n_kf - 1, 0, 25)
# depth association
T = lv.pose_to_affine(np.zeros(6, np.float32))
for k in range(3):
    c, s = gen.scan(np.array([0, 0, 0, 0.5 * k, 0, 0], np.float32), 300 + k, 8)
    h.depth_add_cloud(np.concatenate([c, s]), T, 0.4 * k)
f = np.ones((150, 3), np.float32)
f[:, 0] = rng.uniform(-0.9, 0.9, 150)
f[:, 1] = rng.uniform(-0.6, 0.6, 150)
for _ in range(2):
    d, f3 = h.get_depth(T, f)
print("ok", r.status, r.icp.iterations, int((d > 0).sum()))
