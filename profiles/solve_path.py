import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LVREG_DEBUG_TILES"] = "1"
import bench
import lidar_visual_inertial_slam_b200 as lv
for wl in ("c1", "c3"):
    h = lv.Lvreg()
    ds = bench.make_dataset(wl, bench.SEED, lambda p, l: h.voxelgrid(p, l)[0], lambda m: None)
    for i in range(len(ds["kf_pose"])):
        h.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i])
    ids = np.arange(len(ds["kf_pose"]), dtype=np.int32)
    for j in range(3):
        c, s = ds["scans"][j]
        pose, res, st = h.register_scan(c, s, ids, ds["guess"][j])
        print(wl, j, "iters", res.iterations, "path", h.debug_stage_stats()[5], "profile", np.round(h.iteration_profile(), 1).tolist())
    h.close()
