"""Driver for ncu captures of the registration kernel on the C3 bench data set: a few registrations, nothing else.
LVREG_REG selects the kernel variant."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lidar_visual_inertial_slam_b200 as lv
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
h = lv.Lvreg()
ds = bench.make_dataset(wl, bench.SEED, lambda p, l: h.voxelgrid(p, l)[0], lambda m: print(m, file=sys.stderr))
for i in range(len(ds["kf_pose"])):
    h.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i])
ids = np.arange(len(ds["kf_pose"]), dtype=np.int32)
for j in range(n):
    c, s = ds["scans"][j % bench.N_RING_SCANS]
    pose, res, st = h.register_scan(c, s, ids, ds["guess"][j % bench.N_RING_SCANS])
    print(j, st, res.iterations, h.timings().register_ms)
h.close()
