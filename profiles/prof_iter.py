import sys, numpy as np
sys.path.insert(0,'/root/repo')
import bench, lidar_visual_inertial_slam_b200 as lv
h=lv.Lvreg()
ds=bench.make_dataset("c3", bench.SEED, lambda p,l: h.voxelgrid(p,l)[0], lambda m: None)
for i in range(len(ds["kf_pose"])): h.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i])
ids=np.arange(len(ds["kf_pose"]),dtype=np.int32)
for j in range(4):
    c,s=ds["scans"][j]
    pose,res,st=h.register_scan(c,s,ids,ds["guess"][j])
    prof=h.iteration_profile()
    print("scan",j,"iters",res.iterations,"register_ms %.3f"%h.timings().register_ms)
    print(np.round(prof,1)); print(" sum per phase", np.round(prof.sum(0),1), "total", round(float(prof.sum()),1))
