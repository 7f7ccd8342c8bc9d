import os, sys, numpy as np
os.environ["LVREG_DEBUG_TILES"]="1"; os.environ.setdefault("LVREG_TPQ","1")
sys.path.insert(0,'/root/repo')
import bench, lidar_visual_inertial_slam_b200 as lv
h=lv.Lvreg()
ds=bench.make_dataset("c3", bench.SEED, lambda p,l: h.voxelgrid(p,l)[0], lambda m: None)
for i in range(len(ds["kf_pose"])): h.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i])
ids=np.arange(len(ds["kf_pose"]),dtype=np.int32)
for j in range(2):
    c,s=ds["scans"][j]
    pose,res,st=h.register_scan(c,s,ids,ds["guess"][j])
    t=h.debug_tile_times().astype(np.float64)/1e3
    ntc=(res.n_corner_ds+31)//32
    print("scan",j,"iters",res.iterations,"tiles",len(t),"corner tiles",ntc)
    for name,x in (("corner",t[:ntc]),("surf",t[ntc:])):
        print("  %s tile us: mean %.1f median %.1f p90 %.1f p99 %.1f max %.1f sum %.0f"%(name,x.mean(),np.median(x),np.percentile(x,90),np.percentile(x,99),x.max(),x.sum()))
    print("  iteration profile[1]:", h.iteration_profile()[1])
