import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import lidar_visual_inertial_slam_b200 as lv
from lidar_visual_inertial_slam_b200 import harness as H
gen=H.Generator(H.BEAM128, 0x5EED0000)
mo=H.MapOptimizationMirror()
for k in range(14):
    truth=gen.truth_pose(k,0.2,1.0)
    c,s=gen.scan(truth,1000003*k,8)
    guess=truth if k==0 else gen.guess_pose(k,truth,0.10,0.035)
    t0=time.perf_counter()
    st,pose,res,tim,nkf=mo.handle_scan(c,s,0.2*k,guess)
    w=(time.perf_counter()-t0)*1e3
    print(k,"wall %.2f ms"%w,"upload %.2f ds %.2f map %.2f grid %.2f reg %.2f total %.2f"%(tim.upload_ms,tim.downsample_ms,tim.map_build_ms,tim.grid_build_ms,tim.register_ms,tim.total_ms),"iters",res.iterations,"nkf",nkf,"n",len(c),len(s),res.n_corner_ds,res.n_surf_ds,res.n_corner_map,res.n_surf_map)
