import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from lidar_visual_inertial_slam_b200 import harness as H
gen = H.Generator(H.MID360, 0x5EED0000)
mo = H.MapOptimizationMirror()
scans=[]
for k in range(30):
    truth = gen.truth_pose(k, 0.2, 1.0)
    scans.append((truth,)+gen.scan(truth, 100 + k, 4))
for k,(truth,c,s) in enumerate(scans):
    guess = truth if k == 0 else gen.guess_pose(k, truth, 0.08, 0.02)
    t0=time.perf_counter()
    st, pose, res, tim, nkf = mo.handle_scan(c, s, k * 0.2, guess)
    dt=(time.perf_counter()-t0)*1e3
    print("scan %2d wall %.2f ms st %d kf %d | map %.3f grid %.3f ds %.3f reg %.3f launches %d iters %d"%(k,dt,st,nkf,tim.map_build_ms,tim.grid_build_ms,tim.downsample_ms,tim.register_ms,tim.kernel_launches,res.iterations))
