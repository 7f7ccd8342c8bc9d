import sys, time
sys.path.insert(0,'/root/repo')
from lidar_visual_inertial_slam_b200 import harness as H
for n in (30, 60, 60):
    r = H.replay(H.MID360, 0x5EED0000, n, device=0, period=0.2, gen_threads=4)
    print(n, {k: round(v,4) for k,v in r.items()})
