import sys
sys.path.insert(0,'/root/repo')
import lidar_visual_inertial_slam_b200 as lv
h=lv.Lvreg()
ms,p=h.bench_sort(4000,24,4)
print(ms*1e3/p,"us/pass")
