import sys
sys.path.insert(0,'/root/repo')
import lidar_visual_inertial_slam_b200 as lv
h=lv.Lvreg()
for n,bits in ((11375817,28),(1660000,31),(1000000,24),(500000,24),(262144,24),(200000,24),(100000,24),(62000,24),(20000,24),(4000,24),(4000,8)):
    ms,p=h.bench_sort(n,bits,8)
    print("n=%9d bits=%2d passes=%d  %.1f us/sort  %.1f us/pass  %.0f GB/s per pass"%(n,bits,p,ms*1e3,ms*1e3/p,16.0*n*p/(ms*1e-3)/1e9))
