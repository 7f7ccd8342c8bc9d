# per-phase clock cycles of rs_onesweep_kernel, summed over tiles (thread 0 of each block);
# needs a build with -DLVREG_SORT_PROF (debug only, not the shipped library)
import sys, ctypes
sys.path.insert(0, '/root/repo')
import lidar_visual_inertial_slam_b200 as lv
from lidar_visual_inertial_slam_b200 import binding
h = lv.Lvreg()
lib = ctypes.CDLL('/root/repo/lidar_visual_inertial_slam_b200/liblvreg.so')
names = ["ticket+zero+sync", "key loads", "tile histogram + sync", "publish + look-back", "ballot match + counter chain x16",
         "val loads + scans + offsets + syncs", "stage to smem + sync", "write-out"]
for n, bits in ((11375817, 28), (1660000, 31), (62000, 24)):
    h.bench_sort(n, bits, 2)
    out = (ctypes.c_ulonglong * 16)()
    lib.lvreg_debug_sort_prof(out, 1)
    reps = 4
    ms, p = h.bench_sort(n, bits, reps)
    lib.lvreg_debug_sort_prof(out, 0)
    tiles = ((n + 4095) // 4096) * p * (reps + 2)
    tot = sum(out[:8])
    print("n=%d passes=%d us/pass=%.1f  tiles=%d  cycles/tile=%.0f" % (n, p, ms * 1e3 / p, tiles, tot / tiles))
    print("   look-back: %.2f batches/tile, %.2f of them met an unpublished tile, inclusive prefix found %.1f entries into its batch"
          % (out[10] / tiles, out[11] / max(1, out[10]), out[12] / tiles))
    for k in range(8):
        print("   %-32s %8.0f cyc  %5.1f%%" % (names[k], out[k] / tiles, 100.0 * out[k] / tot))
