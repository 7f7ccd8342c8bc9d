"""Driver for ncu captures of the stage-level kNN kernels (C4 shapes): gated grid, staged grid, brute force."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lidar_visual_inertial_slam_b200 as lv
from benchmarks import knn_sweep as K
rng = np.random.default_rng(4)
mp, side = K.make_map(rng, 1000000)
q = K.make_queries(rng, mp, side, 100000)
h = lv.Lvreg()
h.set_local_map(mp[:16], mp)
for name, variant, nq in (("gated", lv.KNN_GRID_GATED, 100000), ("staged", lv.KNN_GRID_STAGED, 100000),
                          ("brute", lv.KNN_BRUTE, 20000)):
    ms = h.bench_knn5(lv.SURF, q[:nq], variant, 3)
    print(name, nq, "ms", ms, "Gq/s", nq / ms / 1e6, "TFLOP/s (brute)", 8.0 * nq * 1e6 / (ms * 1e-3) / 1e12)
h.close()
