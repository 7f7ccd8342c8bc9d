#!/usr/bin/env python
"""Depth-association row (8f-3) timing: lvreg_depth_add_cloud (one 128-beam scan into the 5 s stack) and
lvreg_get_depth (150 features) on the device vs the CPU oracle restatement.

    python benchmarks/depth_bench.py > profiles/r01_depth_bench.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lidar_visual_inertial_slam_b200 as lv   # noqa: E402
from lidar_visual_inertial_slam_b200 import harness as H   # noqa: E402
from oracle import pyoracle as O               # noqa: E402


def main():
    gen = H.Generator(H.BEAM128, 0x5EED0077)
    h = lv.Lvreg()
    od = O.DepthRegister()
    rng = np.random.default_rng(2)
    scans, Ts = [], []
    for k in range(12):                                  # 10 Hz lidar, LIDAR_SKIP = 0 ... 5 s window holds ~50; 12 keep the CPU leg short
        pose = np.array([0.0, 0.0, 0.01 * k, 0.5 * k, 0.0, 0.0], np.float32)
        c, s = gen.scan(pose, 300 + k, 8)
        scans.append(np.concatenate([c, s]))
        Ts.append(O.pose_to_affine(pose))
    # the two legs run one after the other: the oracle's OpenMP workers keep spinning after a
    # parallel region and would slow the CUDA host thread's synchronisations down
    gpu_add, cpu_add, ns, ms = [], [], [], []
    for rep in range(2):                                 # second round: buffers are recycled, no allocation
        if rep:
            h.depth_clear()
            gpu_add, ns = [], []
        for k in range(12):
            t0 = time.perf_counter()
            ns.append(h.depth_add_cloud(scans[k], Ts[k], 0.4 * k))
            gpu_add.append(time.perf_counter() - t0)
    for k in range(12):
        t0 = time.perf_counter()
        ms.append(od.add_cloud(scans[k], Ts[k], 0.4 * k))
        cpu_add.append(time.perf_counter() - t0)
    assert ns == ms
    same_stack = bool(np.array_equal(h.depth_get_cloud(0), od.cloud()))
    f = np.ones((150, 3), np.float32)
    f[:, 0] = rng.uniform(-0.9, 0.9, 150)
    f[:, 1] = rng.uniform(-0.6, 0.6, 150)
    T4 = np.eye(4)
    T4[:3] = Ts[-1].reshape(3, 4)
    Tinv = np.linalg.inv(T4)[:3].astype(np.float32).reshape(12)
    for _ in range(3):
        gd, g3 = h.get_depth(Tinv, f)
    t0 = time.perf_counter()
    for _ in range(20):
        gd, g3 = h.get_depth(Tinv, f)
    gpu_get = (time.perf_counter() - t0) / 20
    dc = od.cloud()
    t0 = time.perf_counter()
    cd, c3, cl = O.get_depth(dc, Tinv, f)
    cpu_get = time.perf_counter() - t0
    out = {
        "what": "lidar_callback stack (feature_tracker_node.cpp:303-371) and DepthRegister::get_depth (feature_tracker.h:150-283)",
        "points_per_scan": int(np.mean([len(s) for s in scans])), "stack_points": int(len(dc)), "features": 150,
        "features_with_depth": int((gd > 0).sum()),
        "gpu_add_cloud_ms": float(np.median(gpu_add[3:]) * 1e3), "cpu_add_cloud_ms": float(np.median(cpu_add[3:]) * 1e3),
        "gpu_get_depth_ms": gpu_get * 1e3, "cpu_get_depth_ms": cpu_get * 1e3,
        "parity": {"stack_bit_exact": same_stack, "depth_bit_exact": bool(np.array_equal(gd, cd)),
                   "features_3d_bit_exact": bool(np.array_equal(g3, c3)),
                   "local_cloud_bit_exact": bool(np.array_equal(h.depth_get_cloud(1), cl))},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
