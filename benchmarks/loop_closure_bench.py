#!/usr/bin/env python
"""Loop-closure row (8f-2) timing: lvreg_perform_loop_closure (submaps of 1 and 2x25+1 keyframes,
VoxelGrid, grid build, ICP, fitness) on the device vs the CPU oracle restatement, on a synthetic
there-and-back drive through the 128-beam urban world.

    python benchmarks/loop_closure_bench.py > profiles/r01_loop_closure_bench.json"""
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
    n_kf = int(os.environ.get("LOOP_BENCH_KEYFRAMES", "60"))
    gen = H.Generator(H.BEAM128, 0x5EED0042)
    h = lv.Lvreg()
    mo = O.MapOptimization(O.default_params(num_threads=os.cpu_count() or 8))
    # out along the street and back: keyframe k and keyframe n-1-k are 0.5 m apart
    xs = np.concatenate([np.arange(n_kf // 2) * 2.0, (n_kf // 2 - 1 - np.arange(n_kf - n_kf // 2)) * 2.0 + 0.5])
    rng = np.random.default_rng(3)
    for k, x in enumerate(xs):
        pose = np.array([0.0, 0.0, 0.02 * np.sin(k), x, 0.2 * np.cos(0.3 * k), 0.0], np.float32)
        c, s = gen.scan(pose, 900 + k, 8)
        cds, sds = O.voxelgrid(c, 0.2)[0], O.voxelgrid(s, 0.4)[0]
        stored = pose.copy()
        if k == n_kf - 1:
            stored += np.array([0.004, -0.003, 0.01, 0.25, -0.2, 0.05], np.float32)
        h.add_keyframe(cds, sds, stored)
        mo.add_keyframe(cds, sds, stored, 2.0 * k)
    cur, pre = n_kf - 1, 0
    prm = lv.icp_default_params()
    for _ in range(2):
        g = h.perform_loop_closure(cur, pre, 25, prm)
    reps = 10
    t0 = time.perf_counter()
    stages = []
    for _ in range(reps):
        g = h.perform_loop_closure(cur, pre, 25, prm)
        t = h.timings()
        stages.append((t.map_build_ms, t.grid_build_ms, t.register_ms, t.kernel_launches))
    gpu_wall = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    o = mo.perform_loop_closure(cur, pre, 25, O.icp_default_params(num_threads=os.cpu_count() or 8))
    cpu = time.perf_counter() - t0
    st = np.median(np.array(stages), 0)
    out = {
        "what": "performLoopClosure from the candidate pair to the pose constraint (MO:566-613)",
        "keyframes": n_kf, "n_source": g.n_source, "n_target": g.n_target,
        "icp_iterations": g.icp.iterations, "icp_state": g.icp.state, "status": g.status,
        "gpu_ms_wall": gpu_wall * 1e3,
        "gpu_stages_ms": {"submaps_voxelgrid": float(st[0]), "target_grid": float(st[1]), "icp_and_fitness": float(st[2])},
        "gpu_kernel_launches": int(st[3]),
        "cpu_oracle_ms": cpu * 1e3, "cpu_threads": os.cpu_count(),
        "speedup": cpu / gpu_wall,
        "parity": {
            "status_equal": bool(g.status == o.status),
            "sizes_equal": bool((g.n_source, g.n_target) == (o.n_source, o.n_target)),
            "iterations_equal": bool(g.icp.iterations == o.icp.iterations and g.icp.state == o.icp.state),
            "max_abs_T_diff": float(np.abs(g.icp.T - o.icp.T).max()),
            "fitness_gpu": g.icp.fitness, "fitness_cpu": o.icp.fitness,
        },
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
