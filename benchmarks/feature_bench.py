#!/usr/bin/env python
"""FeatureExtraction ("next" row 8f-1) timing: device (lvreg_extract_features, host buffers in, feature
clouds left on the device) vs the CPU oracle restatement, on a synthetic 128 x 2048 ring-ordered scan.

    python benchmarks/feature_bench.py > profiles/r01_feature_bench.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lidar_visual_inertial_slam_b200 as lv   # noqa: E402
from oracle import pyoracle as O               # noqa: E402


def fast_ring_scan(rng, n_scan, horizon):
    """vectorised variant of tests/synth.ring_scan: box room + depth steps, full rings"""
    el = np.deg2rad(np.linspace(-22.5, 22.5, n_scan))[:, None]
    az = (2 * np.pi * np.arange(horizon) / horizon)[None, :]
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)], -1)
    t = np.minimum(np.minimum(30.0 / np.abs(d[..., 0]).clip(1e-9), 20.0 / np.abs(d[..., 1]).clip(1e-9)),
                   3.0 / np.abs(d[..., 2]).clip(1e-9))
    steps = (np.floor(az * 40 / (2 * np.pi)) % 3 == 0)           # every third 9-degree wedge is 35% closer
    t = np.where(steps, t * 0.65, t) + rng.normal(0, 0.01, t.shape)
    keep = rng.uniform(size=t.shape) > 0.02
    pts, rg, col, sr, er = [], [], [], [], []
    count = 0
    for r in range(n_scan):
        sr.append(count - 1 + 5)
        k = np.flatnonzero(keep[r])
        p = d[r, k] * t[r, k, None]
        pts.append(np.concatenate([p, np.full((len(k), 1), float(r))], 1))
        rg.append(t[r, k])
        col.append(k)
        count += len(k)
        er.append(count - 1 - 5)
    return (np.concatenate(pts).astype(np.float32), np.concatenate(rg).astype(np.float32),
            np.concatenate(col).astype(np.int32), np.array(sr, np.int32), np.array(er, np.int32))


def main():
    rng = np.random.default_rng(11)
    pts, rg, col, sr, er = fast_ring_scan(rng, 128, 2048)
    h = lv.Lvreg()
    for _ in range(3):
        c, s, l = h.extract_features(pts, rg, col, sr, er)
    t0 = time.perf_counter()
    reps = 20
    dev = []
    for _ in range(reps):
        c, s, l = h.extract_features(pts, rg, col, sr, er)
        dev.append(h.timings().downsample_ms)
    gpu_wall = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    rc, rs, rl = O.extract_features(pts, rg, col, sr, er)
    cpu = time.perf_counter() - t0
    ok = bool(np.array_equal(c, rc) and np.array_equal(s, rs) and np.array_equal(l, rl))
    print(json.dumps(dict(benchmark="FeatureExtraction 128x2048", points=int(len(pts)), corners=int(len(c)), surf=int(len(s)),
                          gpu_device_ms=float(np.mean(dev)), gpu_wall_ms_incl_copies=gpu_wall * 1e3,
                          cpu_oracle_ms=cpu * 1e3, speedup_wall=cpu / gpu_wall, bit_exact_vs_oracle=ok)))
    h.close()


if __name__ == "__main__":
    main()
