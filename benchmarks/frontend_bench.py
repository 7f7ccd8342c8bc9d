#!/usr/bin/env python
"""Front-end chain (rows 8f-4 + 8f-1): raw 128 x 2048 scan -> deskew + range-image projection ->
FeatureExtraction, on the device (one upload, the deskewed cloud never leaves the GPU) vs the CPU oracle.

    python benchmarks/frontend_bench.py > profiles/r01_frontend_bench.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lidar_visual_inertial_slam_b200 as lv   # noqa: E402
from oracle import pyoracle as O               # noqa: E402
from benchmarks.feature_bench import fast_ring_scan   # noqa: E402


def main():
    rng = np.random.default_rng(11)
    n_scan, horizon = 128, 2048
    pts, rg, col, sr, er = fast_ring_scan(rng, n_scan, horizon)
    ring = pts[:, 3].astype(np.uint16)                    # fast_ring_scan stores the ring in the intensity channel
    az = np.arctan2(pts[:, 1], pts[:, 0])
    rel = ((az % (2 * np.pi)) / (2 * np.pi) * 0.1).astype(np.float32)
    order = np.argsort(rel, kind="stable")               # a spinning sensor delivers points in time order
    pts, ring, rel = pts[order], ring[order], rel[order]
    t = 10.0 - 0.01 + np.arange(60) * 0.002
    rot = np.cumsum(rng.normal(0, 0.002, (60, 3)), 0)
    kw = dict(n_scan=n_scan, horizon_scan=horizon, sensor=0, lidar_min_range=0.5, lidar_max_range=1000.0, deskew=True,
              time_scan_cur=10.0, imu_time=t, imu_rot=rot)
    raw = lv.make_raw_cloud(pts, ring, rel, lv.LAYOUT_VELODYNE)
    h = lv.Lvreg()
    for _ in range(3):
        n = h.project_cloud(raw, layout=lv.LAYOUT_VELODYNE, **kw)
        nc, ns = h.extract_features_projected()
    reps = 20
    proj_ms, feat_ms = [], []
    t0 = time.perf_counter()
    for _ in range(reps):
        n = h.project_cloud(raw, layout=lv.LAYOUT_VELODYNE, **kw)
        proj_ms.append(h.timings().downsample_ms)
        nc, ns = h.extract_features_projected()
        feat_ms.append(h.timings().downsample_ms)
    gpu_wall = (time.perf_counter() - t0) / reps
    g = h.download_projection()
    t0 = time.perf_counter()
    o = O.project_cloud(pts, ring, rel, **kw)
    cpu_proj = time.perf_counter() - t0
    t0 = time.perf_counter()
    oc, os_, ol = O.extract_features(*o)
    cpu_feat = time.perf_counter() - t0
    out = {
        "what": "raw scan -> projectPointCloud/deskew/cloudExtraction (imageProjection.cpp:495-647) -> FeatureExtraction (featureExtraction.cpp:87-245)",
        "raw_points": int(len(pts)), "extracted_points": int(n), "corner_features": int(nc), "surf_features": int(ns),
        "gpu_wall_ms": gpu_wall * 1e3, "gpu_projection_device_ms": float(np.median(proj_ms)),
        "gpu_features_device_ms": float(np.median(feat_ms)),
        "cpu_projection_ms": cpu_proj * 1e3, "cpu_features_ms": cpu_feat * 1e3, "h2d_bytes": int(raw.nbytes),
        "parity": {"projection_bit_exact": bool(all(np.array_equal(a, b) for a, b in zip(g, o))),
                   "feature_counts_equal": bool((nc, ns) == (len(oc), len(os_)))},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
