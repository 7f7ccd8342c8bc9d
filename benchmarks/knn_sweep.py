#!/usr/bin/env python
"""BASELINE C4: kNN micro-benchmark sweep, Nq in {1k..100k} x M in {10k..4M}, grid (gated / exact)
vs brute force, device-resident data, CUDA-event timing inside the library (lvreg_bench_knn5).

    python benchmarks/knn_sweep.py [--quick] > profiles/r01_knn_sweep.json

HBM fraction (grid) uses the compulsory per-query bytes of SURVEY 8d (56 B/query; 96 B/query with the fused
residual); the 16 B x M map term is NOT counted for grid searches (they touch a few cells per query).
FP32 fraction (brute) uses 8 flops per pair against the nominal 148 SM x 128 lanes x 1.965 GHz."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lidar_visual_inertial_slam_b200 as lv   # noqa: E402


def make_map(rng, m):
    """ground + facades at ~6 pts/m^2 (what a 0.4 m VoxelGrid leaves on planar structure)"""
    side = float(np.sqrt(m / 6.0))
    n_ground = int(0.7 * m)
    g = np.stack([rng.uniform(0, side, n_ground), rng.uniform(0, side, n_ground), rng.normal(0, 0.02, n_ground)], 1)
    n_wall = m - n_ground
    wx = rng.integers(0, max(2, int(side / 10)), n_wall) * 10.0 + rng.normal(0, 0.02, n_wall)
    w = np.stack([wx, rng.uniform(0, side, n_wall), rng.uniform(0, 8, n_wall)], 1)
    pts = np.concatenate([g, w]).astype(np.float32)
    return np.concatenate([pts, np.zeros((m, 1), np.float32)], 1), side


def make_queries(rng, mp, side, nq):
    near = mp[rng.integers(0, len(mp), int(0.9 * nq))][:, :3] + rng.normal(0, 0.05, (int(0.9 * nq), 3))
    far = np.stack([rng.uniform(0, side, nq - len(near)), rng.uniform(0, side, nq - len(near)),
                    rng.uniform(0, 8, nq - len(near))], 1)
    q = np.concatenate([near, far]).astype(np.float32)
    return np.concatenate([q, np.zeros((nq, 1), np.float32)], 1)


def run_sweep(device=0, quick=False, log=None):
    """the whole sweep on one GPU; returns the result dict (rows = one per (M, Nq, variant))"""
    nqs = [1000, 10000, 100000] if quick else [1000, 3000, 10000, 30000, 100000]
    ms_ = [10000, 1000000] if quick else [10000, 100000, 1000000, 4000000]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    fp32_peak = 148 * 128 * 1.965e9          # lane-ops/s, nominal
    rng = np.random.default_rng(4)
    h = lv.Lvreg(device=device)
    rows = []
    for m in ms_:
        mp, side = make_map(rng, m)
        info = h.set_local_map(mp[:16], mp)
        for nq in nqs:
            q = make_queries(rng, mp, side, nq)
            for name, variant in (("grid_gated", lv.KNN_GRID_GATED), ("grid_staged", lv.KNN_GRID_STAGED),
                                  ("grid_exact", lv.KNN_GRID_EXACT), ("brute", lv.KNN_BRUTE)):
                if variant == lv.KNN_BRUTE and nq * m > 4e11:
                    continue
                reps = 3 if variant == lv.KNN_BRUTE and nq * m > 1e10 else 10
                ms = h.bench_knn5(lv.SURF, q, variant, reps)
                row = dict(M=m, Nq=nq, variant=name, ms=ms, queries_per_s=nq / (ms * 1e-3))
                if variant == lv.KNN_BRUTE:
                    row["tflops"] = 8.0 * nq * m / (ms * 1e-3) / 1e12
                    row["fp32_frac_of_nominal"] = 8.0 * nq * m / (ms * 1e-3) / fp32_peak
                else:
                    # compulsory per-query traffic only (query in, 5 indices + 5 distances out).  A grid search touches a
                    # few cells per query, not the whole map: the map term of SURVEY 8d (16 B x M) does not apply to it
                    # (it made a 15 us launch over a 64 MB map look like 60 % of HBM); the cells it does read are L2
                    # hits and are reported from ncu (profiles/r02_ncu_knn_*_summary.txt)
                    gbs = 56.0 * nq / (ms * 1e-3) / 1e9
                    row["algorithmic_gbs"] = gbs
                    row["hbm_frac_of_measured"] = gbs / hbm
                rows.append(row)
                if log:
                    log(json.dumps(row))
            # "with fused residual/Jacobian" (SURVEY 8d C4): search + plane fit + residual in one kernel; per query
            # 16 B read + 16 B coefficients + 1 B flag written + 5 x 16 B neighbour coordinates gathered (L2)
            ms = h.bench_residuals(lv.SURF, q, None, 10)
            gbs = 96.0 * nq / (ms * 1e-3) / 1e9
            row = dict(M=m, Nq=nq, variant="grid_gated_fused_residual", ms=ms, queries_per_s=nq / (ms * 1e-3),
                       algorithmic_gbs=gbs, hbm_frac_of_measured=gbs / hbm)
            rows.append(row)
            if log:
                log(json.dumps(row))
    out = dict(benchmark="C4 kNN sweep", grid_cell_m=float(info.grid_cell[1]), hbm_peak_gbs=hbm, rows=rows)
    h.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    out = run_sweep(0, args.quick, lambda m: print(m, file=sys.stderr, flush=True))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
