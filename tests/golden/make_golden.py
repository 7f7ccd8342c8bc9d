"""Generates the golden vectors under tests/golden/ from the CPU oracle (seeded inputs).

The reference ships no tests or fixtures for this path, and it cannot be
built or imported here, so these vectors are produced by oracle/ -- whose dense-math kernels are
pinned bit-exactly against the OpenCV wheel (tests/test_oracle_pins.py).  Run from the repo root:

    python tests/golden/make_golden.py

The .npz files are small on purpose (a few hundred KB in total)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O                    # noqa: E402
from tests.synth import room_world, scan_from_world, ring_scan  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261018)
    # 1. VoxelGrid: keys + centroids
    pts = np.concatenate([rng.uniform(-12, 12, (6000, 3)) * [1, 1, 0.15], rng.uniform(0, 255, (6000, 1))], 1).astype(np.float32)
    out, keys, okeys, _ = O.voxelgrid(pts, 0.4)
    np.savez_compressed(os.path.join(HERE, "voxelgrid.npz"), pts=pts, leaf=np.float32(0.4), out=out, keys=keys, out_keys=okeys)

    # 2. a small registration problem: maps, DS scan, guess -> kNN, residuals, LM step, full loop
    cw, sw = room_world(rng, n_surf=9000, n_corner=2400, half=6.0, height=4.0)
    cm = O.voxelgrid(cw, 0.2)[0]
    sm = O.voxelgrid(sw, 0.4)[0]
    truth = np.array([0.015, -0.02, 0.25, 0.6, -0.9, 0.1], np.float32)
    c, s = scan_from_world(rng, cw, sw, truth, 500, 1500)
    cds = O.voxelgrid(c, 0.2)[0]
    sds = O.voxelgrid(s, 0.4)[0]
    guess = truth + np.array([0.015, -0.01, 0.02, 0.07, -0.05, 0.04], np.float32)
    T = O.pose_to_affine(guess)
    q = O.transform_cloud(sds, T=T)
    knn_idx, knn_d2 = O.knn5_brute(sm, q)
    ccoef, cflag, cnn = O.corner_residuals(cm, cds, guess)
    scoef, sflag, snn = O.surf_residuals(sm, sds, guess)
    ori = np.concatenate([cds[cflag == 1], sds[sflag == 1]])
    coef = np.concatenate([ccoef[cflag == 1], scoef[sflag == 1]])
    conv, pose1, AtA, Atb, x, _ = O.lm_step(ori, coef, 0, guess)
    pose, res, _ = O.scan2map(cm, sm, cds, sds, guess)
    np.savez_compressed(os.path.join(HERE, "registration.npz"), corner_map=cm, surf_map=sm, corner_ds=cds, surf_ds=sds,
                        truth=truth, guess=guess, affine=T, surf_queries=q, knn_idx=knn_idx, knn_d2=knn_d2,
                        corner_coeff=ccoef, corner_flag=cflag, corner_nn=cnn, surf_coeff=scoef, surf_flag=sflag,
                        surf_nn=snn, lm_AtA=AtA, lm_Atb=Atb, lm_x=x, lm_pose=pose1, lm_conv=np.int32(conv),
                        final_pose=pose, iterations=np.int32(res.iterations), converged=np.int32(res.converged),
                        n_sel=np.array(res.n_sel[:res.iterations], np.int32),
                        pose_iter=np.array([list(res.pose_iter[i]) for i in range(res.iterations)], np.float32))
    # 3. FeatureExtraction ("next" row 8f-1)
    frng = np.random.default_rng(77)
    fpts, frg, fcol, fsr, fer = ring_scan(frng, 8, 600)
    fc, fs, fl = O.extract_features(fpts, frg, fcol, fsr, fer, edge_threshold=0.5)
    np.savez_compressed(os.path.join(HERE, "features.npz"), pts=fpts, point_range=frg, point_col_ind=fcol,
                        start_ring_index=fsr, end_ring_index=fer, edge_threshold=np.float32(0.5), corner=fc, surf=fs,
                        label=fl.astype(np.int8))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
