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
    icp_golden()
    depth_golden()
    projection_golden()
    print("golden vectors written to", HERE)


def icp_golden():
    """4. loop-closure ICP ("next" row 8f-2): 1-NN, one Umeyama step, the full align() and the pose correction"""
    rng = np.random.default_rng(20261019)
    cw, sw = room_world(rng, n_surf=9000, n_corner=2400, half=6.0, height=4.0)
    tgt = O.voxelgrid(np.concatenate([cw, sw]), 0.4)[0]
    pose = np.array([0.02, -0.015, 0.06, 0.35, -0.3, 0.1])
    from tests.synth import rot_rpy
    R = rot_rpy(*pose[:3])
    sel = tgt[rng.choice(len(tgt), 1200, replace=False)]
    src = sel.copy()
    src[:, :3] = ((sel[:, :3].astype(np.float64) - pose[3:]) @ R + rng.normal(0, 0.01, (1200, 3))).astype(np.float32)
    idx, d2 = O.nn1(tgt, src)
    s, t = src[:, :3].astype(np.float64), tgt[idx, :3].astype(np.float64)
    mom = np.zeros(17)
    mom[0], mom[1] = len(s), d2.astype(np.float64).sum()
    mom[2:5], mom[5:8], mom[8:17] = s.sum(0), t.sum(0), (t.T @ s).reshape(9)
    T1 = O.umeyama_from_moments(mom)
    res = O.icp_align(src, tgt)
    stored = np.array([0.01, 0.02, -0.4, 3.0, -1.0, 0.2], np.float32)
    np.savez_compressed(os.path.join(HERE, "icp.npz"), src=src, tgt=tgt, pose=pose, nn_idx=idx, nn_d2=d2, moments=mom,
                        first_T=T1, final_T=res.T, iterations=np.int32(res.iterations), state=np.int32(res.state),
                        converged=np.int32(res.converged), fitness=np.float64(res.fitness),
                        n_corr=np.int32(res.n_correspondences), stored_pose=stored,
                        corrected_pose=O.correct_pose(res.T, stored))


def depth_golden():
    """5. LiDAR depth for visual features ("next" row 8f-3): stacked depth cloud + get_depth"""
    rng = np.random.default_rng(20261020)
    cw, sw = room_world(rng, n_surf=30000, n_corner=3000, half=8.0, height=5.0)
    world = np.concatenate([cw, sw]).astype(np.float32)
    from tests.synth import rot_rpy
    reg = O.DepthRegister()
    clouds, Ts, stamps = [], [], []
    for k in range(3):
        pose = np.array([0.01 * k, -0.02, 0.15 * k, 0.4 * k - 1.0, 0.1 * k, 0.05], np.float32)
        R = rot_rpy(*pose[:3])
        sel = world[rng.choice(len(world), 1200, replace=False)]
        loc = np.concatenate([(sel[:, :3].astype(np.float64) - pose[3:].astype(np.float64)) @ R, sel[:, 3:4]], 1).astype(np.float32)
        T = O.pose_to_affine(pose)
        reg.add_cloud(loc, T, 3.0 * k)                # the first cloud expires when the third arrives
        clouds.append(loc); Ts.append(T); stamps.append(3.0 * k)
    depth_cloud = reg.cloud()
    pose = np.array([0.02, -0.01, 0.3, 0.5, -0.2, 0.1], np.float32)
    T4 = np.eye(4)
    T4[:3] = O.pose_to_affine(pose).reshape(3, 4)
    Tinv = np.linalg.inv(T4)[:3].astype(np.float32).reshape(12)
    f = np.ones((60, 3), np.float32)
    f[:, 0] = rng.uniform(-0.9, 0.9, 60)
    f[:, 1] = rng.uniform(-0.6, 0.6, 60)
    dense = world[rng.choice(len(world), 7000, replace=False)]
    d, f3, local = O.get_depth(dense, Tinv, f)
    np.savez_compressed(os.path.join(HERE, "depth.npz"), cloud0=clouds[0], cloud1=clouds[1], cloud2=clouds[2],
                        T_now=np.array(Ts), stamps=np.array(stamps), depth_cloud=depth_cloud, dense=dense, T_inv=Tinv,
                        features=f, depth=d, features_3d=f3, local=local)


def projection_golden():
    """6. deskew + range-image projection ("next" row 8f-4)"""
    rng = np.random.default_rng(20261021)
    n, NS, H = 6000, 4, 1200
    d = rng.normal(0, 1, (n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyzi = np.concatenate([d * rng.uniform(0.3, 28, (n, 1)), rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    ring = rng.integers(0, NS + 1, n).astype(np.uint16)
    rel = np.sort(rng.uniform(0, 0.1, n)).astype(np.float32)
    t = 200.0 - 0.01 + np.arange(30) * 0.005
    rot = np.cumsum(rng.normal(0, 0.004, (30, 3)), 0)
    ext, rg, col, sr, er = O.project_cloud(xyzi, ring, rel, n_scan=NS, horizon_scan=H, sensor=2, lidar_min_range=1.0,
                                           lidar_max_range=25.0, deskew=True, time_scan_cur=200.0, imu_time=t, imu_rot=rot)
    np.savez_compressed(os.path.join(HERE, "projection.npz"), xyzi=xyzi, ring=ring, rel_time=rel, imu_time=t, imu_rot=rot,
                        n_scan=np.int32(NS), horizon=np.int32(H), extracted=ext, point_range=rg, point_col_ind=col,
                        start_ring_index=sr, end_ring_index=er)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "projection":
        projection_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "icp":
        icp_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "depth":
        depth_golden()
    else:
        main()
