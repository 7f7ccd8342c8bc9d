"""Registration-kernel comparison on the C3 bench data set: register_ms per variant, per-iteration phase profile,
staging statistics.  Usage: python profiles/reg_compare.py [c3|c1] > gpurun_out/reg_compare.json"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lidar_visual_inertial_slam_b200 as lv

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
out = {}
ds = None
for variant in ("warm", "tpq", "staged"):
    for debug in ("0", "1"):
        os.environ["LVREG_REG"] = variant
        os.environ["LVREG_DEBUG_TILES"] = debug
        h = lv.Lvreg()
        if ds is None:
            ds = bench.make_dataset(wl, bench.SEED, lambda p, l: h.voxelgrid(p, l)[0], lambda m: print(m, file=sys.stderr))
        for i in range(len(ds["kf_pose"])):
            h.add_keyframe(ds["kf_corner"][i], ds["kf_surf"][i], ds["kf_pose"][i])
        ids = np.arange(len(ds["kf_pose"]), dtype=np.int32)
        reg, tot, iters, poses = [], [], [], []
        for rep in range(3):
            for j in range(bench.N_RING_SCANS):
                c, s = ds["scans"][j]
                pose, res, st = h.register_scan(c, s, ids, ds["guess"][j])
                t = h.timings()
                if rep > 0:
                    reg.append(t.register_ms); tot.append(t.total_ms); iters.append(res.iterations)
                    poses.append(pose.tolist())
        key = variant + ("_debug" if debug == "1" else "")
        c, s = ds["scans"][0]
        pose, res, st = h.register_scan(c, s, ids, ds["guess"][0])
        out[key] = dict(register_ms=float(np.mean(reg)), total_ms=float(np.mean(tot)), iters=float(np.mean(iters)),
                        register_ms_per_iter=float(np.sum(reg) / np.sum(iters)),
                        profile_scan0=np.round(h.iteration_profile(), 1).tolist(), poses=poses[:8])
        if debug == "1":
            out[key]["stage_stats"] = h.debug_stage_stats().tolist()
            tt = h.debug_tile_times()
            if len(tt):
                out[key]["tile_us_pct"] = np.round(np.percentile(tt / 1e3, [5, 25, 50, 75, 95, 99, 100]), 1).tolist()
                nct = (res.n_corner_ds + 31) // 32
                out[key]["tile_us_corner_mean"] = float(tt[:nct].mean() / 1e3)
                out[key]["tile_us_surf_mean"] = float(tt[nct:].mean() / 1e3)
        h.close()
# all variants must produce the same poses
ref = out["tpq"]["poses"]
out["poses_identical"] = all(out[k]["poses"] == ref for k in out if isinstance(out[k], dict))
for k in list(out):
    if isinstance(out[k], dict):
        out[k].pop("poses")
print(json.dumps(out, indent=1))
