import sys, numpy as np
sys.path.insert(0,'/root/repo')
import lidar_visual_inertial_slam_b200 as lv
from oracle import pyoracle as O
from tests.synth import room_world
rng = np.random.default_rng(100)
cw, sw = room_world(rng)
sm = O.voxelgrid(sw, 0.4)[0]
rng = np.random.default_rng(7)
h=lv.Lvreg(); h.set_local_map(sm, sm)
near = sm[rng.choice(len(sm), 4000)] + np.r_[rng.normal(0, 0.05, 3), 0].astype(np.float32)
near = (sm[rng.choice(len(sm), 4000)][:, :3] + rng.normal(0, 0.05, (4000, 3))).astype(np.float32)
far = rng.uniform(-15, 15, (600, 3)).astype(np.float32)
q = np.concatenate([np.concatenate([near, far]), np.zeros((4600, 1), np.float32)], 1)
ridx, rd2 = O.knn5_brute(sm, q)
idx, d2 = h.knn5(lv.SURF, q, lv.KNN_GRID_GATED)
inside = rd2[:,4] < 1.0
bad = np.flatnonzero(inside & ((idx!=ridx).any(1) | (d2!=rd2).any(1)))
print("inside", inside.sum(), "bad", len(bad))
for b in bad[:5]:
    print(q[b], "\n gpu", idx[b], d2[b], "\n ref", ridx[b], rd2[b])
bad2=np.flatnonzero(~inside & (d2[:,4] < 1.0)); print("bad outside", len(bad2))
