"""Small seeded synthetic scenes for the parity tests (numpy; bench-scale data comes from the
C++ generator in the harness library).  Clouds are float32 [n,4] = x,y,z,intensity."""
import numpy as np


def rot_rpy(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def room_world(rng, n_surf=60000, n_corner=12000, half=12.0, height=6.0, noise=0.01):
    """points on the shell of a box room (surf) and on vertical/horizontal edges + poles (corner)"""
    surf = []
    per = n_surf // 6
    for axis in range(3):
        for sgn in (-1, 1):
            p = rng.uniform(-1, 1, (per, 3)) * np.array([half, half, height / 2])
            p[:, axis] = sgn * (half if axis < 2 else height / 2)
            surf.append(p)
    surf = np.concatenate(surf) + rng.normal(0, noise, (per * 6, 3))
    corner = []
    n_lines = 40
    per = n_corner // n_lines
    for k in range(n_lines):
        if k % 2 == 0:   # vertical pole
            base = np.array([rng.uniform(-half, half), rng.uniform(-half, half), 0.0])
            d = np.array([0, 0, 1.0])
            ext = height / 2
        else:            # horizontal edge
            base = np.array([rng.uniform(-half, half), rng.uniform(-half, half), rng.uniform(-height / 2, height / 2)])
            ang = rng.uniform(0, np.pi)
            d = np.array([np.cos(ang), np.sin(ang), 0.0])
            ext = 4.0
        t = rng.uniform(-ext, ext, per)
        corner.append(base + np.outer(t, d))
    corner = np.concatenate(corner) + rng.normal(0, noise, (per * n_lines, 3))

    def with_i(p):
        return np.concatenate([p, rng.uniform(0, 255, (len(p), 1))], axis=1).astype(np.float32)
    return with_i(corner), with_i(surf)


def scan_from_world(rng, corner_w, surf_w, pose, n_corner=800, n_surf=4000, noise=0.01):
    """sample world points and express them in the sensor frame of `pose` = [r,p,y,x,y,z]"""
    R = rot_rpy(*pose[:3])
    t = np.asarray(pose[3:6], dtype=np.float64)

    def pick(w, n):
        sel = w[rng.choice(len(w), size=min(n, len(w)), replace=False)].astype(np.float64)
        loc = (sel[:, :3] - t) @ R      # R^T (p - t)
        loc += rng.normal(0, noise, loc.shape)
        return np.concatenate([loc, sel[:, 3:4]], axis=1).astype(np.float32)
    return pick(corner_w, n_corner), pick(surf_w, n_surf)


def corridor_world(rng, n_surf=40000, n_corner=6000, length=30.0, half_w=2.0, height=3.0, noise=0.01):
    """corridor along x: walls y=+-half_w, floor/ceiling, edge lines parallel to x -> x is unobservable"""
    per = n_surf // 4
    surf = []
    for k in range(4):
        p = np.zeros((per, 3))
        p[:, 0] = rng.uniform(-length, length, per)
        if k < 2:
            p[:, 1] = (-1) ** k * half_w
            p[:, 2] = rng.uniform(-height / 2, height / 2, per)
        else:
            p[:, 1] = rng.uniform(-half_w, half_w, per)
            p[:, 2] = (-1) ** k * height / 2
        surf.append(p)
    surf = np.concatenate(surf) + rng.normal(0, noise, (per * 4, 3))
    per = n_corner // 4
    corner = []
    for sy in (-1, 1):
        for sz in (-1, 1):
            p = np.zeros((per, 3))
            p[:, 0] = rng.uniform(-length, length, per)
            p[:, 1] = sy * half_w
            p[:, 2] = sz * height / 2
            corner.append(p)
    corner = np.concatenate(corner) + rng.normal(0, noise, (per * 4, 3))

    def with_i(p):
        return np.concatenate([p, rng.uniform(0, 255, (len(p), 1))], axis=1).astype(np.float32)
    return with_i(corner), with_i(surf)


def ring_scan(rng, n_scan=16, horizon=1800, drop=0.03, noise=0.01, elev=(-15.0, 15.0), n_pillars=14):
    """A deskewed, ring-ordered cloud with the CloudInfo side channels of imageProjection.cpp:624-647:
    returns (pts [n,4], point_range, point_col_ind, start_ring_index, end_ring_index).  Scene: a box
    room with square pillars, so that there are depth discontinuities (edge features)."""
    pillars = [(rng.uniform(-4, 4), rng.uniform(-3, 3), rng.uniform(0.15, 0.4)) for _ in range(n_pillars)]
    pts, rg, col, sr, er = [], [], [], [], []
    count = 0
    for r in range(n_scan):
        sr.append(count - 1 + 5)
        el = np.deg2rad(elev[0] + (elev[1] - elev[0]) * r / max(1, n_scan - 1))
        for c in range(horizon):
            if rng.uniform() < drop:
                continue
            az = 2 * np.pi * c / horizon
            d = np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
            ts = []
            for ax, half in ((0, 5.0), (1, 4.0), (2, 1.5)):
                if abs(d[ax]) > 1e-9:
                    ts.append(half / abs(d[ax]))
            t = min(ts)
            for (px, py, hw) in pillars:          # axis-aligned square pillars, full height
                for ax, ctr, oth, octr in ((0, px, 1, py), (1, py, 0, px)):
                    if abs(d[ax]) < 1e-9:
                        continue
                    for face in (ctr - hw, ctr + hw):
                        tt = face / d[ax]
                        if 0.3 < tt < t and abs(d[oth] * tt - octr) <= hw:
                            t = tt
            t += rng.normal(0, noise)
            p = d * t
            pts.append([p[0], p[1], p[2], float(r) + c / 10000.0])
            rg.append(t)
            col.append(c)
            count += 1
        er.append(count - 1 - 5)
    return (np.array(pts, np.float32), np.array(rg, np.float32), np.array(col, np.int32),
            np.array(sr, np.int32), np.array(er, np.int32))
