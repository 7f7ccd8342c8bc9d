"""ctypes binding of liblvreg_host.so: synthetic generator + the C++ mapOptimization mirror."""
import ctypes as C
import os

import numpy as np

from .binding import Params, Result, Timings, lib as _cuda_lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liblvreg_host.so")
_lib = None

MID360, BEAM128 = 0, 1


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError("liblvreg_host.so is missing: run __graft_entry__.build()")
        _cuda_lib()                       # liblvreg.so first (rpath would find it too)
        L = C.CDLL(_PATH)
        L.lvh_last_error.restype = C.c_char_p
        L.lvh_gen_create.restype = C.c_void_p
        L.lvh_gen_create.argtypes = [C.c_int, C.c_uint64]
        L.lvh_gen_destroy.argtypes = [C.c_void_p]
        L.lvh_gen_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.lvh_gen_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lvh_truth_pose.argtypes = [C.c_int, C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.lvh_guess_pose.argtypes = [C.c_uint64, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.lvh_mo_create.restype = C.c_void_p
        L.lvh_mo_create.argtypes = [C.c_void_p, C.c_int]
        L.lvh_mo_destroy.argtypes = [C.c_void_p]
        L.lvh_mo_handle.restype = C.c_void_p
        L.lvh_mo_handle.argtypes = [C.c_void_p]
        L.lvh_mo_selection.restype = C.c_size_t
        L.lvh_mo_selection.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


class Generator:
    """Deterministic synthetic world + scans (host/synth.cpp)."""

    def __init__(self, sensor, seed):
        self.L = lib()
        self.sensor = sensor
        self.seed = seed
        self.g = C.c_void_p(self.L.lvh_gen_create(sensor, seed))

    def scan(self, pose, seed, threads=8):
        pose = np.ascontiguousarray(pose, np.float32)
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        self.L.lvh_gen_scan(self.g, pose.ctypes.data_as(C.c_void_p), seed, threads, C.byref(nc), C.byref(ns))
        corner = np.zeros((nc.value, 4), np.float32)
        surf = np.zeros((ns.value, 4), np.float32)
        self.L.lvh_gen_fetch(self.g, corner.ctypes.data_as(C.c_void_p), surf.ctypes.data_as(C.c_void_p))
        return corner, surf

    def truth_pose(self, k, scan_period=0.2, speed=1.0):
        pose = np.zeros(6, np.float32)
        self.L.lvh_truth_pose(self.sensor, self.seed, scan_period, speed, k, pose.ctypes.data_as(C.c_void_p))
        return pose

    def guess_pose(self, k, truth, guess_trans=0.10, guess_rot=0.035):
        truth = np.ascontiguousarray(truth, np.float32)
        g = np.zeros(6, np.float32)
        self.L.lvh_guess_pose(self.seed, guess_trans, guess_rot, k, truth.ctypes.data_as(C.c_void_p),
                              g.ctypes.data_as(C.c_void_p))
        return g

    def __del__(self):
        if getattr(self, "g", None):
            self.L.lvh_gen_destroy(self.g)
            self.g = None


class MapOptimizationMirror:
    """The C++ mirror of the reference's mapOptimization class (host/map_optimization.cpp)."""

    def __init__(self, params=None, device=0):
        self.L = lib()
        self.mo = C.c_void_p(self.L.lvh_mo_create(C.byref(params) if params is not None else None, device))
        if not self.mo:
            raise RuntimeError(self.L.lvh_last_error().decode())

    def handle_scan(self, corner, surf, stamp, guess):
        corner = np.ascontiguousarray(corner, np.float32)
        surf = np.ascontiguousarray(surf, np.float32)
        guess = np.ascontiguousarray(guess, np.float32)
        pose = np.zeros(6, np.float32)
        res, tim, nk = Result(), Timings(), C.c_int(0)
        st = self.L.lvh_mo_handle_scan(self.mo, corner.ctypes.data_as(C.c_void_p), C.c_size_t(len(corner)),
                                       surf.ctypes.data_as(C.c_void_p), C.c_size_t(len(surf)), C.c_double(stamp),
                                       guess.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p),
                                       C.byref(res), C.byref(tim), C.byref(nk))
        if st == -2:
            raise RuntimeError(self.L.lvh_last_error().decode())
        return st, pose, res, tim, nk.value

    def raw_scan_to_pose(self, raw, n_scan, horizon, sensor, guess, ids=None, min_range=0.5, max_range=1000.0,
                         deskew=False, time_scan_cur=0.0, imu_time=None, imu_rot=None, edge_threshold=1.0):
        """ImageProjection + FeatureExtraction + registration mirrors on one handle: raw Livox-layout points
        (uint8 [n,32]) in, pose out -> (status, pose, Result, (n_extracted, n_corner, n_surf))"""
        raw = np.ascontiguousarray(raw, np.uint8)
        pose = np.ascontiguousarray(guess, np.float32).copy()
        res = Result()
        n_out = (C.c_size_t * 3)()
        if deskew:
            t = np.ascontiguousarray(imu_time, np.float64)
            r = np.ascontiguousarray(imu_rot, np.float64).reshape(-1, 3)
            tp, rp, k = t.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), len(t)
        else:
            tp, rp, k = None, None, 0
        if ids is None:
            idp, nid = None, 0
        else:
            ids = np.ascontiguousarray(ids, np.int32)
            idp, nid = ids.ctypes.data_as(C.c_void_p), len(ids)
        st = self.L.lvh_mo_raw_scan_to_pose(self.mo, raw.ctypes.data_as(C.c_void_p), C.c_size_t(len(raw)), int(n_scan),
                                            int(horizon), int(sensor), C.c_float(min_range), C.c_float(max_range),
                                            int(bool(deskew)), C.c_double(time_scan_cur), tp, rp, int(k),
                                            C.c_float(edge_threshold), idp, C.c_size_t(nid),
                                            pose.ctypes.data_as(C.c_void_p), C.byref(res), n_out)
        if st == -2:
            raise RuntimeError(self.L.lvh_last_error().decode())
        return st, pose, res, tuple(int(v) for v in n_out)

    def perform_loop_closure(self):
        """mapOptimization::performLoopClosure -> (queued, key_cur, key_pre, LoopResult)"""
        from .binding import LoopResult
        cur, pre, out = C.c_int(-1), C.c_int(-1), LoopResult()
        st = self.L.lvh_mo_perform_loop_closure(self.mo, C.byref(cur), C.byref(pre), C.byref(out))
        if st == -2:
            raise RuntimeError(self.L.lvh_last_error().decode())
        return bool(st), cur.value, pre.value, out

    def selection(self):
        ids = np.zeros(4096, np.int32)
        n = self.L.lvh_mo_selection(self.mo, ids.ctypes.data_as(C.c_void_p), C.c_size_t(len(ids)))
        return ids[:n].copy()

    def close(self):
        if getattr(self, "mo", None):
            self.L.lvh_mo_destroy(self.mo)
            self.mo = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReplayMirror:
    """one long-lived C++ mapOptimization mirror with pre-sized device buffers; sequences are replayed on it one
    after the other (reset in between), as lvreg_replay does per GPU"""
    KEYS = ["scans", "registered", "keyframes", "converged", "iterations", "queries", "wall_s", "device_ms",
            "max_pos_err", "max_rot_err", "launches"]

    def __init__(self, sensor, device=0):
        self.L = lib()
        self.sensor = sensor
        if sensor == BEAM128:
            rv = (3000000, 12000000, 65536, 400000, 1 << 25)
        else:
            rv = (200000, 800000, 20000, 60000, 1 << 22)
        arr = (C.c_size_t * 5)(*rv)
        self.L.lvh_mo_create_reserved.restype = C.c_void_p
        self.L.lvh_mo_create_reserved.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.L.lvh_replay_on.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_float,
                                         C.c_float, C.c_int, C.c_void_p]
        self.mo = C.c_void_p(self.L.lvh_mo_create_reserved(None, device, arr))
        if not self.mo:
            raise RuntimeError(self.L.lvh_last_error().decode())

    def replay(self, seed, n_scans, period=0.2, speed=1.0, guess_trans=0.10, guess_rot=0.035, gen_threads=8):
        out = (C.c_double * 11)()
        st = self.L.lvh_replay_on(self.mo, self.sensor, C.c_uint64(seed), n_scans, C.c_double(period), C.c_double(speed),
                                  C.c_float(guess_trans), C.c_float(guess_rot), gen_threads, out)
        if st != 0:
            raise RuntimeError(self.L.lvh_last_error().decode())
        return dict(zip(self.KEYS, list(out)))

    def close(self):
        if self.mo:
            self.L.lvh_mo_destroy(self.mo)
            self.mo = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def replay(sensor, seed, n_scans, device=0, period=0.2, speed=1.0, guess_trans=0.10, guess_rot=0.035, gen_threads=8):
    out = (C.c_double * 10)()
    L = lib()
    st = L.lvh_replay(sensor, C.c_uint64(seed), n_scans, C.c_double(period), C.c_double(speed), C.c_float(guess_trans),
                      C.c_float(guess_rot), device, gen_threads, out)
    if st != 0:
        raise RuntimeError(L.lvh_last_error().decode())
    keys = ["scans", "registered", "keyframes", "converged", "iterations", "queries", "wall_s", "device_ms",
            "max_pos_err", "max_rot_err"]
    return dict(zip(keys, list(out)))
