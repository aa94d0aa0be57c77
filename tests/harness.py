"""Test-only ctypes wrapper of tests/host_harness.cpp (the kernels' per-point arithmetic
compiled for the host).  Never imported by the product."""
import ctypes
import os
import subprocess

import numpy as np

from macaque_3d_pose_estimation_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
SO = os.path.join(BUILD, "libm3d_hostmath.so")
SRC = os.path.join(HERE, "host_harness.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "macaque_3d_pose_estimation_b200", "csrc")

_h = None


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("m3d_math.cuh", "m3d_point.cuh", "m3d_rig.h", "m3d_cert.h", "m3d_ransac_cert.cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    global _h
    if _h is not None:
        return _h
    if _stale():
        os.makedirs(BUILD, exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    _h = ctypes.CDLL(SO)
    _h.hh_last_error.restype = ctypes.c_char_p
    return _h


def cam_structs(cams):
    """M3DCam array from oracle CamSpec objects."""
    arr = (_lib.M3DCam * max(len(cams), 1))()
    for i, c in enumerate(cams):
        s = arr[i]
        s.model = c.model
        s.n_dist = c.dist.size
        for j in range(9):
            s.K[j] = float(c.K.flat[j])
        for j in range(14):
            s.dist[j] = float(c.dist[j]) if j < c.dist.size else 0.0
        for j in range(3):
            s.rvec[j] = float(c.rvec[j])
            s.tvec[j] = float(c.tvec[j])
        s.xi = float(c.xi)
    return arr


def _p(a):
    return ctypes.c_void_p(a.ctypes.data if a is not None else 0)


def _ok(rc):
    if rc != 0:
        raise RuntimeError(load().hh_last_error().decode())


def extrinsics(cams):
    M = np.empty((len(cams), 4, 4))
    _ok(load().hh_extrinsics(cam_structs(cams), len(cams), _p(M)))
    return M


def undistort(cams, xy):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    out = np.empty_like(xy)
    _ok(load().hh_undistort(cam_structs(cams), len(cams), _p(xy), ctypes.c_int64(xy.shape[1]), _p(out)))
    return out


def project(cams, p3d):
    p3d = np.ascontiguousarray(p3d, dtype=np.float64).reshape(-1, 3)
    out = np.empty((len(cams), p3d.shape[0], 2))
    _ok(load().hh_project(cam_structs(cams), len(cams), _p(p3d), ctypes.c_int64(p3d.shape[0]), _p(out)))
    return out


def triangulate_error(cams, xy, undistort=True):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    n = xy.shape[1]
    p3d = np.empty((n, 3))
    err = np.empty(n)
    _ok(load().hh_triangulate_error(cam_structs(cams), len(cams), _p(xy), ctypes.c_int64(n),
                                    int(undistort), _p(p3d), _p(err)))
    return p3d, err


def ransac(cams, xy, undistort=True, min_cams=2, threshold=0.5, init_best=200.0):
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    C, n = len(cams), xy.shape[1]
    p3d = np.empty((n, 3))
    picked = np.empty((C, n, 1), dtype=np.uint8)
    xyp = np.empty((C, n, 2))
    err = np.empty(n)
    sub = np.empty(n, dtype=np.int32)
    nev = np.empty(n, dtype=np.int32)
    _ok(load().hh_ransac(cam_structs(cams), C, _p(xy), ctypes.c_int64(n), int(undistort), int(min_cams),
                         ctypes.c_double(threshold), ctypes.c_double(init_best), _p(p3d), _p(picked),
                         _p(xyp), _p(err), _p(sub), _p(nev)))
    return p3d, picked.view(np.bool_), xyp, err, sub, nev


def cert_tables(cams):
    """Pair-certificate tables of csrc/m3d_cert.h: (ok_mask, inv_mf (C,), E (pairs, 10))."""
    C = len(cams)
    ok = ctypes.c_int32(0)
    inv = np.zeros(max(C, 1))
    E = np.zeros((max(C * (C - 1) // 2, 1), 10))
    _ok(load().hh_cert(cam_structs(cams), C, ctypes.byref(ok), _p(inv), _p(E)))
    return ok.value, inv[:C], E[:C * (C - 1) // 2]


def ransac_cert(cams, xy, undistort=True, min_cams=2, threshold=0.5, init_best=200.0, use_cert=True,
                return_solved=False, general=False):
    """The pruned subset search (csrc/m3d_ransac_cert.cuh ransac_cert_point) on the host.  8-camera pinhole
    rigs run the compile-time-count instantiation of the headline kernels (straight-line undistortion and half
    budgets); ``general=True`` forces the run-time-count form every other rig takes."""
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    C, n = len(cams), xy.shape[1]
    p3d = np.empty((n, 3))
    picked = np.empty((C, n, 1), dtype=np.uint8)
    xyp = np.empty((C, n, 2))
    err = np.empty(n)
    sub = np.empty(n, dtype=np.int32)
    nev = np.empty(n, dtype=np.int32)
    solved = np.empty(n, dtype=np.int32)
    _ok(load().hh_ransac_cert(cam_structs(cams), C, _p(xy), ctypes.c_int64(n), int(undistort), int(min_cams),
                              ctypes.c_double(threshold), ctypes.c_double(init_best), int(bool(use_cert)) | (2 if general else 0), _p(p3d),
                              _p(picked), _p(xyp), _p(err), _p(sub), _p(nev), _p(solved)))
    res = (p3d, picked.view(np.bool_), xyp, err, sub, nev)
    return res + (solved,) if return_solved else res
