"""Cross-view association geometry of the reference's step 2 on the GPU
(src/pipeline/step2_crossviewmatching.py, src/utils/multicam_toolbox.py).

Drop-in functions (same names, argument meaning and return shapes):
  geometry_affinity2(points_set, dimGroup, config_path, camparam)   step2:373-432
  matchSVT(S, dimGroup, *, alpha, _lambda, dual_stochastic_SVT, ...)  step2:130-216
  calc_3dpose(kp_2d, config_path, camparam)                           step2:436-461
  undistortPoints(config_path, pos_2d, omnidir, camparam)             mct:393-431
  triangulatePoints(config_path, pos_2d_undist, frame_use, use_optim_extrin, camparam)  mct:433-486
plus batched forms over many frames (`*_batch`), which is how the kernels are meant to be
used (the reference calls them once per keyframe in Python loops, step2:899-928).

``camparam`` is the dict of step2.get_camparam (:67-75): camera_id, K, xi, D, rvecs, tvecs,
pmat.  Reading it from YAML/HDF5 (config_path with camparam=None) is the reference's file
I/O and is not reimplemented: pass camparam.
"""
import ctypes

import numpy as np

from . import _lib
from .cameras import (CameraGroup, OmnidirCamera, _default_device, _is_torch, _ptr, _ret, _stream, _to_dev,
                      torch)

THR_KP = 0.1   # step2_crossviewmatching.py:21


def _need_camparam(camparam):
    if camparam is None:
        raise NotImplementedError(
            "camparam=None (read cam_intrinsic.h5 / cam_extrinsic_optim.h5 through config_path, "
            "step2_crossviewmatching.py:35-75) is file I/O outside the accelerated path: pass camparam")


_group_cache = {}


def group_from_camparam(camparam):
    """CameraGroup of OmnidirCamera objects for a step2 camparam dict (cached by identity of
    the parameter values)."""
    _need_camparam(camparam)
    key = tuple(np.asarray(camparam[k][i], dtype=np.float64).tobytes()
                for k in ("K", "xi", "D", "rvecs", "tvecs") for i in range(len(camparam["K"])))
    cg = _group_cache.get(key)
    if cg is None:
        cams = []
        for i in range(len(camparam["K"])):
            cams.append(OmnidirCamera(K=np.asarray(camparam["K"][i], dtype=np.float64),
                                      xi=np.asarray(camparam["xi"][i], dtype=np.float64).ravel()[:1],
                                      D=np.asarray(camparam["D"][i], dtype=np.float64).ravel(),
                                      rvec=np.asarray(camparam["rvecs"][i], dtype=np.float64).ravel(),
                                      tvec=np.asarray(camparam["tvecs"][i], dtype=np.float64).ravel(),
                                      name=str(camparam["camera_id"][i]) if "camera_id" in camparam else str(i)))
        cg = CameraGroup(cams)
        if len(_group_cache) > 16:
            _group_cache.clear()
        _group_cache[key] = cg
    return cg


# ------------------------------------------------------------------------------------------
# batched kernels
# ------------------------------------------------------------------------------------------

def geometry_affinity_batch(cgroup, kp, dim, thr_kp=THR_KP, return_dist=False):
    """kp (F,M,J,3) undistorted x, y, score; dim (F,C+1) int32 cumulative detection counts per
    camera (rows may describe fewer than M detections: the rest is padding).
    Returns aff (F,M,M) [and the mean ray distance matrix (F,M,M)]."""
    device = kp.device.index if _is_torch(kp) and kp.device.type == "cuda" else cgroup._dev()
    like_torch = _is_torch(kp)
    rig = cgroup._rig(device)
    k = _to_dev(kp, device)
    F, M, J, _ = k.shape
    d = dim if _is_torch(dim) else torch.from_numpy(np.ascontiguousarray(dim, dtype=np.int32))
    d = d.to(device="cuda:%d" % device, dtype=torch.int32).contiguous()
    assert tuple(d.shape) == (F, len(cgroup.cameras) + 1), "dim must be (F, C+1)"
    aff = torch.empty((F, M, M), dtype=torch.float64, device=k.device)
    dist = torch.empty((F, M, M), dtype=torch.float64, device=k.device) if return_dist else None
    _lib.check(rig._lib.m3d_ray_affinity(rig.handle, _ptr(k), ctypes.c_void_p(d.data_ptr()), F, M, J,
                                         float(thr_kp), _ptr(aff), _ptr(dist), _stream(device)),
               "m3d_ray_affinity")
    if return_dist:
        return _ret(aff, like_torch), _ret(dist, like_torch)
    return _ret(aff, like_torch)


def match_svt_batch(W, dim, n_cams, alpha=0.1, _lambda=50.0, mu=64.0, tol=5e-4, maxIter=500,
                    return_iters=False, device=None):
    """W (F,M,M) affinities, dim (F,C+1) -> match (F,M,M) uint8 (matchSVT with pselect = 1,
    dual_stochastic_SVT = False, per frame)."""
    like_torch = _is_torch(W)
    if device is None:
        device = W.device.index if like_torch and W.device.type == "cuda" else _default_device()
    lib = _lib.require_gpu()
    w = _to_dev(W, device)
    F, M, _ = w.shape
    d = dim if _is_torch(dim) else torch.from_numpy(np.ascontiguousarray(dim, dtype=np.int32))
    d = d.to(device="cuda:%d" % device, dtype=torch.int32).contiguous()
    out = torch.empty((F, M, M), dtype=torch.uint8, device=w.device)
    its = torch.empty((F,), dtype=torch.int32, device=w.device)
    _lib.check(lib.m3d_match_svt(_ptr(w), ctypes.c_void_p(d.data_ptr()), F, M, int(n_cams), float(alpha),
                                 float(_lambda), float(mu), float(tol), int(maxIter),
                                 ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(its.data_ptr()), int(device),
                                 _stream(device)), "m3d_match_svt")
    if return_iters:
        return _ret(out, like_torch), _ret(its, like_torch)
    return _ret(out, like_torch)


def triangulate_ls_batch(cgroup, xy_undist, use):
    """Inhomogeneous least squares of mct.triangulatePoints on (C,N,2) undistorted points with
    use (C,N) bool/uint8 -> (N,3)."""
    device, like_torch = cgroup._device_of(xy_undist)
    rig = cgroup._rig(device)
    x = _to_dev(xy_undist, device)
    C, N, _ = x.shape
    u = use if _is_torch(use) else torch.from_numpy(np.ascontiguousarray(use).astype(np.uint8))
    u = u.to(device=x.device, dtype=torch.uint8).contiguous()
    out = torch.empty((N, 3), dtype=torch.float64, device=x.device)
    _lib.check(rig._lib.m3d_triangulate_ls(rig.handle, _ptr(x), ctypes.c_void_p(u.data_ptr()), N, _ptr(out),
                                           _stream(device)), "m3d_triangulate_ls")
    return _ret(out, like_torch)


# ------------------------------------------------------------------------------------------
# reference-shaped entry points
# ------------------------------------------------------------------------------------------

def geometry_affinity2(points_set, dimGroup, config_path=None, camparam=None):
    """step2_crossviewmatching.py:373-432 — (M,J,3) undistorted keypoints + scores and the
    cumulative per-camera detection counts -> (M,M) affinity."""
    cg = group_from_camparam(camparam)
    pts = np.asarray(points_set, dtype=np.float64)
    return geometry_affinity_batch(cg, pts[None], np.asarray(dimGroup, dtype=np.int32)[None])[0]


def matchSVT(S, dimGroup, *, alpha=0.1, pselect=1, tol=5e-4, maxIter=500, verbose=False,
             eigenvalues=False, _lambda=50, mu=64, dual_stochastic_SVT=True):
    """step2_crossviewmatching.py:130-216.  The reference's only call site uses
    dual_stochastic_SVT=False, pselect=1 (step2:589-595); other settings are not on the path."""
    if dual_stochastic_SVT or pselect != 1 or eigenvalues:
        raise NotImplementedError("matchSVT: only pselect=1, dual_stochastic_SVT=False, eigenvalues=False "
                                  "(the configuration of step2_crossviewmatching.py:589-595) is accelerated")
    S = np.asarray(S, dtype=np.float64)
    dg = np.asarray(dimGroup, dtype=np.int32)
    return match_svt_batch(S[None], dg[None], len(dg) - 1, alpha=alpha, _lambda=_lambda, mu=mu, tol=tol,
                           maxIter=maxIter)[0]


def undistortPoints(config_path, pos_2d, omnidir=False, camparam=None):
    """multicam_toolbox.py:393-431: list of (n,2) pixel arrays per camera -> list of (n,2)
    undistorted arrays (omnidir model; the pinhole branch reads mtx/dist from HDF5 only)."""
    _need_camparam(camparam)
    if not omnidir:
        raise NotImplementedError("undistortPoints(omnidir=False) reads mtx/dist from cam_intrinsic.h5 "
                                  "(multicam_toolbox.py:422-429); use cameras.Camera.undistort_points")
    cg = group_from_camparam(camparam)
    out = []
    for cam, p in zip(cg.cameras, pos_2d):
        p = np.asarray(p, dtype=np.float64) + 0.0
        out.append(np.squeeze(cam.undistort_points(p.reshape(1, -1, 2))))
    return out


def triangulatePoints(config_path, pos_2d_undist, frame_use, use_optim_extrin=True, camparam=None):
    """multicam_toolbox.py:433-486: per-camera list of (n,2) undistorted points, frame_use (n,C)
    -> (n,3) by X = -pinv(A[:, :3]) @ A[:, 3]."""
    cg = group_from_camparam(camparam)
    xy = np.stack([np.asarray(p, dtype=np.float64).reshape(-1, 2) for p in pos_2d_undist])
    use = np.ascontiguousarray(np.asarray(frame_use, dtype=bool).T)
    return triangulate_ls_batch(cg, xy, use)


def calc_3dpose(kp_2d, config_path=None, camparam=None, thr_kp=THR_KP):
    """step2_crossviewmatching.py:436-461 (thr 0.1) / step3_crossframematching.py:254-272
    (thr 0.3): kp_2d (C,J,3) raw pixels + score -> (J,3)."""
    cg = group_from_camparam(camparam)
    kp = np.asarray(kp_2d, dtype=np.float64)
    und = cg.undistort_points(np.ascontiguousarray(kp[:, :, :2]))
    use = ~(np.isnan(kp[:, :, 0]) | (kp[:, :, 2] < thr_kp))
    return triangulate_ls_batch(cg, und, use)
