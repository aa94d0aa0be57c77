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


# ------------------------------------------------------------------------------------------
# keyframe association (MultiEstimator.predict_data, step2_crossviewmatching.py:502-713)
# ------------------------------------------------------------------------------------------

ALPHA_ID = 0.2          # step2:22
MODEL_CFG = {"joint_num": 17, "alpha_SVT": 0.5, "lambda_SVT": 50, "dual_stochastic_SVT": False}   # step2:25-31


def reproject(i_cam, p3d, camparam=None, config_path=""):
    """step2_crossviewmatching.py:465-489: project (n,3) points into camera i_cam (omnidir model)."""
    cg = group_from_camparam(camparam)
    return cg.cameras[i_cam].project(np.asarray(p3d, dtype=np.float64)).reshape(-1, 2)


class MultiEstimator:
    """Matching and 3D reconstruction across cameras for a single keyframe — the host logic of
    step2_crossviewmatching.py:494-713 with every geometric step on the GPU kernels
    (ray affinity, SVT association, LS triangulation, omnidir reprojection).  Drawing
    (``show=True``) and the unused spectral initialisation (:577-586) are not reproduced."""

    def __init__(self, cfg=None, debug=False):
        self.cfg = cfg
        self.debug = debug

    def predict_data(self, info_dict, show=False, plt_id=0, camparam=None, bcomb_prev=None):
        import itertools
        _need_camparam(camparam)
        if show:
            raise NotImplementedError("predict_data(show=True) draws with matplotlib (step2:648-693)")
        n_cam = len(info_dict)
        dimGroup = [0]
        cnt = 0
        for cam_id in range(n_cam):
            cnt += len(info_dict[cam_id][0])
            dimGroup.append(cnt)
        dimGroup = np.array(dimGroup)
        info_list = []
        for cam_id in range(n_cam):
            info_list.extend(info_dict[cam_id][0])
        if not info_list:
            return [], [], []
        M = len(info_list)
        n_kp = MODEL_CFG["joint_num"]
        pose2d = np.array([det["pose2d"] for det in info_list]).reshape(M, n_kp, 2)
        pose_score = np.array([det["pose2d_raw"] for det in info_list]).reshape(M, n_kp, 3)[..., 2]
        kp_mat = np.concatenate([pose2d, pose_score[..., np.newaxis]], axis=2)
        sub2cam = np.zeros(M, dtype=int)
        for idx in range(len(dimGroup) - 1):
            sub2cam[dimGroup[idx]:dimGroup[idx + 1]] = idx
        cid_list = [det["cid"] for det in info_list]

        geo_aff = geometry_affinity2(kp_mat.copy(), dimGroup, self.cfg, camparam=camparam)      # :554
        cid = np.asarray(cid_list)
        cid_mat = ((sub2cam[:, None] != sub2cam[None, :]) & (cid[:, None] >= 0) &
                   (cid[:, None] == cid[None, :])).astype(np.float64)                          # :557-561
        W = ALPHA_ID * cid_mat + (1 - ALPHA_ID) * geo_aff                                      # :572-575
        W *= (geo_aff > 0)
        W = np.nan_to_num(W)
        match_mat = matchSVT(W, dimGroup, alpha=MODEL_CFG["alpha_SVT"], _lambda=MODEL_CFG["lambda_SVT"],
                             dual_stochastic_SVT=MODEL_CFG["dual_stochastic_SVT"])              # :589-595
        col_sums = match_mat.sum(axis=0)                                                       # :598-607
        matched_cols = np.nonzero(col_sums > 1.9)[0]
        bin_match = match_mat[:, matched_cols] > 0.9
        matched_list = [[] for _ in range(bin_match.shape[1])]
        for sub_idx, row in enumerate(bin_match):
            if row.sum() != 0:
                matched_list[row.argmax()].append(sub_idx)
        matched_list = [np.array(lst) for lst in matched_list]

        def get_best_comb(person_idxs):                                                        # :610-646
            person_idxs = np.asarray(person_idxs, dtype=int)
            cam_ids = sub2cam[person_idxs]
            cam_groups = [person_idxs[np.where(cam_ids == c)].tolist() or [None] for c in range(n_cam)]
            combos = list(itertools.product(*cam_groups))
            if len(combos) == 1:
                return person_idxs
            errors = []
            for combo in combos:
                kp2d = np.zeros((n_cam, n_kp, 3))
                for c, sub_idx in enumerate(combo):
                    if sub_idx is not None:
                        kp2d[c] = info_list[sub_idx]["pose2d_raw"]
                p3d = calc_3dpose(kp2d, self.cfg, camparam=camparam)
                derrs = []
                for c, sub_idx in enumerate(combo):
                    if sub_idx is None:
                        continue
                    raw = np.asarray(info_list[sub_idx]["pose2d_raw"])
                    ok = raw[:, 2] > THR_KP
                    derrs.append(raw[ok, :2] - reproject(c, p3d, camparam=camparam)[ok])
                errors.append(np.sqrt((np.vstack(derrs) ** 2).mean()) if derrs else np.inf)
            best = combos[int(np.argmin(errors))]
            return np.array([i for i in best if i is not None], dtype=int)

        refined = []
        for person in matched_list:                                                            # :649-657
            best = get_best_comb(person)
            refined.append(best)
            leftover = set(person.tolist()) - set(best.tolist())
            if len(leftover) > 1:
                refined.append(get_best_comb(np.array(list(leftover), dtype=int)))
        P3d_list, matched_list2, bcomb_list = [], [], []
        for person_idxs in refined:                                                            # :696-713
            if person_idxs.shape[0] < 2:
                continue
            kp2d = np.zeros((n_cam, n_kp, 3))
            for sub_idx in person_idxs:
                kp2d[sub2cam[sub_idx]] = info_list[sub_idx]["pose2d_raw"]
            P3d_list.append(calc_3dpose(kp2d, self.cfg, camparam=camparam))
            bcomb = -np.ones(n_cam, dtype=int)
            for sub_idx in person_idxs:
                bcomb[sub2cam[sub_idx]] = info_list[sub_idx]["bbox_id"][1]
            matched_list2.append(person_idxs)
            bcomb_list.append(bcomb)
        return matched_list2, P3d_list, bcomb_list


# ------------------------------------------------------------------------------------------
# step 3 users of the same arithmetic (step3_crossframematching.py:254-302)
# ------------------------------------------------------------------------------------------

def calc_3dpose_batch(kp_2d, camparam, thr_kp=0.3):
    """All frames of a tracklet in one launch: kp_2d (F,C,J,3) raw pixels + score (NaN rows for
    cameras that do not see the animal) -> (F,J,3).  Per frame this is step3's calc_3dpose
    (:254-272, score gate 0.3) = omnidir undistortion + mct.triangulatePoints."""
    cg = group_from_camparam(camparam)
    kp = np.asarray(kp_2d, dtype=np.float64)
    F, C, J, _ = kp.shape
    flat = np.ascontiguousarray(kp.transpose(1, 0, 2, 3)).reshape(C, F * J, 3)
    und = cg.undistort_points(np.ascontiguousarray(flat[:, :, :2]))
    with np.errstate(invalid="ignore"):
        use = ~(np.isnan(flat[:, :, 0]) | (flat[:, :, 2] < thr_kp))
    return triangulate_ls_batch(cg, np.nan_to_num(und), use).reshape(F, J, 3)


def calc_3dtrace(p2d, camparam, thr_kp=0.3):
    """step3's calc_3dtrace (:274-302) on the (F,C,J,3) array it assembles per tracklet: frames
    seen by fewer than two cameras give NaN; returns the per-frame nanmedian over keypoints (F,3)."""
    import warnings
    kp = np.asarray(p2d, dtype=np.float64)
    p3d = calc_3dpose_batch(kp, camparam, thr_kp)
    seen = (~np.isnan(kp[:, :, :, 0]).all(axis=2)).sum(axis=1)
    p3d[seen < 2] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmedian(p3d, axis=1)
