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
    undistorted arrays.  omnidir=True: the Mei model from camparam's K / xi / D (:404-420).
    omnidir=False: cv2.undistortPoints with the pinhole ``mtx`` / ``dist`` of every camera (:421-429;
    the reference reads them from cam_intrinsic.h5 — here they are the entries ``camparam['mtx']``,
    ``camparam['dist']``, as ``calib_io.read_camparam`` fills them)."""
    _need_camparam(camparam)
    if omnidir:
        cams = group_from_camparam(camparam).cameras
    else:
        if "mtx" not in camparam or "dist" not in camparam:
            raise KeyError("undistortPoints(omnidir=False) needs camparam['mtx'] and camparam['dist'] "
                           "(multicam_toolbox.py:423-425 reads them from cam_intrinsic.h5)")
        from .cameras import Camera
        cams = [Camera(matrix=np.asarray(m, dtype=np.float64), dist=np.asarray(d, dtype=np.float64).ravel(),
                       name=str(i)) for i, (m, d) in enumerate(zip(camparam["mtx"], camparam["dist"]))]
    out = []
    for cam, p in zip(cams, pos_2d):
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
# keyframe association, batched over keyframes (MultiEstimator.predict_data, step2:502-713)
# ------------------------------------------------------------------------------------------

ALPHA_ID = 0.2          # step2:22
MODEL_CFG = {"joint_num": 17, "alpha_SVT": 0.5, "lambda_SVT": 50, "dual_stochastic_SVT": False}   # step2:25-31


def reproject(i_cam, p3d, camparam=None, config_path=""):
    """step2_crossviewmatching.py:465-489: project (n,3) points into camera i_cam (omnidir model)."""
    cg = group_from_camparam(camparam)
    return cg.cameras[i_cam].project(np.asarray(p3d, dtype=np.float64)).reshape(-1, 2)


def _cam_of(dim, M):
    """Camera index of every detection slot from the cumulative counts: dim (F,C+1) -> (F,M); padding
    slots (m >= dim[:, C]) get C."""
    return (np.arange(M)[None, :, None] >= dim[:, None, 1:]).sum(axis=2)


def _ls_persons(cgroup, und_sel, score_sel, thr_kp):
    """und_sel (P, C, J, 2), score_sel (P, C, J) -> (P, J, 3): calc_3dpose (step2:436-461) for P persons in
    one launch (a camera without a member has score 0 everywhere)."""
    P, C, J, _ = und_sel.shape
    xy = np.ascontiguousarray(und_sel.transpose(1, 0, 2, 3)).reshape(C, P * J, 2)
    sc = np.ascontiguousarray(score_sel.transpose(1, 0, 2)).reshape(C, P * J)
    with np.errstate(invalid="ignore"):
        use = ~(np.isnan(xy[:, :, 0]) | (sc < thr_kp) | np.isnan(sc))
    return triangulate_ls_batch(cgroup, np.nan_to_num(xy), use).reshape(P, J, 3)


def _combo_rmse(cgroup, kp_raw_sel, p3d, thr_kp):
    """Reprojection RMSE of get_best_comb (step2:626-641): kp_raw_sel (P, C, J, 3) raw pixels + score (score
    0 = camera not in the combination), p3d (P, J, 3) -> (P,) over the keypoints with score > thr."""
    P, C, J, _ = kp_raw_sel.shape
    proj = cgroup.project(p3d.reshape(-1, 3)).reshape(C, P, J, 2).transpose(1, 0, 2, 3)
    keep = kp_raw_sel[..., 2] > thr_kp
    d2 = ((kp_raw_sel[..., :2] - proj) ** 2).sum(axis=-1)
    n = 2.0 * keep.sum(axis=(1, 2))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(n > 0, np.sqrt(np.where(keep, d2, 0.0).sum(axis=(1, 2)) / np.maximum(n, 1)), np.inf)


def _resolve_duplicate_frames(cgroup, frames, label, cam_of, kp_raw, und, thr_kp):
    """Host resolution of the keyframes in which a cluster holds two detections of one camera
    (get_best_comb, step2_crossviewmatching.py:610-657): all candidate combinations of all such clusters are
    scored in one batch (LS triangulation + reprojection RMSE), then the leftover round.  ``label`` /
    ``cam_of`` / ``kp_raw`` / ``und`` are the rows of ``frames`` only.  Returns [(frame, (column, 0 | 1),
    member row)] for every cluster of these frames."""
    import itertools
    n, M = label.shape
    C = len(cgroup.cameras)
    fi, mi = np.nonzero(label >= 0)
    key = fi.astype(np.int64) * M + label[fi, mi]
    order = np.lexsort((mi, key))
    fi, mi, key = fi[order], mi[order], key[order]
    ci = cam_of[fi, mi]
    ukey, start = np.unique(key, return_index=True)
    cnt = np.diff(np.append(start, key.size))
    members = -np.ones((ukey.size, C), dtype=np.int64)
    cl = np.repeat(np.arange(ukey.size), cnt)
    first = np.ones(key.size, dtype=bool)
    pair = cl * C + ci
    first[1:] = pair[1:] != pair[:-1]
    members[cl[first], ci[first]] = mi[first]
    dup_clusters = np.unique(cl[~first])
    persons = []
    for k in np.setdiff1d(np.arange(ukey.size), dup_clusters):
        persons.append((int(frames[ukey[k] // M]), (int(ukey[k] % M), 0), members[k]))

    def score_round(groups):
        combos, owner = [], []
        for gi, (f, dets) in enumerate(groups):
            per_cam = [[m for m in dets if cam_of[f, m] == c] or [-1] for c in range(C)]
            for combo in itertools.product(*per_cam):
                combos.append(combo)
                owner.append(gi)
        combos = np.array(combos, dtype=np.int64)
        owner = np.array(owner)
        fr = np.array([groups[g][0] for g in owner])
        has = combos >= 0
        safe = np.where(has, combos, 0)
        raw_sel = kp_raw[fr[:, None], safe] * has[..., None, None]
        und_sel = np.where(has[..., None, None], und[fr[:, None], safe], 0.0)
        p3d = _ls_persons(cgroup, und_sel, raw_sel[..., 2], thr_kp)
        err = _combo_rmse(cgroup, raw_sel, p3d, thr_kp)
        best = []
        for gi in range(len(groups)):
            idx = np.nonzero(owner == gi)[0]
            best.append(combos[idx[int(np.argmin(err[idx]))]])              # first minimum, like np.argmin
        return best

    if dup_clusters.size:
        groups = [(int(ukey[k] // M), mi[cl == k].tolist()) for k in dup_clusters]
        best = score_round(groups)
        left_groups, left_of = [], []
        for gi, k in enumerate(dup_clusters):
            f, dets = groups[gi]
            persons.append((int(frames[f]), (int(ukey[k] % M), 0), best[gi]))
            rest = sorted(set(dets) - set(int(m) for m in best[gi] if m >= 0))
            if len(rest) > 1:                                               # step2:653-657
                left_groups.append((f, rest))
                left_of.append(k)
        if left_groups:
            # a leftover group with one detection per camera is taken as is, otherwise scored again
            for (f, rest), k, b in zip(left_groups, left_of, score_round(left_groups)):
                persons.append((int(frames[f]), (int(ukey[k] % M), 1), b))
    return [p for p in persons if (p[2] >= 0).sum() >= 2]                   # step2:697-698


def associate_batch(cgroup, kp_raw, dim, cid=None, bbox_id=None, thr_kp=THR_KP, alpha_id=0.2, alpha_svt=0.5,
                    lambda_svt=50.0, timing=None):
    """Cross-view association + reconstruction of F keyframes at once — what
    MultiEstimator.predict_data (step2_crossviewmatching.py:502-713) does per keyframe.

      kp_raw (F, M, J, 3)  raw pixels + score of every detection, grouped by camera (padding beyond dim[:, C]);
                           numpy, or a float64 CUDA tensor (then nothing of size F*M*J crosses the bus)
      dim    (F, C+1)      cumulative detection counts per camera (dimGroup)
      cid    (F, M) int    identity label per detection or -1 (None: no identity term)
      bbox_id (F, M) int   tracklet id reported back per person and camera (None: the detection index)

    Everything of size F runs on the device, all keyframes per launch: undistortion of the detections, ray
    affinity, association weights, SVT matching, cluster labels, member tables of the persons, least-squares
    triangulation.  The host sees the (F, M) labels, the (P, C) member tables and the result.  Keyframes in
    which a cluster holds two detections of one camera (rare) are resolved on the host by scoring ALL
    their candidate combinations in one batch (``_resolve_duplicate_frames``).

    Returns {'label' (F,M) person column per detection or -1, 'frame' (P,), 'members' (P,C) detection index
    per camera or -1, 'p3d' (P,J,3), 'bcomb' (P,C)}; persons are ordered by frame, then like the reference
    (ascending cluster column; a cluster's leftover combination follows its best one)."""
    import time as _time
    t_last = [_time.perf_counter()]

    def lap(name):
        if timing is not None:
            torch.cuda.synchronize()
            now = _time.perf_counter()
            timing[name] = timing.get(name, 0.0) + now - t_last[0]
            t_last[0] = now

    device = cgroup._dev()
    lib = _lib.require_gpu()
    rig = cgroup._rig(device)
    dev = "cuda:%d" % device
    C = len(cgroup.cameras)
    if _is_torch(kp_raw):
        d_raw = kp_raw.to(device=dev, dtype=torch.float64).contiguous()
    else:
        d_raw = torch.from_numpy(np.ascontiguousarray(kp_raw, dtype=np.float64)).to(dev)
    F, M, J, _ = d_raw.shape
    dim = np.ascontiguousarray(dim.cpu().numpy() if _is_torch(dim) else dim, dtype=np.int32)
    assert dim.shape == (F, C + 1)
    d_dim = torch.from_numpy(dim).to(dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = _stream(device)
    lap("upload")
    d_und = torch.empty_like(d_raw)
    _lib.check(lib.m3d_undistort_detections(rig.handle, vp(d_raw), vp(d_dim), F, M, J, vp(d_und), st),
               "m3d_undistort_detections")
    aff = geometry_affinity_batch(cgroup, torch.nan_to_num(d_und), d_dim, thr_kp)
    ids = np.full((F, M), -1, dtype=np.int32) if cid is None else \
        np.ascontiguousarray(cid.cpu().numpy() if _is_torch(cid) else cid, dtype=np.int32)
    d_cid = torch.from_numpy(ids).to(dev)
    W = torch.empty_like(aff)
    _lib.check(lib.m3d_association_weights(_ptr(aff), vp(d_cid), vp(d_dim), F, M, C, float(alpha_id), _ptr(W),
                                           int(device), st), "m3d_association_weights")
    del aff
    lap("undistort+affinity+weights")
    match = match_svt_batch(W, d_dim, C, alpha=alpha_svt, _lambda=lambda_svt, device=device)
    del W
    lap("svt")
    d_label = torch.empty((F, M), dtype=torch.int32, device=dev)
    _lib.check(lib.m3d_match_clusters(vp(match), vp(d_dim), F, M, C, vp(d_label), int(device), st),
               "m3d_match_clusters")
    del match
    # persons as member tables, (frame, column) order: count, prefix sums, write
    d_count = torch.empty((F,), dtype=torch.int32, device=dev)
    d_dup = torch.empty((F,), dtype=torch.uint8, device=dev)
    _lib.check(lib.m3d_cluster_members(vp(d_label), vp(d_dim), F, M, C, None, vp(d_count), vp(d_dup), None, None,
                                       None, int(device), st), "m3d_cluster_members")
    csum = torch.cumsum(d_count, 0, dtype=torch.int64)
    P0 = int(csum[-1].item()) if F else 0
    d_off = (csum - d_count).contiguous()
    d_frame = torch.empty((P0,), dtype=torch.int32, device=dev)
    d_col = torch.empty((P0,), dtype=torch.int32, device=dev)
    d_mem = torch.empty((P0, C), dtype=torch.int32, device=dev)
    if P0:
        _lib.check(lib.m3d_cluster_members(vp(d_label), vp(d_dim), F, M, C, vp(d_off), vp(d_count), vp(d_dup),
                                           vp(d_frame), vp(d_col), vp(d_mem), int(device), st), "m3d_cluster_members")
    label = d_label.cpu().numpy()
    dup_frames = torch.nonzero(d_dup).reshape(-1)
    lap("clusters")
    if dup_frames.numel():
        # the rare frames with duplicate detections: host bookkeeping on their rows only, merged by (frame, column)
        fr_h = dup_frames.cpu().numpy()
        lab_s = label[fr_h]
        extra = _resolve_duplicate_frames(cgroup, fr_h, lab_s, _cam_of(dim[fr_h], M), d_raw[dup_frames].cpu().numpy(),
                                          d_und[dup_frames][..., :2].cpu().numpy(), thr_kp)
        frame = np.concatenate([d_frame.cpu().numpy().astype(np.int64), np.array([p[0] for p in extra], dtype=np.int64)])
        col = np.concatenate([d_col.cpu().numpy().astype(np.int64) * 2,
                              np.array([2 * p[1][0] + p[1][1] for p in extra], dtype=np.int64)])
        mem = np.concatenate([d_mem.cpu().numpy().astype(np.int64),
                              np.array([p[2] for p in extra], dtype=np.int64).reshape(-1, C)])
        order = np.lexsort((col, frame))
        frame, mem = frame[order], mem[order]
        d_frame = torch.from_numpy(frame.astype(np.int32)).to(dev)
        d_mem = torch.from_numpy(np.ascontiguousarray(mem, dtype=np.int32)).to(dev)
        lap("duplicates")
    else:
        frame = d_frame.cpu().numpy().astype(np.int64)
        mem = d_mem.cpu().numpy().astype(np.int64)
    P = frame.shape[0]
    d_p3d = torch.empty((P, J, 3), dtype=torch.float64, device=dev)
    if P:
        _lib.check(lib.m3d_triangulate_ls_members(rig.handle, vp(d_und), vp(d_frame), vp(d_mem), P, M, J,
                                                  float(thr_kp), vp(d_p3d), st), "m3d_triangulate_ls_members")
    p3d = d_p3d.cpu().numpy()
    has = mem >= 0
    if bbox_id is None:
        src = mem
    else:
        bb = np.asarray(bbox_id.cpu().numpy() if _is_torch(bbox_id) else bbox_id)
        src = np.where(has, bb[frame[:, None], np.where(has, mem, 0)], -1)
    lap("triangulate+download")
    return {"label": label, "frame": frame, "members": mem, "p3d": p3d, "bcomb": np.where(has, src, -1)}


class MultiEstimator:
    """Matching and 3D reconstruction across cameras for a single keyframe: the call shape of
    step2_crossviewmatching.py:494-713 on ``associate_batch`` with F = 1 (the reference's per-keyframe
    loop, step2:899-928, is better served by calling ``associate_batch`` once for all keyframes).
    Drawing (``show=True``) is not reproduced."""

    def __init__(self, cfg=None, debug=False):
        self.cfg = cfg
        self.debug = debug

    def predict_data(self, info_dict, show=False, plt_id=0, camparam=None, bcomb_prev=None):
        _need_camparam(camparam)
        if show:
            raise NotImplementedError("predict_data(show=True) draws with matplotlib (step2:648-693)")
        cg = group_from_camparam(camparam)
        dets = [d for cam_id in range(len(info_dict)) for d in info_dict[cam_id][0]]
        if not dets:
            return [], [], []
        J = MODEL_CFG["joint_num"]
        dim = np.cumsum([0] + [len(info_dict[c][0]) for c in range(len(info_dict))]).astype(np.int32)[None]
        kp_raw = np.array([d["pose2d_raw"] for d in dets], dtype=np.float64).reshape(1, len(dets), J, 3)
        cid = np.array([d["cid"] for d in dets], dtype=np.int32)[None]
        bbox = np.array([d["bbox_id"][1] for d in dets], dtype=np.int64)[None]
        res = associate_batch(cg, kp_raw, dim, cid, bbox, thr_kp=THR_KP, alpha_id=ALPHA_ID,
                              alpha_svt=MODEL_CFG["alpha_SVT"], lambda_svt=MODEL_CFG["lambda_SVT"])
        matched = [row[row >= 0] for row in res["members"]]
        return matched, list(res["p3d"]), list(res["bcomb"])


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


# ------------------------------------------------------------------------------------------
# step-3 call sites of the same arithmetic (tracklets -> keypoint arrays -> batched kernels)
# ------------------------------------------------------------------------------------------

def tracklet_keypoints(trk, T, frames, n_kp):
    """The (F, C, J, 3) keypoint array step 3 assembles frame by frame for one tracklet
    (step3_crossframematching.py:289-297, 1116-1125): ``T[i_cam][i_frame]`` is the list of 2D tracks of a
    camera in a frame (entry[0] = bbox id, entry[5] = (J,3) keypoints), ``trk[i_frame][i_cam]`` the bbox id
    the tracklet uses there (-1 = none).  Cameras without a matching track are NaN."""
    frames = np.asarray(frames, dtype=int).ravel()
    n_cam = len(T)
    out = np.full((frames.size, n_cam, n_kp, 3), np.nan)
    for k, i_frame in enumerate(frames):
        for i_cam in range(n_cam):
            want = trk[i_frame][i_cam]
            for tt in T[i_cam][i_frame]:
                if tt[0] == want:
                    out[k, i_cam] = np.asarray(tt[5], dtype=np.float64)
    return out


def calc_p3d(T, trk, i_frame, camparam, n_kp=MODEL_CFG["joint_num"]):
    """get_graph's calc_p3d (step3:1115-1130): mean 3D position of a tracklet's keypoints in one frame."""
    import warnings
    p3d = calc_3dpose_batch(tracklet_keypoints(trk, T, [i_frame], n_kp), camparam)[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.nanmean(p3d, axis=0)


def calc_3dtrace_tracklet(trk, T, frames, camparam, config_path=None, n_kp=MODEL_CFG["joint_num"]):
    """step3's calc_3dtrace with its own signature (:274-302): per-frame median 3D position of a tracklet
    over ``frames`` (one launch for all of them); rows of the other frames are NaN, like the reference's."""
    frames = np.asarray(frames, dtype=int).ravel()
    n_frame = len(T[0])
    trace = np.full((n_frame, 3), np.nan)
    if frames.size:
        trace[frames] = calc_3dtrace(tracklet_keypoints(trk, T, frames, n_kp), camparam)
        seen = np.array([(np.asarray(trk[f]) >= 0).sum() for f in frames])
        trace[frames[seen < 2]] = np.nan                                   # :284-285
    return trace


def trace_distance(p1, p2):
    """step3's calc_dist_pose on two traces: RMS distance over the frames both have (step3:304-310)."""
    d = np.linalg.norm(np.asarray(p1) - np.asarray(p2), axis=-1)
    ok = ~np.isnan(d)
    return float(np.sqrt(np.mean(d[ok] ** 2))) if ok.any() else np.nan
