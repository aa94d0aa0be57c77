"""Host side of ``CameraGroup.optim_points`` / ``optim_points_jointlenfix`` (reference
aniposelib/cameras.py:1116-1270): the cheap O(F*J) pre-processing in numpy — gap interpolation, the
median-filtered track that sets the smoothness scale, the start vector x0 — and the call into the GPU
solver ``m3d_optim_points`` (csrc/m3d_optim.cu), which owns the residuals, the Jacobian and the
Levenberg-Marquardt / CG iteration."""
import ctypes

import numpy as np

from . import _lib
from ._device import _ptr, _stream, _to_dev, torch

LOSSES = {"linear": 0, "soft_l1": 1, "huber": 2}


def interpolate_columns(p3ds):
    """Linear interpolation over the NaN gaps of every (joint, coordinate) track along the frame
    axis, ends held (np.interp); a track without a single finite value becomes 0
    (cameras.py:138-145 applied along axis 0 at :1148)."""
    a = np.array(p3ds, dtype=np.float64, copy=True)
    flat = a.reshape(a.shape[0], -1)
    t = np.arange(flat.shape[0])
    for k in range(flat.shape[1]):
        col = flat[:, k]
        bad = np.isnan(col)
        if bad.all():
            col[:] = 0.0
        elif bad.any():
            col[bad] = np.interp(t[bad], t[~bad], col[~bad])
    return flat.reshape(a.shape)


def median_filter_columns(vals, size=7):
    """Running median of width ``size`` along axis 0 on a reflect-padded copy (cameras.py:129-133,
    called with size = 7 at :1150).  scipy.signal.medfilt zero-pads, but only inside the reflected
    margin of size + 5 samples, which is cut off again: the result is the plain reflect-padded median."""
    a = np.asarray(vals, dtype=np.float64)
    flat = a.reshape(a.shape[0], -1)
    h = size // 2
    pad = np.pad(flat, ((h, h), (0, 0)), mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(pad, size, axis=0)      # (F, cols, size)
    return np.median(win, axis=2).reshape(a.shape)


def smoothness_scale(p3ds_intp, scale_smooth):
    """scale_smooth / mean |first difference of the median-filtered track| (cameras.py:1150-1153)."""
    med = median_filter_columns(p3ds_intp, 7)
    return float(scale_smooth) * (1.0 / np.mean(np.abs(np.diff(med, axis=0))))


def initial_lengths(p3ds, constraints, constraints_weak):
    """Start values of the limb lengths: per-constraint median length, zeros and outliers
    (> median + 5 MAD over all constraints) replaced by the overall median (cameras.py:1664-1690)."""
    def med_len(cons):
        out = np.empty(len(cons))
        for i, (a, b) in enumerate(cons):
            out[i] = np.median(np.linalg.norm(p3ds[:, a] - p3ds[:, b], axis=1))
        return out
    strong, weak = med_len(constraints), med_len(constraints_weak)
    both = np.hstack([strong, weak])
    med = np.median(both)
    if med == 0:
        med = 1e-3
    mad = np.median(np.abs(both - med))
    for arr in (strong, weak):
        arr[arr == 0] = med
        arr[arr > med + mad * 5] = med
    return strong, weak


def _cons(c):
    a = np.asarray(c, dtype=np.int32).reshape(-1, 2) if len(c) else np.zeros((0, 2), dtype=np.int32)
    return np.ascontiguousarray(a)


def _call(cgroup, points, params, scores, constraints, constraints_weak, scale_smooth_full, scale_length,
          scale_length_weak, rp, reproj_loss, n_deriv, fix_lengths, ftol, max_iter, mode, out=None):
    if reproj_loss not in LOSSES:
        raise ValueError("reproj_loss must be one of %s" % sorted(LOSSES))
    n_cams, n_frames, n_joints, _ = points.shape
    device, _ = cgroup._device_of(points)
    rig = cgroup._rig(device)
    p2d = _to_dev(points, device)
    sc = _to_dev(scores, device) if scores is not None else None
    cs, cw = _cons(constraints), _cons(constraints_weak)
    x = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64)).to(p2d.device)
    info = np.zeros(8)
    o = out
    _lib.check(rig._lib.m3d_optim_points(
        rig.handle, _ptr(p2d), _ptr(sc) if sc is not None else None, int(n_frames), int(n_joints),
        cs.ctypes.data_as(ctypes.c_void_p), int(cs.shape[0]), cw.ctypes.data_as(ctypes.c_void_p), int(cw.shape[0]),
        float(scale_smooth_full), float(scale_length), float(scale_length_weak), float(rp), LOSSES[reproj_loss],
        int(n_deriv), int(bool(fix_lengths)), float(ftol), int(max_iter), int(mode), _ptr(x),
        _ptr(o) if o is not None else None, info.ctypes.data_as(ctypes.c_void_p), _stream(device)),
        "m3d_optim_points")
    return x, info


def residual_sizes(n_cams, n_frames, n_joints, n_cons, n_weak, n_deriv):
    n1 = n_cams * n_frames * n_joints * 2
    n2 = max(0, n_frames - n_deriv) * n_joints * 3
    return n1, n2, n_cons * n_frames, n_weak * n_frames


def error_fun(cgroup, params, p2ds, constraints=(), constraints_weak=(), scores=None, scale_smooth=10000,
              scale_length=1, scale_length_weak=0.2, reproj_error_threshold=100, reproj_loss="soft_l1",
              n_deriv_smooth=1, joint_len=None, dense=False):
    """``CameraGroup._error_fun_triangulation`` (cameras.py:1560-1620; ``joint_len`` given = the
    ``_jointlenfix`` form :1356-1416) evaluated by the GPU kernels: the same residual vector, in the same
    order (reprojection residuals of the non-NaN 2D coordinates, smoothness, strong and weak lengths)."""
    n_cams, n_frames, n_joints, _ = p2ds.shape
    K, Kw = len(constraints), len(constraints_weak)
    sizes = residual_sizes(n_cams, n_frames, n_joints, K, Kw, n_deriv_smooth)
    full = np.asarray(params, dtype=np.float64)
    if joint_len is not None:
        full = np.hstack([full[:n_frames * n_joints * 3], np.asarray(joint_len, dtype=np.float64)])
    device, _ = cgroup._device_of(p2ds)
    out = torch.empty((max(sum(sizes), full.size),), dtype=torch.float64, device="cuda:%d" % device)
    _call(cgroup, p2ds, full, scores, constraints, constraints_weak, scale_smooth, scale_length, scale_length_weak,
          reproj_error_threshold, reproj_loss, n_deriv_smooth, joint_len is not None, 0.0, 0, 1, out)
    r = out[:sum(sizes)].cpu().numpy()
    if dense:
        return r
    keep = ~np.isnan(np.asarray(p2ds, dtype=np.float64).reshape(-1))
    return np.hstack([r[:sizes[0]][keep], r[sizes[0]:]])


def jvp(cgroup, params, v, p2ds, constraints=(), constraints_weak=(), scores=None, scale_smooth=10000,
        scale_length=1, scale_length_weak=0.2, reproj_error_threshold=100, reproj_loss="soft_l1",
        n_deriv_smooth=1, fix_lengths=False):
    """J(params) @ v with the solver's exact Jacobian blocks, in the layout of ``error_fun``."""
    n_cams, n_frames, n_joints, _ = p2ds.shape
    K, Kw = len(constraints), len(constraints_weak)
    sizes = residual_sizes(n_cams, n_frames, n_joints, K, Kw, n_deriv_smooth)
    device, _ = cgroup._device_of(p2ds)
    out = torch.zeros((max(sum(sizes), len(params)),), dtype=torch.float64, device="cuda:%d" % device)
    nv = n_frames * n_joints * 3 + (0 if fix_lengths else K + Kw)
    out[:nv] = torch.from_numpy(np.asarray(v, dtype=np.float64)[:nv]).to(out.device)
    _call(cgroup, p2ds, params, scores, constraints, constraints_weak, scale_smooth, scale_length, scale_length_weak,
          reproj_error_threshold, reproj_loss, n_deriv_smooth, fix_lengths, 0.0, 0, 2, out)
    r = out[:sum(sizes)].cpu().numpy()
    keep = ~np.isnan(np.asarray(p2ds, dtype=np.float64).reshape(-1))
    return np.hstack([r[:sizes[0]][keep], r[sizes[0]:]])


def optim_points(cgroup, points, p3ds, constraints=(), constraints_weak=(), scale_smooth=4, scale_length=2,
                 scale_length_weak=0.5, reproj_error_threshold=15, reproj_loss="soft_l1", n_deriv_smooth=1,
                 scores=None, verbose=False, joint_len=None, ftol=1e-4, max_iter=60, return_info=False):
    """``CameraGroup.optim_points`` (``joint_len`` None) / ``optim_points_jointlenfix``
    (cameras.py:1116-1270).  points (C,F,J,2), p3ds (F,J,3) -> (p3ds_new (F,J,3), joint_len (K+Kw,)).

    Same objective, same start vector; the GPU solver uses exact derivatives and runs to ``ftol``
    (default 1e-4: tighter than the reference's 1e-3, so the final cost is not above the
    reference's) instead of reproducing scipy's trust-region trajectory step by step."""
    assert points.shape[0] == len(cgroup.cameras), \
        "Invalid points shape, first dim should be equal to" \
        " number of cameras ({}), but shape is {}".format(len(cgroup.cameras), points.shape)
    p3ds = np.asarray(p3ds.cpu().numpy() if hasattr(p3ds, "cpu") else p3ds, dtype=np.float64)
    n_cams, n_frames, n_joints, _ = points.shape
    constraints = [tuple(c) for c in np.asarray(constraints).reshape(-1, 2)] if len(constraints) else []
    constraints_weak = [tuple(c) for c in np.asarray(constraints_weak).reshape(-1, 2)] if len(constraints_weak) else []
    intp = interpolate_columns(p3ds)
    s_full = smoothness_scale(intp, scale_smooth)
    strong, weak = initial_lengths(intp, constraints, constraints_weak)
    fix = joint_len is not None
    lens = np.asarray(joint_len, dtype=np.float64) if fix else np.hstack([strong, weak])
    x0 = np.hstack([intp.ravel(), lens])
    if not fix:
        x0[~np.isfinite(x0)] = 0
    else:
        x0[:intp.size][~np.isfinite(x0[:intp.size])] = 0
    x, info = _call(cgroup, points, x0, scores, constraints, constraints_weak, s_full, scale_length,
                    scale_length_weak, reproj_error_threshold, reproj_loss, n_deriv_smooth, fix, ftol, max_iter, 0)
    xs = x.cpu().numpy()
    new = xs[:p3ds.size].reshape(p3ds.shape)
    jl = lens if fix else xs[p3ds.size:]
    if verbose:
        print("optim_points: cost %.6g -> %.6g in %d LM steps (%d CG iterations, %d residual evaluations), status %d"
              % (info[1], info[0], info[2], info[3], info[4], info[5]))
    if return_info:
        return new, jl, {"cost": info[0], "cost0": info[1], "lm_steps": int(info[2]), "cg_iterations": int(info[3]),
                         "evaluations": int(info[4]), "status": int(info[5]), "scale_smooth_full": s_full, "x0": x0}
    return new, jl
