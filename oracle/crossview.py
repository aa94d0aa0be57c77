"""Oracle (test infrastructure, see oracle/__init__.py): numpy restatement of the
cross-view association arithmetic.

Reference:
  deproject               src/pipeline/step2_crossviewmatching.py:327-355
  calc_dist_btw_lines     step2_crossviewmatching.py:359-369
  geometry_affinity2      step2_crossviewmatching.py:373-432
  matchSVT                step2_crossviewmatching.py:130-216
  mct.triangulatePoints   src/utils/multicam_toolbox.py:433-486
"""
import numpy as np


def _camera_of(dimGroup, i):
    return int(np.searchsorted(dimGroup, i, side="right") - 1)


def ray_distance_matrix(cams, points_set, dimGroup, thr_kp=0.1):
    """dist_mat of geometry_affinity2 before the sigmoid (step2:388-424)."""
    points_set = np.asarray(points_set, dtype=np.float64)
    M, J, _ = points_set.shape
    D = np.full((M, M), 300.0)
    np.fill_diagonal(D, 0.0)
    Rs = [c.extrinsics()[:3, :3] for c in cams]
    ts = [c.tvec for c in cams]
    cam_of = [_camera_of(dimGroup, i) for i in range(M)]
    near, far = [], []
    for i in range(M):
        Rinv = np.linalg.inv(Rs[cam_of[i]])
        t = ts[cam_of[i]]
        xy1 = np.hstack([points_set[i, :, :2], np.ones((J, 1))])
        near.append((Rinv @ (xy1 * 0.0 - t).T).T)
        far.append((Rinv @ (xy1 * 1000.0 - t).T).T)
    S = points_set[:, :, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(M):
            for j in range(i + 1, M):
                if cam_of[i] == cam_of[j]:
                    continue
                ok = (S[i] > thr_kp) & (S[j] > thr_kp)
                if ok.sum() < 3:
                    continue
                d1 = far[i][ok] - near[i][ok]
                d1 = d1 / np.linalg.norm(d1, axis=1, keepdims=True)
                d2 = far[j][ok] - near[j][ok]
                d2 = d2 / np.linalg.norm(d2, axis=1, keepdims=True)
                c = np.cross(d1, d2)
                d = np.abs(np.sum((near[j][ok] - near[i][ok]) * c, axis=1)) / np.linalg.norm(c, axis=1)
                D[i, j] = D[j, i] = np.mean(d)
    return D


def geometry_affinity(cams, points_set, dimGroup, thr_kp=0.1):
    """geometry_affinity2 (step2:373-432): points_set (M,J,3) undistorted x, y, score."""
    D = ray_distance_matrix(cams, points_set, dimGroup, thr_kp)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        valid = D < 300.0
        mu = D[valid].mean()
        sd = D[valid].std()
        aff = -(D - mu) / sd
        aff = 1 / (1 + np.exp(-5 * aff))
        aff[D > 150.0] = 0
    return aff


def match_svt(S, dimGroup, alpha=0.1, lam=50.0, mu=64.0, tol=5e-4, max_iter=500, return_info=False):
    """matchSVT with pselect = 1, dual_stochastic_SVT = False (step2:130-216)."""
    S = np.array(S, dtype=np.float64, copy=True)
    N = S.shape[0]
    S[np.arange(N), np.arange(N)] = 0
    S = (S + S.T) / 2
    X = S.copy()
    Y = np.zeros_like(S)
    W = alpha - S
    it = 0
    for it in range(max_iter):
        X0 = X.copy()
        U, s, Vh = np.linalg.svd(Y / mu + X, full_matrices=False)
        Q = U @ np.diag(np.maximum(s - lam / mu, 0)) @ Vh
        X = Q - (W + Y) / mu
        for i in range(len(dimGroup) - 1):
            a, b = int(dimGroup[i]), int(dimGroup[i + 1])
            X[a:b, a:b] = 0
        X[np.arange(N), np.arange(N)] = 1
        X = np.clip(X, 0, 1)
        X = (X + X.T) / 2
        Y = Y + mu * (X - Q)
        pRes = np.linalg.norm(X - Q) / N
        dRes = mu * np.linalg.norm(X - X0) / N
        if pRes < tol and dRes < tol:
            break
        if pRes > 10 * dRes:
            mu *= 2
        elif dRes > 10 * pRes:
            mu /= 2
    X = (X + X.T) / 2
    match = (X > 0.5).astype(np.uint8)
    if return_info:
        return match, it, X
    return match


def triangulate_ls(cams, xy_undist, frame_use):
    """mct.triangulatePoints (multicam_toolbox.py:458-486): xy_undist (C,n,2),
    frame_use (n,C) bool -> (n,3); X = -pinv(A[:, :3]) @ A[:, 3]."""
    xy = np.asarray(xy_undist, dtype=np.float64)
    use = np.asarray(frame_use, dtype=bool)
    n, C = use.shape
    P = [c.extrinsics()[:3, :] for c in cams]
    out = np.zeros((n, 3))
    for i in range(n):
        if use[i].sum() < 2:
            out[i] = np.nan
            continue
        rows = []
        for c in range(C):
            if use[i, c]:
                rows.append(xy[c, i, 0] * P[c][2] - P[c][0])
                rows.append(xy[c, i, 1] * P[c][2] - P[c][1])
        A = np.vstack(rows)
        out[i] = -(np.linalg.pinv(A[:, :3]) @ A[:, 3])
    return out
