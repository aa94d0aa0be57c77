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


def associate_frame(cams, kp_raw, dimGroup, cid, bbox_id, thr_kp=0.1, alpha_id=0.2, alpha_svt=0.5, lam=50.0):
    """MultiEstimator.predict_data for one keyframe in the reference's own loop structure
    (step2_crossviewmatching.py:502-713) with the camera model taken from ``cams`` (the reference hard-wires
    cv2.omnidir there, which this container cannot execute): affinity :554, identity term :557-575,
    matchSVT :589-595, clusters :598-607, get_best_comb :610-646, leftovers :649-657, 3D poses :696-713.
    kp_raw (M,J,3) raw pixels + score.  Returns (matched_list, P3d_list, bcomb_list)."""
    import itertools
    kp_raw = np.asarray(kp_raw, dtype=np.float64)
    M, J, _ = kp_raw.shape
    C = len(cams)
    dimGroup = np.asarray(dimGroup)
    sub2cam = np.zeros(M, dtype=int)
    for idx in range(C):
        sub2cam[dimGroup[idx]:dimGroup[idx + 1]] = idx
    und = np.stack([cams[sub2cam[i]].undistort(kp_raw[i, :, :2]) for i in range(M)]) if M else np.zeros((0, J, 2))
    kp_mat = np.concatenate([und, kp_raw[:, :, 2:3]], axis=2)
    geo = geometry_affinity(cams, kp_mat, dimGroup, thr_kp)
    cid = np.asarray(cid)
    cid_mat = np.zeros((M, M))
    for i in range(M):
        for j in range(M):
            if sub2cam[i] != sub2cam[j] and cid[i] >= 0 and cid[i] == cid[j]:
                cid_mat[i, j] = 1
    W = alpha_id * cid_mat + (1 - alpha_id) * geo
    W *= (geo > 0)
    W = np.nan_to_num(W)
    match_mat = match_svt(W, dimGroup, alpha=alpha_svt, lam=lam)
    bin_match = match_mat[:, np.nonzero(np.sum(match_mat, axis=0) > 1.9)[0]] > 0.9
    bin_match = bin_match.reshape(M, -1)
    matched_list = [[] for _ in range(bin_match.shape[1])]
    for sub_imgid, row in enumerate(bin_match):
        if row.sum() != 0:
            matched_list[np.argmax(row)].append(sub_imgid)
    # (an empty cluster — all rows of a person column claimed by earlier columns — makes the reference's
    # float-typed empty index array raise; it carries no person either way)
    matched_list = [np.array(x, dtype=int) for x in matched_list if len(x)]

    def pose3d(kp2d):                                            # calc_3dpose, step2:436-461
        u = np.stack([cams[c].undistort(kp2d[c, :, :2]) for c in range(C)])
        use = ~(np.isnan(kp2d[:, :, 0]) | (kp2d[:, :, 2] < thr_kp))
        return triangulate_ls(cams, np.nan_to_num(u), use.T)

    def get_best_comb(person):
        cam_list = sub2cam[person]
        groups = [np.array(person)[np.argwhere(cam_list == c).ravel()].tolist() or [None] for c in range(C)]
        combs = list(itertools.product(*groups))
        if len(combs) == 1:
            return np.array(person)
        errs = []
        for comb in combs:
            kp2d = np.zeros((C, J, 3))
            for c, i in enumerate(comb):
                if i is not None:
                    kp2d[c] = kp_raw[i]
            p3d = pose3d(kp2d)
            d = []
            for c, i in enumerate(comb):
                if i is not None:
                    keep = kp_raw[i][:, 2] > thr_kp
                    d.append(kp_raw[i][keep, :2] - cams[c].project(p3d)[keep])
            errs.append(np.sqrt(np.mean(np.concatenate(d, axis=0) ** 2)))
        best = combs[int(np.argmin(errs))]
        return np.array([i for i in best if i is not None])

    refined = []
    for person in matched_list:
        b = get_best_comb(person)
        refined.append(b)
        rest = set(person.tolist()) - set(b.tolist())
        if len(rest) > 1:
            refined.append(get_best_comb(np.array(list(rest))))
    out_m, out_p, out_b = [], [], []
    for person in refined:
        if person.shape[0] < 2:
            continue
        kp2d = np.zeros((C, J, 3))
        for i in person:
            kp2d[sub2cam[i]] = kp_raw[i]
        out_p.append(pose3d(kp2d))
        bc = -np.ones(C, dtype=int)
        for i in person:
            bc[sub2cam[i]] = bbox_id[i]
        out_m.append(person)
        out_b.append(bc)
    return out_m, out_p, out_b
