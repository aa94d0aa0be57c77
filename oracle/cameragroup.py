"""Oracle (test infrastructure, see oracle/__init__.py): numpy restatement of the
reference's CameraGroup triangulation core.

Reference (all /root/reference/src/third_party/aniposelib/cameras.py):
  triangulate_simple               :20-32     DLT rows + SVD, vh[-1][:3]/vh[-1][3]
  CameraGroup.project              :580-591
  CameraGroup.triangulate          :593-637
  CameraGroup.triangulate_possible :639-724   exhaustive subset search with early exit
  CameraGroup.triangulate_ransac   :726-743   = triangulate_possible with P = 1
  CameraGroup.reprojection_error   :746-783

Two forms are provided:
  * vectorised (``triangulate``, ``reprojection_error``, ``triangulate_ransac``):
    points grouped by validity pattern, batched LAPACK SVD — the checker used by
    the parity tests;
  * loop-faithful (``*_loops``): one OpenCV / LAPACK call per point / subset /
    camera exactly like the reference's Python loops — used as the timed CPU
    baseline ("port") in bench.py because it has the reference's cost structure.
Both are pinned against the reference's own output in tests/golden/.
"""
import itertools

import numpy as np

from . import camera_math as cm


# ----------------------------------------------------------------------------
# vectorised checker
# ----------------------------------------------------------------------------

def undistort_points(cams, points):
    """Per-camera undistortion of a (C, N, 2) array (cameras.py:608-614)."""
    points = np.asarray(points, dtype=np.float64)
    out = np.empty(points.shape)
    for c, cam in enumerate(cams):
        out[c] = cam.undistort(points[c])
    return out


def project(cams, p3d):
    """CameraGroup.project: (N,3) -> (C,N,2)  (cameras.py:580-591)."""
    p3d = np.asarray(p3d, dtype=np.float64).reshape(-1, 3)
    out = np.empty((len(cams), p3d.shape[0], 2))
    for c, cam in enumerate(cams):
        out[c] = cam.project(p3d)
    return out


def _dlt_rows(U, Ms):
    """Rows x*M[2]-M[0], y*M[2]-M[1] for U (k, n, 2), Ms (k, 4, 4) -> (n, 2k, 4)
    (cameras.py:27-28)."""
    k, n, _ = U.shape
    A = np.empty((n, 2 * k, 4))
    for j in range(k):
        A[:, 2 * j, :] = U[j, :, 0:1] * Ms[j][2][None, :] - Ms[j][0][None, :]
        A[:, 2 * j + 1, :] = U[j, :, 1:2] * Ms[j][2][None, :] - Ms[j][1][None, :]
    return A


def _dlt_solve(A):
    """Smallest right singular vector of each (2k x 4) system, dehomogenised
    (cameras.py:29-31)."""
    if A.shape[0] == 0:
        return np.empty((0, 3))
    with np.errstate(invalid="ignore", divide="ignore"):
        _, _, vh = np.linalg.svd(A, full_matrices=True)
        p = vh[:, -1, :]
        return p[:, :3] / p[:, 3:4]


def triangulate_undistorted(cams, U):
    """DLT on already-undistorted points U (C, N, 2); a camera is used for a point
    iff its x is not NaN; fewer than two -> NaN  (cameras.py:616-632)."""
    U = np.asarray(U, dtype=np.float64)
    C, N, _ = U.shape
    Ms = np.array([cam.extrinsics() for cam in cams]) if C else np.empty((0, 4, 4))
    out = np.full((N, 3), np.nan)
    good = ~np.isnan(U[:, :, 0])                      # (C, N)
    weights = (1 << np.arange(C, dtype=np.int64))[:, None]
    key = (good * weights).sum(axis=0)
    for m in np.unique(key):
        idx = np.nonzero(key == m)[0]
        cs = [c for c in range(C) if (int(m) >> c) & 1]
        if len(cs) < 2:
            continue
        A = _dlt_rows(U[cs][:, idx], Ms[cs])
        out[idx] = _dlt_solve(A)
    return out


def triangulate(cams, points, undistort=True):
    """CameraGroup.triangulate (cameras.py:593-637).  points (C,N,2) or (C,2)."""
    points = np.asarray(points, dtype=np.float64)
    assert points.shape[0] == len(cams), \
        "Invalid points shape, first dim should be equal to" \
        " number of cameras ({}), but shape is {}".format(len(cams), points.shape)
    one_point = False
    if points.ndim == 2:
        points = points.reshape(-1, 1, 2)
        one_point = True
    U = undistort_points(cams, points) if undistort else points
    out = triangulate_undistorted(cams, U)
    return out[0] if one_point else out


def reprojection_error(cams, p3ds, p2ds, mean=False):
    """CameraGroup.reprojection_error (cameras.py:746-783)."""
    p3ds = np.asarray(p3ds, dtype=np.float64)
    p2ds = np.asarray(p2ds, dtype=np.float64)
    one_point = False
    if p3ds.ndim == 1 and p2ds.ndim == 2:
        p3ds = p3ds.reshape(1, 3)
        p2ds = p2ds.reshape(-1, 1, 2)
        one_point = True
    n_cams, n_points, _ = p2ds.shape
    assert p3ds.shape == (n_points, 3), \
        "shapes of 2D and 3D points are not consistent: " \
        "2D={}, 3D={}".format(p2ds.shape, p3ds.shape)
    errors = np.empty((n_cams, n_points, 2))
    for c, cam in enumerate(cams):
        errors[c] = p2ds[c] - cam.project(p3ds)
    if mean:
        with np.errstate(invalid="ignore", divide="ignore"):
            ex, ey = errors[:, :, 0], errors[:, :, 1]
            nrm = np.sqrt(ex * ex + ey * ey)
            good = ~np.isnan(nrm)
            nrm[~good] = 0
            denom = np.sum(good, axis=0).astype("float64")
            denom[denom < 1.5] = np.nan
            errors = np.sum(nrm, axis=0) / denom
    if one_point:
        if mean:
            errors = float(errors[0])
        else:
            errors = errors.reshape(-1, 2)
    return errors


def subset_cameras_of(valid_cams, s):
    """Cameras of enumeration step ``s`` for the ascending valid-camera list
    (itertools.product order of cameras.py:689): camera valid_cams[j] is included
    iff bit (k-1-j) of s is 0."""
    k = len(valid_cams)
    return [valid_cams[j] for j in range(k) if not (s >> (k - 1 - j)) & 1]


def triangulate_ransac(cams, points, undistort=True, min_cams=2, threshold=0.5,
                       init_best=200.0, return_stats=False):
    """CameraGroup.triangulate_ransac / triangulate_possible with P = 1
    (cameras.py:639-743).

    Returns (out (N,3), picked (C,N,1) bool, points_2d (C,N,2), errors (N,)) and,
    with return_stats, also (subset_index (N,) int64 [-1 = nothing selected],
    n_evaluated (N,) int64 = subsets the reference would have triangulated).
    """
    points = np.asarray(points, dtype=np.float64)
    assert points.shape[0] == len(cams), \
        "Invalid points shape, first dim should be equal to" \
        " number of cameras ({}), but shape is {}".format(len(cams), points.shape)
    C, N, _ = points.shape
    out = np.full((N, 3), np.nan)
    picked = np.zeros((C, N, 1), dtype=bool)
    errors = np.zeros(N)
    points_2d = np.full((C, N, 2), np.nan)
    subset_index = np.full(N, -1, dtype=np.int64)
    n_eval = np.zeros(N, dtype=np.int64)

    U = undistort_points(cams, points) if undistort else points
    Ms = np.array([cam.extrinsics() for cam in cams]) if C else np.empty((0, 4, 4))
    valid = ~np.isnan(points[:, :, 0])                 # validity on RAW x (:658-659)
    usable = valid & ~np.isnan(U[:, :, 0])             # survives inside triangulate (:630)
    weights = (1 << np.arange(C, dtype=np.int64))[:, None]
    key = (valid * weights).sum(axis=0) + ((usable * weights).sum(axis=0) << C)

    for m in np.unique(key):
        idx = np.nonzero(key == m)[0]
        vmask = int(m) & ((1 << C) - 1)
        umask = int(m) >> C
        V = [c for c in range(C) if (vmask >> c) & 1]
        k = len(V)
        n = idx.size
        best_err = np.full(n, float(init_best))
        best_s = np.full(n, -1, dtype=np.int64)
        best_X = np.full((n, 3), np.nan)
        done = np.zeros(n, dtype=bool)
        for s in range(1 << k):
            S = subset_cameras_of(V, s)
            if len(S) < min_cams and len(S) != k:
                continue
            live = ~done
            if not live.any():
                break
            li = np.nonzero(live)[0]
            n_eval[idx[li]] += 1
            Su = [c for c in S if (umask >> c) & 1]
            if len(Su) < 2:
                continue                               # X = nan -> err = nan -> never accepted
            pid = idx[li]
            X = _dlt_solve(_dlt_rows(U[Su][:, pid], Ms[Su]))
            # mean reprojection error against RAW points over S (:701); cameras whose
            # residual is NaN (y missing) drop out of numerator and count.
            with np.errstate(invalid="ignore", divide="ignore"):
                tot = np.zeros(li.size)
                cnt = np.zeros(li.size)
                for c in S:
                    e = points[c, pid] - cams[c].project(X)
                    nr = np.sqrt(e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1])
                    g = ~np.isnan(nr)
                    tot += np.where(g, nr, 0.0)
                    cnt += g
                cnt[cnt < 1.5] = np.nan
                err = tot / cnt
                acc = err < best_err[li]
            a = li[acc]
            best_err[a] = err[acc]
            best_s[a] = s
            best_X[a] = X[acc]
            done[a] = best_err[a] < threshold
        sel = best_s >= 0
        pid = idx[sel]
        out[pid] = best_X[sel]
        errors[pid] = best_err[sel]
        subset_index[pid] = best_s[sel]
        for j, c in enumerate(V):
            inc = sel & (((best_s >> (k - 1 - j)) & 1) == 0)
            picked[c, idx[inc], 0] = True
            points_2d[c, idx[inc]] = points[c, idx[inc]]
    if return_stats:
        return out, picked, points_2d, errors, subset_index, n_eval
    return out, picked, points_2d, errors


def triangulate_possible(cams, points, undistort=True, min_cams=2, threshold=0.5, init_best=200.0,
                         return_stats=False):
    """CameraGroup.triangulate_possible for P candidates per camera (cameras.py:639-724), in the
    reference's own loop structure: per point, itertools.product over the cameras that have a
    valid candidate (ascending), each offering its valid candidates (ascending) and then None.

    Returns (out (N,3), picked (C,N,P) bool, points_2d (C,N,2), errors (N,)) and, with
    return_stats, (index (N,) of the accepted combination in product order, -1 = none,
    n_evaluated (N,))."""
    import itertools
    points = np.asarray(points, dtype=np.float64)
    assert points.shape[0] == len(cams), \
        "Invalid points shape, first dim should be equal to" \
        " number of cameras ({}), but shape is {}".format(len(cams), points.shape)
    C, N, P, _ = points.shape
    out = np.full((N, 3), np.nan)
    picked_vals = np.zeros((C, N, P), dtype=bool)
    errors = np.zeros(N)
    points_2d = np.full((C, N, 2), np.nan)
    index = np.full(N, -1, dtype=np.int64)
    n_eval = np.zeros(N, dtype=np.int64)
    for n in range(N):
        options = []
        for c in range(C):
            cand = [(c, x) for x in range(P) if not np.isnan(points[c, n, x, 0])]      # :658-671
            if cand:
                options.append(cand + [None])
        n_cams_max = len(options)
        best_error, best = float(init_best), None
        for ix, combo in enumerate(itertools.product(*options)):
            pk = [q for q in combo if q is not None]
            if len(pk) < min_cams and len(pk) != n_cams_max:                           # :691
                continue
            n_eval[n] += 1
            cn = [q[0] for q in pk]
            xn = [q[1] for q in pk]
            pts = points[cn, n, xn] if pk else np.zeros((0, 2))
            sub = [cams[c] for c in cn]
            with np.errstate(all="ignore"):
                p3d = triangulate(sub, pts[:, None, :], undistort=undistort)[0] if len(pk) else np.full(3, np.nan)
                err = reprojection_error(sub, p3d[None], pts[:, None, :], mean=True)[0] if len(pk) else np.nan
            if err < best_error:                                                       # :703
                best = (p3d, pts, pk, err, ix)
                best_error = err
                if best_error < threshold:                                             # :712
                    break
        if best is not None:
            p3d, pts, pk, err, ix = best
            out[n] = p3d
            cn = [q[0] for q in pk]
            xn = [q[1] for q in pk]
            picked_vals[cn, n, xn] = True
            errors[n] = err
            points_2d[cn, n] = pts
            index[n] = ix
    if return_stats:
        return out, picked_vals, points_2d, errors, index, n_eval
    return out, picked_vals, points_2d, errors


# ----------------------------------------------------------------------------
# loop-faithful port (CPU baseline: same call structure as the reference)
# ----------------------------------------------------------------------------

def _cv2():
    import cv2
    return cv2


def _cv_undistort(cam, pts):
    cv2 = _cv2()
    p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 1, 2)
    if cam.model == cm.MODEL_PINHOLE:
        o = cv2.undistortPoints(p, cam.K, cam.dist)
    elif cam.model == cm.MODEL_FISHEYE:
        o = cv2.fisheye.undistortPoints(p, cam.K, cam.dist)
    else:
        return cam.undistort(pts)
    return o.reshape(np.shape(pts))


def _cv_project(cam, p3d):
    cv2 = _cv2()
    X = np.ascontiguousarray(p3d, dtype=np.float64).reshape(-1, 1, 3)
    if cam.model == cm.MODEL_PINHOLE:
        o, _ = cv2.projectPoints(X, cam.rvec, cam.tvec, cam.K, cam.dist)
    elif cam.model == cm.MODEL_FISHEYE:
        o, _ = cv2.fisheye.projectPoints(X, cam.rvec, cam.tvec, cam.K, cam.dist)
    else:
        return cam.project(p3d).reshape(-1, 1, 2)
    return o


def _svd_point(pts, mats):
    n = len(mats)
    A = np.zeros((n * 2, 4))
    for i in range(n):
        x, y = pts[i]
        A[2 * i] = x * mats[i][2] - mats[i][0]
        A[2 * i + 1] = y * mats[i][2] - mats[i][1]
    _, _, vh = np.linalg.svd(A, full_matrices=True)
    p = vh[-1]
    return p[:3] / p[3]


def triangulate_loops(cams, points, undistort=True):
    """One LAPACK call per point, like the loop at cameras.py:628-632."""
    points = np.asarray(points, dtype=np.float64)
    one_point = points.ndim == 2
    if one_point:
        points = points.reshape(-1, 1, 2)
    if undistort:
        new = np.empty(points.shape)
        for c, cam in enumerate(cams):
            new[c] = _cv_undistort(cam, np.copy(points[c]))
        points = new
    n_points = points.shape[1]
    out = np.full((n_points, 3), np.nan)
    mats = np.array([cam.extrinsics() for cam in cams]).reshape(-1, 4, 4)
    for ip in range(n_points):
        sub = points[:, ip, :]
        good = ~np.isnan(sub[:, 0])
        if np.sum(good) >= 2:
            out[ip] = _svd_point(sub[good], mats[good])
    return out[0] if one_point else out


def reprojection_error_loops(cams, p3ds, p2ds, mean=False):
    """Per-camera OpenCV projection, like cameras.py:764-775."""
    p3ds = np.asarray(p3ds, dtype=np.float64)
    p2ds = np.asarray(p2ds, dtype=np.float64)
    one_point = p3ds.ndim == 1 and p2ds.ndim == 2
    if one_point:
        p3ds = p3ds.reshape(1, 3)
        p2ds = p2ds.reshape(-1, 1, 2)
    n_cams, n_points, _ = p2ds.shape
    errors = np.empty((n_cams, n_points, 2))
    for c, cam in enumerate(cams):
        errors[c] = p2ds[c] - _cv_project(cam, p3ds).reshape(n_points, 2)
    if mean:
        with np.errstate(invalid="ignore", divide="ignore"):
            nrm = np.linalg.norm(errors, axis=2)
            good = ~np.isnan(nrm)
            nrm[~good] = 0
            denom = np.sum(good, axis=0).astype("float64")
            denom[denom < 1.5] = np.nan
            errors = np.sum(nrm, axis=0) / denom
    if one_point:
        return float(errors[0]) if mean else errors.reshape(-1, 2)
    return errors


def triangulate_ransac_loops(cams, points, undistort=True, min_cams=2, threshold=0.5):
    """Point-by-point, subset-by-subset search with one-point OpenCV / LAPACK
    calls per subset, like cameras.py:683-722."""
    points = np.asarray(points, dtype=np.float64)
    C, N, _ = points.shape
    out = np.full((N, 3), np.nan)
    picked = np.zeros((C, N, 1), dtype=bool)
    errors = np.zeros(N)
    points_2d = np.full((C, N, 2), np.nan)
    for ip in range(N):
        V = [c for c in range(C) if not np.isnan(points[c, ip, 0])]
        options = [((c,), ()) for c in V]
        best = None
        best_error = 200
        for choice in itertools.product(*options):
            S = [c for t in choice for c in t]
            if len(S) < min_cams and len(S) != len(V):
                continue
            sub = [cams[c] for c in S]
            pts = points[S, ip]
            p3d = triangulate_loops(sub, pts, undistort=undistort)
            err = reprojection_error_loops(sub, p3d, pts, mean=True)
            if err < best_error:
                best = (S, p3d, err, pts)
                best_error = err
                if best_error < threshold:
                    break
        if best is not None:
            S, p3d, err, pts = best
            out[ip] = p3d
            picked[S, ip, 0] = True
            errors[ip] = err
            points_2d[S, ip] = pts
    return out, picked, points_2d, errors
