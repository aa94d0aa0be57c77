"""Oracle (test infrastructure, see oracle/__init__.py): numpy restatement of the
OpenCV point maps the reference calls per camera.

Reference call sites (all under /root/reference/src/third_party/aniposelib/):
  Camera.undistort_points          cameras.py:310-316  -> cv2.undistortPoints(pts, K, dist)
  Camera.project                   cameras.py:318-323  -> cv2.projectPoints(X, rvec, tvec, K, dist)
  FisheyeCamera.undistort_points   cameras.py:376-382  -> cv2.fisheye.undistortPoints(pts, K, D)
  FisheyeCamera.project            cameras.py:384-390  -> cv2.fisheye.projectPoints(X, rvec, tvec, K, D)
  OmnidirCamera.undistort_points   cameras.py:498-507  -> cv2.omnidir.undistortPoints(pts, K, D, xi, I)
  OmnidirCamera.project            cameras.py:509-516  -> cv2.omnidir.projectPoints(X, rvec, tvec, K, xi, D)
  make_M                           utils.py:9-15       -> cv2.Rodrigues

Third-party arithmetic restated here: OpenCV calib3d (reference pins
opencv-python 4.11.0.86, getting_started.md:25; checked here against the
installed 4.13.0 in tests/test_oracle_golden.py) and opencv_contrib ccalib
omnidir (pinned opencv-contrib-python 4.11.0.86; NOT installed in this image ->
the omnidir functions below follow the published algorithm and are
"parity unpinned").

All functions are vectorised over points, float64, and never mutate inputs.
"""
import numpy as np

MODEL_PINHOLE = 0
MODEL_FISHEYE = 1
MODEL_OMNIDIR = 2


def rodrigues(rvec):
    """cv2.Rodrigues(rvec)[0]: R = cos(t) I + (1-cos t) r r^T + sin(t) [r]x ;
    |rvec| < DBL_EPSILON -> I.  (utils.py:11)"""
    r = np.asarray(rvec, dtype=np.float64).ravel()
    theta = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    itheta = 1.0 / theta
    rx, ry, rz = r * itheta
    rrt = np.array([[rx * rx, rx * ry, rx * rz],
                    [rx * ry, ry * ry, ry * rz],
                    [rx * rz, ry * rz, rz * rz]])
    r_x = np.array([[0, -rz, ry], [rz, 0, -rx], [-ry, rx, 0]])
    return c * np.eye(3) + c1 * rrt + s * r_x


def make_M(rvec, tvec):
    """4x4 extrinsics [R|t; 0 0 0 1]  (utils.py:9-15)."""
    out = np.zeros((4, 4))
    out[:3, :3] = rodrigues(rvec)
    out[:3, 3] = np.asarray(tvec, dtype=np.float64).ravel()
    out[3, 3] = 1
    return out


def _k14(dist):
    """OpenCV zero-pads the distortion vector to 14 entries
    (k1,k2,p1,p2,k3,k4,k5,k6,s1,s2,s3,s4,tauX,tauY)."""
    d = np.asarray(dist, dtype=np.float64).ravel()
    if d.size not in (4, 5, 8, 12, 14):
        raise ValueError("distortion vector must have 4, 5, 8, 12 or 14 entries")
    k = np.zeros(14)
    k[:d.size] = d
    if k[12] != 0.0 or k[13] != 0.0:
        raise NotImplementedError("tilted sensor model (tauX, tauY) not restated")
    return k


def undistort_pinhole(points, K, dist, iters=5):
    """cv2.undistortPoints(points, K, dist) with no R / P: exactly ``iters`` (=5,
    the TermCriteria(MAX_ITER, 5) default) fixed-point iterations, with the
    ``icdist < 0`` bail-out that restores the initial guess.  Skew is ignored by
    OpenCV.  points: (..., 2) pixels -> (..., 2) normalised coordinates."""
    pts = np.asarray(points, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    k = _k14(dist)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    u, v = pts[..., 0], pts[..., 1]
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        x0 = (u - cx) * ifx
        y0 = (v - cy) * ify
        x, y = x0.copy(), y0.copy()
        active = np.ones(x.shape, dtype=bool)
        for _ in range(iters):
            r2 = x * x + y * y
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / \
                     (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            bail = active & (icdist < 0)
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
            xn = (x0 - dx) * icdist
            yn = (y0 - dy) * icdist
            upd = active & ~bail
            x = np.where(upd, xn, np.where(bail, x0, x))
            y = np.where(upd, yn, np.where(bail, y0, y))
            active = upd
    return np.stack([x, y], axis=-1)


def _to_camera(p3d, R, t):
    X = np.asarray(p3d, dtype=np.float64).reshape(-1, 3)
    xc = R[0, 0] * X[:, 0] + R[0, 1] * X[:, 1] + R[0, 2] * X[:, 2] + t[0]
    yc = R[1, 0] * X[:, 0] + R[1, 1] * X[:, 1] + R[1, 2] * X[:, 2] + t[1]
    zc = R[2, 0] * X[:, 0] + R[2, 1] * X[:, 1] + R[2, 2] * X[:, 2] + t[2]
    return xc, yc, zc


def project_pinhole(p3d, rvec, tvec, K, dist):
    """cv2.projectPoints(p3d, rvec, tvec, K, dist)[0] -> (N, 2).  z == 0 -> 1/z := 1;
    points behind the camera are NOT rejected."""
    K = np.asarray(K, dtype=np.float64)
    k = _k14(dist)
    R = rodrigues(rvec)
    t = np.asarray(tvec, dtype=np.float64).ravel()
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        xc, yc, zc = _to_camera(p3d, R, t)
        iz = np.where(zc != 0, 1.0 / np.where(zc != 0, zc, 1.0), 1.0)
        iz = np.where(np.isnan(zc), np.nan, iz)
        x = xc * iz
        y = yc * iz
        r2 = x * x + y * y
        r4 = r2 * r2
        r6 = r4 * r2
        a1 = 2 * x * y
        a2 = r2 + 2 * x * x
        a3 = r2 + 2 * y * y
        cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6
        icdist2 = 1.0 / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6)
        xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4
        yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4
        u = xd * fx + cx
        v = yd * fy + cy
    return np.stack([u, v], axis=-1)


def undistort_fisheye(points, K, D):
    """cv2.fisheye.undistortPoints(points, K, D) (R = I, no P), OpenCV >= 4.5
    semantics: Newton on theta, <= 10 iterations, stop |fix| < 1e-8; a
    non-converged or sign-flipped theta yields (-1e6, -1e6)."""
    pts = np.asarray(points, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    k = np.zeros(4)
    d = np.asarray(D, dtype=np.float64).ravel()
    k[:min(4, d.size)] = d[:4]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    alpha = K[0, 1] / fx
    shape = pts.shape
    P = pts.reshape(-1, 2)
    out = np.empty_like(P)
    half_pi = np.pi / 2.0
    for i in range(P.shape[0]):
        pw1 = (P[i, 1] - cy) / fy
        pw0 = (P[i, 0] - cx) / fx
        if alpha != 0.0:
            pw0 = pw0 - alpha * pw1
        theta_d = np.sqrt(pw0 * pw0 + pw1 * pw1)
        # std::min(std::max(-pi/2, theta_d), pi/2): a NaN theta_d becomes -pi/2
        theta_d = theta_d if (-half_pi < theta_d) else -half_pi
        theta_d = theta_d if (theta_d < half_pi) else half_pi
        converged = False
        theta = theta_d
        scale = 0.0
        if abs(theta_d) > 1e-8:
            for _ in range(10):
                t2 = theta * theta
                t4 = t2 * t2
                t6 = t4 * t2
                t8 = t6 * t2
                k0t2, k1t4, k2t6, k3t8 = k[0] * t2, k[1] * t4, k[2] * t6, k[3] * t8
                fix = (theta * (1 + k0t2 + k1t4 + k2t6 + k3t8) - theta_d) / \
                      (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8)
                theta = theta - fix
                if abs(fix) < 1e-8:
                    converged = True
                    break
            scale = np.tan(theta) / theta_d
        else:
            converged = True
        flipped = (theta_d < 0 and theta > 0) or (theta_d > 0 and theta < 0)
        if converged and not flipped:
            out[i, 0] = pw0 * scale
            out[i, 1] = pw1 * scale
        else:
            out[i, 0] = -1000000.0
            out[i, 1] = -1000000.0
    return out.reshape(shape)


def project_fisheye(p3d, rvec, tvec, K, D):
    """cv2.fisheye.projectPoints(p3d, rvec, tvec, K, D)[0] -> (N, 2)."""
    K = np.asarray(K, dtype=np.float64)
    k = np.zeros(4)
    d = np.asarray(D, dtype=np.float64).ravel()
    k[:min(4, d.size)] = d[:4]
    R = rodrigues(rvec)
    t = np.asarray(tvec, dtype=np.float64).ravel()
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    alpha = K[0, 1] / fx
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        xc, yc, zc = _to_camera(p3d, R, t)
        x = xc / zc
        y = yc / zc
        r2 = x * x + y * y
        r = np.sqrt(r2)
        theta = np.arctan(r)
        t2 = theta * theta
        t3 = t2 * theta
        t5 = t3 * t2
        t7 = t5 * t2
        t9 = t7 * t2
        theta_d = theta + k[0] * t3 + k[1] * t5 + k[2] * t7 + k[3] * t9
        inv_r = np.where(r > 1e-8, 1.0 / np.where(r > 1e-8, r, 1.0), 1.0)
        cdist = np.where(r > 1e-8, theta_d * inv_r, 1.0)
        cdist = np.where(np.isnan(r), np.nan, cdist)
        xd1 = x * cdist
        xd2 = y * cdist
        u = fx * (xd1 + alpha * xd2) + cx
        v = fy * xd2 + cy
    return np.stack([u, v], axis=-1)


def undistort_omnidir(points, K, D, xi):
    """cv2.omnidir.undistortPoints(points, K, D, xi, R=I) — Mei unified model.
    PARITY UNPINNED: restated from the published opencv_contrib ccalib algorithm
    (20 fixed-point iterations with in-place x-then-y update, lift to the unit
    sphere with xi, perspective re-projection)."""
    pts = np.asarray(points, dtype=np.float64)
    K = np.asarray(K, dtype=np.float64)
    d = np.asarray(D, dtype=np.float64).ravel()
    k1, k2, p1, p2 = d[0], d[1], d[2], d[3]
    xi = float(np.asarray(xi, dtype=np.float64).ravel()[0])
    fx, fy, cx, cy, s = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]
    u, v = pts[..., 0], pts[..., 1]
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        ppx = (u * fy - cx * fy - s * (v - cy)) / (fx * fy)
        ppy = (v - cy) / fy
        pux, puy = ppx.copy(), ppy.copy()
        for _ in range(20):
            r2 = pux * pux + puy * puy
            r4 = r2 * r2
            den = 1 + k1 * r2 + k2 * r4
            pux = (ppx - 2 * p1 * pux * puy - p2 * (r2 + 2 * pux * pux)) / den
            puy = (ppy - 2 * p2 * pux * puy - p1 * (r2 + 2 * puy * puy)) / den
        r2 = pux * pux + puy * puy
        a = r2 + 1
        b = 2 * xi * r2
        cc = r2 * xi * xi - 1
        Zs = (-b + np.sqrt(b * b - 4 * a * cc)) / (2 * a)
        Xw0 = pux * (Zs + xi)
        Xw1 = puy * (Zs + xi)
        Xw2 = Zs
        nrm = np.sqrt(Xw0 * Xw0 + Xw1 * Xw1 + Xw2 * Xw2)
        Xs0, Xs1, Xs2 = Xw0 / nrm, Xw1 / nrm, Xw2 / nrm
        x = Xs0 / Xs2
        y = Xs1 / Xs2
    return np.stack([x, y], axis=-1)


def project_omnidir(p3d, rvec, tvec, K, xi, D):
    """cv2.omnidir.projectPoints(p3d, rvec, tvec, K, xi, D)[0] -> (N, 2).
    PARITY UNPINNED (see undistort_omnidir)."""
    K = np.asarray(K, dtype=np.float64)
    d = np.asarray(D, dtype=np.float64).ravel()
    k1, k2, p1, p2 = d[0], d[1], d[2], d[3]
    xi = float(np.asarray(xi, dtype=np.float64).ravel()[0])
    R = rodrigues(rvec)
    t = np.asarray(tvec, dtype=np.float64).ravel()
    fx, fy, cx, cy, s = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        xc, yc, zc = _to_camera(p3d, R, t)
        nrm = np.sqrt(xc * xc + yc * yc + zc * zc)
        Xs0, Xs1, Xs2 = xc / nrm, yc / nrm, zc / nrm
        xu = Xs0 / (Xs2 + xi)
        yu = Xs1 / (Xs2 + xi)
        r2 = xu * xu + yu * yu
        r4 = r2 * r2
        rad = 1 + k1 * r2 + k2 * r4
        xd = xu * rad + 2 * p1 * xu * yu + p2 * (r2 + 2 * xu * xu)
        yd = yu * rad + p1 * (r2 + 2 * yu * yu) + 2 * p2 * xu * yu
        u = fx * xd + s * yd + cx
        v = fy * yd + cy
    return np.stack([u, v], axis=-1)


class CamSpec:
    """Plain parameter record for one camera (oracle-side twin of the reference's
    Camera / FisheyeCamera / OmnidirCamera objects, cameras.py:173-556)."""

    def __init__(self, model, K, dist, rvec, tvec, xi=0.0, name=None):
        self.model = int(model)
        self.K = np.array(K, dtype=np.float64).reshape(3, 3)
        self.dist = np.array(dist, dtype=np.float64).ravel()
        self.rvec = np.array(rvec, dtype=np.float64).ravel()
        self.tvec = np.array(tvec, dtype=np.float64).ravel()
        self.xi = float(np.asarray(xi, dtype=np.float64).ravel()[0])
        self.name = name

    def undistort(self, pts):
        if self.model == MODEL_PINHOLE:
            return undistort_pinhole(pts, self.K, self.dist)
        if self.model == MODEL_FISHEYE:
            return undistort_fisheye(pts, self.K, self.dist)
        return undistort_omnidir(pts, self.K, self.dist, self.xi)

    def project(self, p3d):
        if self.model == MODEL_PINHOLE:
            return project_pinhole(p3d, self.rvec, self.tvec, self.K, self.dist)
        if self.model == MODEL_FISHEYE:
            return project_fisheye(p3d, self.rvec, self.tvec, self.K, self.dist)
        return project_omnidir(p3d, self.rvec, self.tvec, self.K, self.xi, self.dist)

    def extrinsics(self):
        return make_M(self.rvec, self.tvec)
