"""Seeded synthetic rigs and keypoint tracks for tests, golden
vectors and bench.py.  Host-side numpy only; no arithmetic of the hot path lives
here — 2D observations are produced by whatever ``project`` the caller passes
(the GPU CameraGroup in bench.py, the reference / oracle when making goldens).

Rig geometry follows the reference's templates: 8 cameras named "1".."C",
2048x1536 images (configs/calibration_tmpl.toml:1-83), intrinsics in the range
printed in notebooks/bbox_optimisation_algorithm.ipynb, cage about 2.2 m wide.
"""
import numpy as np

IMG_SIZE = (2048, 1536)
N_JOINTS = 17          # macaque model, model/pose/macaque.py


def _rodrigues_inv(R):
    """Rotation matrix -> rotation vector (host helper for rig construction)."""
    tr = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    theta = np.arccos(tr)
    if theta < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if np.pi - theta < 1e-6:
        # near pi: take axis from the symmetric part
        B = (R + np.eye(3)) / 2.0
        axis = np.sqrt(np.clip(np.diag(B), 0, None))
        i = int(np.argmax(axis))
        axis = B[i] / axis[i]
        axis /= np.linalg.norm(axis)
        if np.dot(axis, w) < 0:
            axis = -axis
        return axis * theta
    return w / (2.0 * np.sin(theta)) * theta


def make_rig(n_cams=8, model="pinhole", seed=20261018, radius=2000.0):
    """List of camera dicts with the keys of the reference's ``Camera.get_dict``
    (cameras.py:191-199; + 'fisheye' :363, + 'omnidir','xi','K','D' :442-451)."""
    rng = np.random.default_rng(seed)
    cams = []
    for i in range(n_cams):
        phi = 2.0 * np.pi * i / n_cams
        centre = np.array([radius * np.cos(phi), radius * np.sin(phi), 500.0 + 100.0 * (i % 3)])
        z = -centre / np.linalg.norm(centre)
        up = np.array([0.0, 0.0, 1.0])
        x = np.cross(z, up)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])
        # small random roll / pointing error so no two cameras are mirror images
        d = rng.normal(0, 0.02, size=3)
        th = np.linalg.norm(d)
        kx = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        dR = np.eye(3) + np.sin(th) / th * kx + (1 - np.cos(th)) / th ** 2 * (kx @ kx)
        R = dR @ R
        tvec = -R @ centre
        fx = rng.uniform(1190, 1290)
        fy = fx * rng.uniform(0.995, 1.005)
        cx = 1030 + rng.uniform(-30, 30)
        cy = 725 + rng.uniform(-35, 35)
        cam = {
            "name": str(i + 1),
            "size": list(IMG_SIZE),
            "matrix": [[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]],
            "rotation": _rodrigues_inv(R).tolist(),
            "translation": tvec.tolist(),
        }
        if model == "pinhole":
            cam["distortions"] = [rng.uniform(-0.25, -0.05), rng.uniform(0, 0.08),
                                  rng.normal(0, 1e-3), rng.normal(0, 1e-3), rng.uniform(0, 0.01)]
        elif model == "pinhole8":
            cam["distortions"] = [rng.uniform(-0.25, -0.05), rng.uniform(0, 0.08),
                                  rng.normal(0, 1e-3), rng.normal(0, 1e-3), rng.uniform(0, 0.01),
                                  rng.uniform(0, 0.02), rng.uniform(0, 0.01), rng.uniform(0, 0.002)]
        elif model == "fisheye":
            base = np.array([0.05, -0.01, 0.002, -0.0005])
            cam["distortions"] = (base * rng.uniform(0.5, 1.5, size=4)).tolist()
            cam["fisheye"] = True
        elif model == "omnidir":
            xi = rng.uniform(0.9, 1.6)
            f = 1240.0 * (1.0 + xi)
            cam["distortions"] = [0.0, 0.0, 0.0, 0.0]
            cam["omnidir"] = True
            cam["xi"] = [xi]
            cam["K"] = [[f * rng.uniform(0.98, 1.02), rng.uniform(-8, 16), cx],
                        [0.0, f * rng.uniform(0.98, 1.02), cy], [0.0, 0.0, 1.0]]
            cam["D"] = [rng.uniform(-0.2, -0.05), rng.uniform(0, 0.05),
                        rng.normal(0, 5e-4), rng.normal(0, 5e-4)]
        else:
            raise ValueError("unknown camera model " + str(model))
        cams.append(cam)
    return cams


def make_tracks(n_frames, n_animals, n_joints=N_JOINTS, seed=20261018):
    """3D joint tracks (F, A, J, 3) in mm: per-animal root random walk inside the
    cage plus a fixed skeleton offset per joint."""
    rng = np.random.default_rng(seed + 1)
    root0 = rng.uniform([-900, -900, 0], [900, 900, 1500], size=(n_animals, 3))
    steps = rng.normal(0, 15.0, size=(n_frames, n_animals, 3))
    root = root0[None] + np.cumsum(steps, axis=0)
    root = np.clip(root, [-1000, -1000, 0], [1000, 1000, 1600])
    skel = rng.normal(0, 120.0, size=(n_animals, n_joints, 3))
    return root[:, :, None, :] + skel[None]


def corrupt(p2d, seed=20261018, noise=0.3, p_outlier=0.0, sigma_outlier=60.0, p_missing=0.0):
    """Add detector noise, gross outliers and missing views to clean (C, N, 2)
    projections.  Missing views are NaN in both coordinates (the score-threshold
    gate of step4_aniposefiltering.py:225-226)."""
    rng = np.random.default_rng(seed + 2)
    p = np.array(p2d, dtype=np.float64, copy=True)
    C, N, _ = p.shape
    p += rng.normal(0, noise, size=p.shape)
    if p_outlier > 0:
        o = rng.random((C, N)) < p_outlier
        p[o] += rng.normal(0, sigma_outlier, size=(int(o.sum()), 2))
    if p_missing > 0:
        m = rng.random((C, N)) < p_missing
        p[m] = np.nan
    return p


def make_detection_series(n_frames, n_joints, n_possible, seed, jump=0.05, low=0.1):
    """Synthetic 2D detections (F, J, P, 3) [x, y, score]: smooth tracks + detector noise, gross
    jumps, low-score frames; extra candidates are a near duplicate (within 5 px: remove_dups) or a
    distractor."""
    rng = np.random.default_rng(seed)
    F, J, P = n_frames, n_joints, n_possible
    pos = rng.uniform([300, 300], [1700, 1200], size=(J, 2))[None] + np.cumsum(rng.normal(0, 4.0, size=(F, J, 2)), axis=0)
    pts = np.zeros((F, J, P, 3))
    pts[:, :, 0, :2] = pos + rng.normal(0, 0.7, size=(F, J, 2))
    pts[:, :, 0, 2] = rng.uniform(0.35, 1.0, size=(F, J))
    jm = rng.random((F, J)) < jump
    pts[jm, 0, :2] += rng.normal(0, 60.0, size=(int(jm.sum()), 2))
    lw = rng.random((F, J)) < low
    pts[lw, 0, 2] = rng.uniform(0.0, 0.29, size=int(lw.sum()))
    for p in range(1, P):
        near = rng.random((F, J)) < 0.3
        pts[:, :, p, :2] = np.where(near[..., None], pts[:, :, 0, :2] + rng.normal(0, 1.5, size=(F, J, 2)),
                                    pos + rng.normal(0, 40.0, size=(F, J, 2)))
        pts[:, :, p, 2] = rng.uniform(0.1, 0.9, size=(F, J))
    # a stretch of missing detections at the start and in the middle of one series
    pts[:4, 0, :, 2] = 0.05
    pts[F // 2:F // 2 + 5, min(1, J - 1), :, 2] = 0.05
    return pts
