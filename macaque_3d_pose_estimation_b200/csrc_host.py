"""Small host-side helpers that mirror arithmetic of csrc/ (no GPU needed)."""
import numpy as np


def rodrigues(rvec):
    """cv2.Rodrigues(rvec) -> (3,3): theta < DBL_EPSILON -> I, else cos I + (1 - cos) r r^T + sin [r]x
    (same formula as csrc/m3d_rig.h rodrigues, utils.py:9-15 make_M)."""
    r = np.asarray(rvec, dtype=np.float64).ravel()
    th = np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])
    if th < 2.220446049250313e-16:
        return np.eye(3)
    c, s = np.cos(th), np.sin(th)
    x, y, z = r / th
    rrt = np.array([[x * x, x * y, x * z], [x * y, y * y, y * z], [x * z, y * z, z * z]])
    rx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return c * np.eye(3) + (1 - c) * rrt + s * rx
