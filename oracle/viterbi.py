"""Oracle for the 2D keypoint Viterbi filter of the step-4 stage.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``src/third_party/anipose/filter_pose.py`` with numpy / scipy:
  remove_dups            :26-46   (cKDTree pairs within 5 px of the same frame -> later one dropped)
  viterbi_path           :48-120  (particles = candidates of the last n_back frames, score
                                   halved per frame of age; log-space Viterbi; first-max ties)
  filter_pose_viterbi    :151-186 (score threshold, one series per joint; the multiprocessing
                                   Pool only distributes independent series)
  wrap_points            :332-343
Caller: src/pipeline/step4_aniposefiltering.py:144-170 (score_threshold 0.3, n_back 3,
offset_threshold 25, one candidate per frame, one call per (animal, camera)).

The transition log-probability is written in the closed form the reference's scipy calls reduce
to (scipy 1.18: stats.norm.logcdf -> special.log_ndtr(z) = log1p(-erfc(z / sqrt 2) / 2) for
z >= -1; special.logsumexp([hi, lo], b=[1, -1]) = log1p(-exp(lo - hi)) + hi, and -inf when the
two are equal); tests/test_oracle_golden.py pins it bit-for-bit against golden vectors produced
by executing the reference itself (oracle/make_golden.py, cases ``viterbi_*``).
"""
import numpy as np
from scipy import special

LOG_MISSING = np.log(0.001)


def remove_dups(pts, thres=5.0):
    """filter_pose.py:26-46 for candidates (F, P, 2): candidate q of a frame is dropped when an
    earlier candidate p < q of the same frame lies within ``thres`` (non-finite coordinates are
    moved to 1e9 first, so two missing candidates also 'collide' — a no-op)."""
    pts = np.asarray(pts, dtype=np.float64)
    q = np.where(np.isfinite(pts), pts, 1e9)
    out = pts.copy()
    P = pts.shape[1]
    for b in range(1, P):
        for a in range(b):
            d = np.sqrt(((q[:, a] - q[:, b]) ** 2).sum(axis=1))
            out[d <= thres, b] = np.nan
    return out


def log_ndtr(z):
    z = np.asarray(z, dtype=np.float64)
    t = z * np.sqrt(0.5)
    with np.errstate(all="ignore"):
        return np.where(z < -1.0, np.log(special.erfcx(-t) / 2) - t * t, np.log1p(-special.erfc(t) / 2))


def transition_logprob(dists, thres_dist):
    """filter_pose.py:89-94: log(Phi((d+2)/s) - Phi((d-2)/s)), floored at -100."""
    hi = log_ndtr((dists + 2) / thres_dist)
    lo = log_ndtr((dists - 2) / thres_dist)
    with np.errstate(all="ignore"):
        out = np.where(hi == lo, -np.inf, np.log1p(-np.exp(lo - hi)) + hi)
    out = np.where(out < -100, -100.0, out)
    return out


def build_particles(points, scores, n_back=3):
    """filter_pose.py:51-73.  Returns particles (F, n_max, 3), valid (F,), and src (F, n_max): the
    candidate each particle came from as age * P + index (-1 for the missing-point particle)."""
    F, P, _ = points.shape
    points_nans = remove_dups(points, thres=5)
    ok = ~np.isnan(points_nans[:, :, 0])
    n_max = int(ok.sum(axis=1).max()) * n_back + 1 if F else 1
    particles = np.zeros((F, n_max, 3))
    src = np.full((F, n_max), -1, dtype=np.int64)
    valid = np.zeros(F, dtype=np.int64)
    for i in range(F):
        s = 0
        for j in range(n_back):
            if i - j < 0:
                break
            ixs = np.where(ok[i - j])[0]
            n = len(ixs)
            particles[i, s:s + n, :2] = points[i - j, ixs]
            particles[i, s:s + n, 2] = scores[i - j, ixs] * np.power(2.0, -j)
            src[i, s:s + n] = j * P + ixs
            s += n
        if s == 0:
            particles[i, 0] = [-1, -1, 0.001]
            s = 1
        valid[i] = s
    return particles, valid, src


def viterbi_path(points, scores, n_back=3, thres_dist=30, return_choice=False):
    """filter_pose.py:48-120 for one series: points (F, P, 2), scores (F, P)."""
    points = np.asarray(points, dtype=np.float64)
    scores = np.asarray(scores, dtype=np.float64)
    F = points.shape[0]
    particles, valid, src = build_particles(points, scores, n_back)
    n_particles = int(valid.max())
    T = np.full((F, n_particles), -np.inf)
    back = np.zeros((F, n_particles), dtype=np.int64)
    with np.errstate(all="ignore"):
        T[0, :valid[0]] = np.log(particles[0, :valid[0], 2])
        back[0] = -1
        for i in range(1, F):
            va, vb = valid[i - 1], valid[i]
            pa = particles[i - 1, :va, :2]
            pb = particles[i, :vb, :2]
            d = pb[:, None, :] - pa[None, :, :]
            dists = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])     # (vb, va) = cdist(pa, pb).T
            P_trans = transition_logprob(dists, thres_dist)
            P_trans[pb[:, 0] == -1, :] = LOG_MISSING
            P_trans[:, pa[:, 0] == -1] = LOG_MISSING
            possible = T[i - 1, :va] + P_trans
            T[i, :vb] = np.max(possible, axis=1) + np.log(particles[i, :vb, 2])
            back[i, :vb] = np.argmax(possible, axis=1)
    out = np.zeros(F, dtype=np.int64)
    out[-1] = np.argmax(T[-1])
    for i in range(F - 1, 0, -1):
        out[i - 1] = back[i, out[i]]
    trace = particles[np.arange(F), out]
    if return_choice:
        return trace[:, :2], trace[:, 2], src[np.arange(F), out]
    return trace[:, :2], trace[:, 2]


def filter_pose_viterbi(config, all_points, bodyparts=None, return_choice=False):
    """filter_pose.py:151-186 (single process; like the reference it writes the score
    threshold into ``all_points`` in place)."""
    F, J, P, _ = all_points.shape
    points_full = all_points[:, :, :, :2]
    scores_full = all_points[:, :, :, 2]
    points_full[scores_full < config["filter"]["score_threshold"]] = np.nan
    points = np.full((F, J, 2), np.nan)
    scores = np.empty((F, J))
    choice = np.zeros((F, J), dtype=np.int64)
    for j in range(J):
        r = viterbi_path(points_full[:, j], scores_full[:, j], config["filter"]["n_back"],
                         config["filter"]["offset_threshold"], return_choice=True)
        points[:, j], scores[:, j], choice[:, j] = r
    if return_choice:
        return points, scores, choice
    return points, scores


def wrap_points(points, scores):
    """filter_pose.py:332-343."""
    if points.ndim == 3:
        points = points[:, :, None]
        scores = scores[:, :, None]
    F, J, P, _ = points.shape
    out = np.full((F, J, P, 3), np.nan)
    out[..., :2] = points
    out[..., 2] = scores
    return out
