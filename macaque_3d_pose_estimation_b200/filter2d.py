"""2D keypoint Viterbi filter of the step-4 stage on the GPU: drop-in for
``src/third_party/anipose/filter_pose.py`` ``filter_pose_viterbi`` / ``wrap_points`` and for the
filter loop of ``src/pipeline/step4_aniposefiltering.py:144-170``.  Every series of a call runs
in ONE launch of ``k_viterbi`` (csrc/m3d_viterbi.cu) through ``m3d_viterbi_filter``; there is no
CPU fallback."""
import numpy as np

from . import _lib
from ._device import _default_device, _ptr, _stream, torch

DUP_THRES = 5.0          # viterbi_path calls remove_dups(points, thres=5)  (filter_pose.py:51)


def viterbi_series(cand, n_back=3, thres_dist=30.0, score_threshold=-np.inf, device=None, return_choice=False):
    """viterbi_path (filter_pose.py:48-120) for S independent series at once.

    cand: (S, F, P, 3) array or CUDA tensor of x, y, score.  Returns points (S, F, 2), scores
    (S, F) [and choice (S, F) int32: age * P + candidate, -1 = missing-point particle]."""
    lib = _lib.require_gpu()
    device = _default_device() if device is None else device
    like_torch = torch is not None and isinstance(cand, torch.Tensor)
    t = cand if like_torch else torch.from_numpy(np.ascontiguousarray(cand, dtype=np.float64))
    t = t.to("cuda:%d" % device, torch.float64).contiguous()
    assert t.dim() == 4 and t.shape[3] == 3, "cand must have shape (S, F, P, 3), but got {}".format(tuple(t.shape))
    S, F, P = int(t.shape[0]), int(t.shape[1]), int(t.shape[2])
    out = torch.empty((S, F, 3), dtype=torch.float64, device=t.device)
    choice = torch.empty((S, F), dtype=torch.int32, device=t.device) if return_choice else None
    with torch.cuda.device(device):
        _lib.check(lib.m3d_viterbi_filter(_ptr(t), S, F, P, int(n_back), float(thres_dist), float(score_threshold),
                                          DUP_THRES, _ptr(out), _ptr(choice), int(device), _stream(device)),
                   "m3d_viterbi_filter")
    pts, sc = out[..., :2], out[..., 2]
    if not like_torch:
        pts, sc = pts.cpu().numpy(), sc.cpu().numpy()
        choice = choice.cpu().numpy() if return_choice else None
    return (pts, sc, choice) if return_choice else (pts, sc)


def filter_pose_viterbi(config, all_points, bodyparts=None):
    """filter_pose.py:151-186: all_points (F, J, P, 3) -> points (F, J, 2), scores (F, J).  Like the
    reference, the score threshold is written into ``all_points`` in place (:157)."""
    n_frames, n_joints, n_possible, _ = all_points.shape
    points_full = all_points[:, :, :, :2]
    scores_full = all_points[:, :, :, 2]
    points_full[scores_full < config['filter']['score_threshold']] = np.nan
    cand = np.ascontiguousarray(np.transpose(all_points, (1, 0, 2, 3)))          # (J, F, P, 3)
    pts, sc = viterbi_series(cand, config['filter']['n_back'], config['filter']['offset_threshold'],
                             config['filter']['score_threshold'])
    return np.ascontiguousarray(pts.transpose(1, 0, 2)), np.ascontiguousarray(sc.T)


def wrap_points(points, scores):
    """filter_pose.py:332-343."""
    if len(points.shape) == 3:
        points = points[:, :, None]
        scores = scores[:, :, None]
    n_frames, n_joints, n_possible, _ = points.shape
    all_points = np.full((n_frames, n_joints, n_possible, 3), np.nan, dtype='float64')
    all_points[:, :, :, :2] = points
    all_points[:, :, :, 2] = scores
    return all_points


STEP4_FILTER = {'score_threshold': 0.3, 'n_back': 3, 'offset_threshold': 25}    # step4:146-150


def filter_stage(kp2d, config=None):
    """The 2D filter loop of step4_aniposefiltering.proc (:144-167): kp2d (A, F, C, J, 3) as stored in
    kp2d.pickle -> kp2d_f (F, J, A, 3, C) as stored in kp2d_f.pickle, all A * C * J series in one launch."""
    f = dict(STEP4_FILTER if config is None else config['filter'])
    kp2d = np.asarray(kp2d, dtype=np.float64)
    A, F, C, J, _ = kp2d.shape
    cand = np.ascontiguousarray(kp2d.transpose(0, 2, 3, 1, 4)).reshape(A * C * J, F, 1, 3)   # series (a, c, j)
    pts, sc = viterbi_series(cand, f['n_back'], f['offset_threshold'], f['score_threshold'])
    res = np.concatenate([pts, sc[..., None]], axis=-1).reshape(A, C, J, F, 3)
    return np.ascontiguousarray(res.transpose(3, 2, 0, 4, 1))                                # (F, J, A, 3, C)
