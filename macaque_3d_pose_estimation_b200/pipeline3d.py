"""The 3D stage of the reference pipeline (src/pipeline/step4_aniposefiltering.py:219-339)
on the GPU CameraGroup: score gating -> triangulate / triangulate_ransac ->
reprojection error -> per-joint score and error arrays.

All animals are reconstructed in ONE kernel call (the reference loops over animals and
points in Python, step4:219, cameras.py:628/683).  ``run_filter_stage`` / ``run_stage`` /
``run_step4`` are the file-to-file forms (kp2d.pickle -> kp2d_f.pickle -> kp3d.pickle); the
calibration assembly from the lab's HDF5 files (step4:101-138) stays reference territory.
The ``optim=True`` branch (CameraGroup.optim_points, step4:228-291) is outside the
accelerated path.
"""
import numpy as np


def reconstruct(cgroup, kp2d_f, score_threshold=0.5, ransac=False, optim=False, min_cams=3):
    """kp2d_f: (A, C, F, J, 3) float64 [x, y, score] — the array step4 holds after
    ``kp2d_f.transpose((2,4,0,1,3))`` (step4:190).  Like the reference, keypoints whose
    score is below the threshold are set to NaN IN PLACE and the scores of unused views are
    overwritten with 2 (step4:225-226, 287, 314).

    Returns {'kp3d': (A,F,J,3), 'kp3d_score': (A,F,J), 'kp3d_err': (A,F,J), 'num_cams': (A,F,J)}.
    ``min_cams`` is the value step4 passes to triangulate_ransac (:297).
    """
    if optim:
        raise NotImplementedError(
            "optim=True (CameraGroup.optim_points, step4_aniposefiltering.py:247-271) is outside "
            "the accelerated hot path; run with config['triangulation']['optim'] = false")
    n_animal, n_cam, n_frame, n_kp, _ = kp2d_f.shape
    assert n_cam == len(cgroup.cameras), \
        "kp2d_f has {} cameras, camera group has {}".format(n_cam, len(cgroup.cameras))
    all_points_raw = kp2d_f[..., :2]
    all_scores = kp2d_f[..., 2]
    bad = all_scores < score_threshold                                   # step4:225-226
    all_points_raw[bad] = np.nan

    # (A, C, F, J, 2) -> (C, A*F*J, 2): one launch for every animal
    pts = np.ascontiguousarray(all_points_raw.transpose(1, 0, 2, 3, 4)).reshape(n_cam, -1, 2)
    if ransac:
        p3d, picked, p2ds, errors = cgroup.triangulate_ransac(pts, min_cams=min_cams)   # step4:296-297
        picked_shaped = p2ds.reshape(n_cam, n_animal, n_frame, n_kp, 2)
        good = ~np.isnan(picked_shaped[..., 0])                           # step4:299-300
        num_cams = picked.sum(axis=0).sum(axis=1).reshape(n_animal, n_frame, n_kp).astype('float')
    else:
        p3d, errors = cgroup.triangulate_with_error(pts)                  # step4:306-307 fused
        good = ~np.isnan(pts.reshape(n_cam, n_animal, n_frame, n_kp, 2)[..., 0])
        num_cams = good.sum(axis=0).astype('float')                      # step4:308-309
    kp3d = p3d.reshape(n_animal, n_frame, n_kp, 3)
    E = errors.reshape(n_animal, n_frame, n_kp).copy()

    good_a = good.transpose(1, 0, 2, 3)                                  # (A, C, F, J)
    all_scores[~good_a] = 2                                              # step4:314
    S = np.min(all_scores, axis=1)                                       # step4:315
    few = num_cams < 2
    S[few] = np.nan                                                      # step4:317-319
    E[few] = np.nan
    num_cams[few] = np.nan
    return {'kp3d': kp3d, 'kp3d_score': S, 'kp3d_err': E, 'num_cams': num_cams}


def run_stage(result_dir, camera_ids, config=None, joint_len=None):
    """File-to-file form of the 3D stage: the part of step4_aniposefiltering.proc after the 2D
    filter (:172-339).  Reads ``<result_dir>/kp2d_f.pickle`` (array (F, J, A, 3, C) as written at
    :169-170), ``calibration.toml`` (:138, loaded like :212-213) and, unless ``config`` is given,
    ``config.toml`` (:102-104); writes ``kp3d.pickle`` with the reference's dict layout
    {kp3d (A,F,J,3), kp3d_score (A,F,J), kp3d_err (A,F,J), joint_len} (:332-339) and returns it.
    ``camera_ids`` is the camera_id list of calib/config.yaml (:107-110, 196-199)."""
    import os
    import pickle

    from .cameras import CameraGroup

    with open(os.path.join(result_dir, 'kp2d_f.pickle'), 'rb') as f:
        kp2d_f = pickle.load(f)
    if config is None:
        import toml
        config = toml.load(os.path.join(result_dir, 'config.toml'))
    tri = config['triangulation']
    cgroup = CameraGroup.load(os.path.join(result_dir, 'calibration.toml'))
    cgroup = cgroup.subset_cameras_names([str(i) for i in camera_ids])
    kp = np.ascontiguousarray(np.asarray(kp2d_f, dtype=np.float64).transpose((2, 4, 0, 1, 3)))   # (A, C, F, J, 3), :190
    res = reconstruct(cgroup, kp, score_threshold=tri['score_threshold'], ransac=bool(tri.get('ransac', False)),
                      optim=bool(tri.get('optim', False)))
    data = {'kp3d': res['kp3d'], 'kp3d_score': res['kp3d_score'], 'kp3d_err': res['kp3d_err'],
            'joint_len': [] if joint_len is None else joint_len}
    with open(os.path.join(result_dir, 'kp3d.pickle'), 'wb') as f:
        pickle.dump(data, f)
    return data


def run_filter_stage(result_dir):
    """File-to-file form of the 2D filter of step4_aniposefiltering.proc (:140-170): reads
    ``kp2d.pickle`` (array (A, F, C, J, 3)), runs the Viterbi filter with the constants of
    :146-150 on every (animal, camera, joint) series in one launch, writes ``kp2d_f.pickle``
    ((F, J, A, 3, C)) and returns it."""
    import os
    import pickle

    from . import filter2d

    with open(os.path.join(result_dir, 'kp2d.pickle'), 'rb') as f:
        kp2d = pickle.load(f)
    kp2d_f = filter2d.filter_stage(kp2d)
    with open(os.path.join(result_dir, 'kp2d_f.pickle'), 'wb') as f:
        pickle.dump(kp2d_f, f)
    return kp2d_f


def run_step4(result_dir, camera_ids, config=None, joint_len=None):
    """kp2d.pickle + calibration.toml (+ config.toml) -> kp2d_f.pickle, kp3d.pickle: everything of
    step4_aniposefiltering.proc after the calibration assembly (:140-339)."""
    run_filter_stage(result_dir)
    return run_stage(result_dir, camera_ids, config=config, joint_len=joint_len)
