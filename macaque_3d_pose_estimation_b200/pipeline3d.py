"""The 3D stage of the reference pipeline (src/pipeline/step4_aniposefiltering.py:219-339)
on the GPU CameraGroup: score gating -> triangulate / triangulate_ransac ->
reprojection error -> per-joint score and error arrays.

All animals are reconstructed in ONE kernel call (the reference loops over animals and
points in Python, step4:219, cameras.py:628/683).  ``run_filter_stage`` / ``run_stage`` /
``run_step4`` are the file-to-file forms (kp2d.pickle -> kp2d_f.pickle -> kp3d.pickle).
The ``optim=True`` branch — the template's default, config_tmpl.toml:64-65 — runs the reference's
per-animal sequence (initial triangulation, ``optim_points`` or ``optim_points_jointlenfix``,
step4:228-291) on the GPU solver of csrc/m3d_optim.cu.
"""
import numpy as np

# macaque model (step4_aniposefiltering.py:198-202, model/pose/macaque.py)
BODYPARTS = ['nose', 'left_eye', 'right_eye', 'left_ear', 'right_ear', 'left_shoulder', 'right_shoulder',
             'left_elbow', 'right_elbow', 'left_wrist', 'right_wrist', 'left_hip', 'right_hip', 'left_knee',
             'right_knee', 'left_ankle', 'right_ankle']


def load_constraints(config, bodyparts, key='constraints'):
    """Joint-name pairs of config['triangulation'][key] -> joint-index pairs (step4:32-41)."""
    index = {bp: i for i, bp in enumerate(bodyparts)}
    out = []
    for a, b in config['triangulation'].get(key, []):
        for nm in (a, b):
            assert nm in index, 'Bodypart {} from constraints not found in list of bodyparts'.format(nm)
        out.append([index[a], index[b]])
    return out


def correct_coordinate_frame(config, points_3d, bodyparts):
    """Rotate / centre one animal's (F, J, 3) points into the frame config['triangulation'] names
    with 'axes' and 'reference_point' (step4:43-88).  Returns (points, M, centre)."""
    index = {bp: i for i, bp in enumerate(bodyparts)}
    axis_of = {'x': 0, 'y': 1, 'z': 2}
    tri = config['triangulation']
    (a_name, a_l, a_r), (b_name, b_l, b_r) = tri['axes'][0], tri['axes'][1]
    a_dir, b_dir = axis_of[a_name], axis_of[b_name]
    c_dir = ({0, 1, 2} - {a_dir, b_dir}).pop()

    def med(pts, j):
        return np.nanmedian(pts[:, j], axis=0)
    a_diff = med(points_3d, index[a_r]) - med(points_3d, index[a_l])
    b_raw = med(points_3d, index[b_r]) - med(points_3d, index[b_l])
    b_diff = b_raw - a_diff * np.dot(b_raw, a_diff) / np.dot(a_diff, a_diff)     # Gram-Schmidt
    M = np.zeros((3, 3))
    M[a_dir], M[b_dir] = a_diff, b_diff
    right_handed = (a_dir, b_dir) in [(0, 1), (2, 0), (1, 2)]
    M[c_dir] = np.cross(a_diff, b_diff) if right_handed else np.cross(b_diff, a_diff)
    M /= np.linalg.norm(M, axis=1)[:, None]
    adj = points_3d.dot(M.T)
    centre = med(adj, index[tri['reference_point']])
    return adj - centre, M, centre


def _reconstruct_optim(cgroup, kp2d_f, config, ransac, joint_len_median, bodyparts, solver):
    """The optim branch of step4:228-291, animal by animal like the reference (the temporal and
    limb-length terms couple all frames of one animal)."""
    n_animal, n_cam, n_frame, n_kp, _ = kp2d_f.shape
    tri = config['triangulation']
    constraints = load_constraints(config, bodyparts)
    constraints_weak = load_constraints(config, bodyparts, 'constraints_weak')
    kp3d = np.zeros((n_animal, n_frame, n_kp, 3))
    E = np.zeros((n_animal, n_frame, n_kp))
    S = np.zeros((n_animal, n_frame, n_kp))
    NC = np.zeros((n_animal, n_frame, n_kp))
    joint_len = []
    for ia in range(n_animal):
        pts = np.ascontiguousarray(kp2d_f[ia, :, :, :, :2])                       # (C, F, J, 2)
        scores = kp2d_f[ia, :, :, :, 2]
        flat = pts.reshape(n_cam, n_frame * n_kp, 2)
        if ransac:
            init = cgroup.triangulate_ransac(flat, outputs="points_3d")[0]        # step4:236-237
        else:
            init = cgroup.triangulate(flat)
        init = init.reshape(n_frame, n_kp, 3)
        if np.isfinite(init[:, :, 0]).sum() < 20:                                # step4:242-245
            p3d = init
        else:
            kw = dict(constraints=constraints, constraints_weak=constraints_weak, scale_smooth=tri['scale_smooth'],
                      scale_length=tri['scale_length'], scale_length_weak=tri['scale_length_weak'],
                      n_deriv_smooth=tri['n_deriv_smooth'], reproj_error_threshold=tri['reproj_error_threshold'])
            kw.update(solver)
            if joint_len_median is None:
                p3d, jl = cgroup.optim_points(pts, init, **kw)                    # step4:247-258
            else:
                p3d, jl = cgroup.optim_points_jointlenfix(pts, init, joint_len_median, **kw)   # step4:259-271
            joint_len.append(jl)
        err = cgroup.reprojection_error(p3d.reshape(-1, 3), flat, mean=True).reshape(n_frame, n_kp)   # step4:278-280
        good = ~np.isnan(pts[..., 0])
        num_cams = good.sum(axis=0).astype('float')
        scores[~good] = 2                                                        # step4:287
        s3 = np.min(scores, axis=0)
        s3[num_cams < 1] = np.nan                                                # step4:290-291
        err = err.copy()
        err[num_cams < 1] = np.nan
        kp3d[ia], E[ia], S[ia], NC[ia] = p3d, err, s3, num_cams
    return {'kp3d': kp3d, 'kp3d_score': S, 'kp3d_err': E, 'num_cams': NC, 'joint_len': joint_len}


def reconstruct(cgroup, kp2d_f, score_threshold=0.5, ransac=False, optim=False, min_cams=3, config=None,
                joint_len_median=None, bodyparts=BODYPARTS, solver=None):
    """kp2d_f: (A, C, F, J, 3) float64 [x, y, score] — the array step4 holds after
    ``kp2d_f.transpose((2,4,0,1,3))`` (step4:190).  Like the reference, keypoints whose
    score is below the threshold are set to NaN IN PLACE and the scores of unused views are
    overwritten with 2 (step4:225-226, 287, 314).

    Returns {'kp3d': (A,F,J,3), 'kp3d_score': (A,F,J), 'kp3d_err': (A,F,J), 'num_cams': (A,F,J)}.
    ``min_cams`` is the value step4 passes to triangulate_ransac (:297).  ``optim`` needs ``config``
    (the dict of config.toml: constraint lists and weights, config_tmpl.toml:66-97); ``joint_len_median``
    switches to ``optim_points_jointlenfix`` (:259-271); the result then also carries 'joint_len'.
    """
    n_animal, n_cam, n_frame, n_kp, _ = kp2d_f.shape
    assert n_cam == len(cgroup.cameras), \
        "kp2d_f has {} cameras, camera group has {}".format(n_cam, len(cgroup.cameras))
    all_points_raw = kp2d_f[..., :2]
    all_scores = kp2d_f[..., 2]
    bad = all_scores < score_threshold                                   # step4:225-226
    all_points_raw[bad] = np.nan
    if optim:
        assert config is not None, "optim=True needs the config dict (constraints and weights)"
        return _reconstruct_optim(cgroup, kp2d_f, config, ransac, joint_len_median, bodyparts, solver or {})

    # (A, C, F, J, 2) -> (C, A*F*J, 2): one launch for every animal
    pts = np.ascontiguousarray(all_points_raw.transpose(1, 0, 2, 3, 4)).reshape(n_cam, -1, 2)
    if ransac:
        # step4:296-300 reads points_2d only for its NaN pattern, which IS picked_vals (a picked camera has a
        # finite raw x, cameras.py:658-659; everything else is NaN, :720): the masked copy of the input stays
        # on the device and 128 of the 168 result bytes per joint-instance never cross PCIe
        p3d, picked, _, errors = cgroup.triangulate_ransac(pts, min_cams=min_cams, outputs="picked")
        good = np.asarray(picked).reshape(n_cam, n_animal, n_frame, n_kp).astype(bool)
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
    ``camera_ids`` is the camera_id list of calib/config.yaml (:107-110, 196-199); ``joint_len``: the
    array of joint_len.npy (:176-181) — its median fixes the limb lengths, as in the reference."""
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
    jl_median = None if joint_len is None else np.median(np.asarray(joint_len), axis=0)     # :179-181
    res = reconstruct(cgroup, kp, score_threshold=tri['score_threshold'], ransac=bool(tri.get('ransac', False)),
                      optim=bool(tri.get('optim', False)), config=config, joint_len_median=jl_median)
    kp3d = res['kp3d']
    if 'reference_point' in tri and 'axes' in tri:                                           # :321-327
        kp3d = np.stack([correct_coordinate_frame(config, kp3d[a], BODYPARTS)[0] for a in range(kp3d.shape[0])])
    data = {'kp3d': kp3d, 'kp3d_score': res['kp3d_score'], 'kp3d_err': res['kp3d_err'],
            'joint_len': res.get('joint_len', [])}
    # the reference writes kp3d_fxdJointLen.pickle when fixed limb lengths were supplied (:334-339)
    fname = 'kp3d.pickle' if joint_len is None else 'kp3d_fxdJointLen.pickle'
    with open(os.path.join(result_dir, fname), 'wb') as f:
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
