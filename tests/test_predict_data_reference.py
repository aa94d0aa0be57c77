"""Keyframe association against the EXECUTED reference: ``tests/golden/predict_data_*.npz`` hold the return values
of the unmodified ``MultiEstimator.predict_data`` (step2_crossviewmatching.py:502-713), run frame by frame by
oracle/make_golden_step2.py with a stand-in for the two ``cv2.omnidir`` functions it reaches (opencv-contrib is not
installed; the stand-in is the oracle's restated Mei model, so the omnidir arithmetic stays unpinned — everything
else of the method, from the identity weighting to get_best_comb's combination order and the bcomb bookkeeping, is
the reference's own execution).

CPU tier: the loop-faithful restatement oracle/crossview.py ``associate_frame`` reproduces the goldens.
GPU tier: ``crossview.associate_batch`` (all keyframes per launch) reproduces them."""
import os

import numpy as np
import pytest

from oracle import crossview as ocv
from oracle import fixtures

NAMES = ["predict_data_dups", "predict_data_clean"]


def _load(name):
    g = dict(np.load(os.path.join(fixtures.GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    return g, fixtures.cams_from_arrays(g)


def _check_frame(g, f, members, p3d, bcomb, tol):
    sel = np.nonzero(g["person_frame"] == f)[0]
    assert len(members) == len(sel), "frame %d: %d persons, reference %d" % (f, len(members), len(sel))
    for k, mem, P, b in zip(sel, members, p3d, bcomb):
        ref_mem = g["person_members"][k]
        assert sorted(np.asarray(mem)[np.asarray(mem) >= 0].tolist()) == sorted(ref_mem[ref_mem >= 0].tolist()), (f, k)
        assert np.array_equal(np.asarray(b), g["bcomb"][k]), (f, k)
        assert np.array_equal(np.isnan(P), np.isnan(g["p3d"][k])), (f, k)
        assert np.nanmax(np.abs(np.asarray(P) - g["p3d"][k]), initial=0.0) <= tol, (f, k)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_associate_frame_equals_executed_predict_data(name):
    g, cams = _load(name)
    for f in range(g["kp_raw"].shape[0]):
        n = int(g["dim"][f, -1])
        m, p, b = ocv.associate_frame(cams, g["kp_raw"][f, :n], g["dim"][f], g["cid"][f, :n], g["bbox"][f, :n])
        _check_frame(g, f, m, p, b, 1e-7)
    assert (g["person_frame"] >= 0).all() and len(g["person_frame"]) >= g["kp_raw"].shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_associate_batch_equals_executed_predict_data(name):
    from macaque_3d_pose_estimation_b200 import crossview as cv
    g, cams = _load(name)
    from tests.test_gpu_parity import group_from_golden
    cg = group_from_golden(g)
    res = cv.associate_batch(cg, g["kp_raw"], g["dim"], g["cid"], g["bbox"])
    for f in range(g["kp_raw"].shape[0]):
        sel = np.nonzero(res["frame"] == f)[0]
        _check_frame(g, f, [res["members"][k] for k in sel], [res["p3d"][k] for k in sel],
                     [res["bcomb"][k] for k in sel], 1e-6)


def _step3_containers(g):
    """The reference's containers rebuilt from the golden arrays: T[i_cam][i_frame] = list of tracks (entry[0] =
    bbox id, entry[5] = (J,3) keypoints), trk[a][i_frame][i_cam] = bbox id or -1."""
    kp, trk = g["kp"], g["trk"]
    A, F, C, J, _ = kp.shape
    T = [[[] for _ in range(F)] for _ in range(C)]
    for f in range(F):
        for c in range(C):
            for a in range(A):
                if trk[a, f, c] >= 0:
                    T[c][f].append([int(trk[a, f, c]), 0, 0, 0, 0, kp[a, f, c].tolist()])
    return T, trk


@pytest.mark.gpu
def test_gpu_step3_traces_equal_executed_reference():
    """step3_crossframematching.py calc_3dpose (:254-272), calc_3dtrace (:274-302), calc_dist_pose (:304-311),
    executed by oracle/make_golden_step2.py (same cv2.omnidir stand-in), against the batched call-site shims."""
    from macaque_3d_pose_estimation_b200 import crossview as cv
    from oracle import camera_math as cm
    g, cams = _load("step3_traces")
    C = g["rig_model"].shape[0]
    camparam = {"camera_id": [str(x) for x in g["rig_names"]],
                "K": [g["rig_K"][i] for i in range(C)], "xi": [np.array([[g["rig_xi"][i]]]) for i in range(C)],
                "D": [g["rig_dist"][i, :4].reshape(1, 4) for i in range(C)],
                "rvecs": [g["rig_rvec"][i].reshape(3, 1) for i in range(C)],
                "tvecs": [g["rig_tvec"][i].reshape(3, 1) for i in range(C)],
                "pmat": [np.hstack([cm.rodrigues(g["rig_rvec"][i]), g["rig_tvec"][i].reshape(3, 1)]) for i in range(C)]}
    T, trk = _step3_containers(g)
    J = g["kp"].shape[3]
    traces = []
    for a in range(2):
        tr = cv.calc_3dtrace_tracklet(trk[a], T, g["frames"], camparam, "", J)
        ref = g["trace%d" % a]
        assert np.array_equal(np.isnan(tr), np.isnan(ref))
        assert np.nanmax(np.abs(tr - ref)) <= 1e-6
        traces.append(tr)
    assert abs(cv.trace_distance(traces[0], traces[1]) - float(g["rmse"])) <= 1e-6
    poses = cv.calc_3dpose_batch(g["kp"][0, :4], camparam, thr_kp=0.3)
    assert np.array_equal(np.isnan(poses), np.isnan(g["poses"]))
    assert np.nanmax(np.abs(poses - g["poses"])) <= 1e-6
    p = cv.calc_p3d(T, trk[0], 3, camparam, n_kp=J)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        assert np.allclose(p, np.nanmean(g["poses"][3], axis=0), rtol=0, atol=1e-6)
