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
