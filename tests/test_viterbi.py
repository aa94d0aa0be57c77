"""2D keypoint Viterbi filter (anipose/filter_pose.py:26-120, 151-186, 332-343).

CPU tier: oracle/viterbi.py against golden vectors produced by executing the reference
(oracle/make_golden.py, cases viterbi_*): the filter's outputs are SELECTED input detections
(x, y, score * 2^-age), so the comparison is bit-exact.
GPU tier: k_viterbi through the C ABI (m3d_viterbi_filter) against the goldens and the oracle.  The
kernel evaluates the transition log-probabilities with CUDA's erfc / log1p / exp (a few ulp from
scipy's); a decision can differ only where two path scores agree to ~1e-14 relative, which the
seeded cases below do not contain (asserted: zero differing frames)."""
import glob
import os

import numpy as np
import pytest

from oracle import viterbi as ov

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "viterbi_*.npz")))


def _cfg(g):
    return {"filter": {"score_threshold": float(g["score_threshold"]), "n_back": int(g["n_back"]),
                       "offset_threshold": float(g["offset_threshold"])}}


def _series(F, J, P, seed):
    from macaque_3d_pose_estimation_b200 import synth
    return synth.make_detection_series(F, J, P, seed)


def test_goldens_present():
    assert len(NAMES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    inp = g["all_points"].copy()
    pts, sc = ov.filter_pose_viterbi(_cfg(g), inp)
    assert np.array_equal(pts, g["points"], equal_nan=True)
    assert np.array_equal(sc, g["scores"], equal_nan=True)
    assert np.array_equal(ov.wrap_points(pts, sc), g["wrapped"], equal_nan=True)
    # like the reference, the threshold is written into the caller's array
    low = g["all_points"][..., 2] < float(g["score_threshold"])
    assert np.isnan(inp[..., 0][low]).all() and not np.isnan(inp[..., 0][~low]).any()


def test_oracle_remove_dups_and_missing():
    pts = np.array([[[10.0, 10.0], [12.0, 13.0], [100.0, 100.0]],      # second within 5 px of the first
                    [[np.nan, np.nan], [np.nan, np.nan], [50.0, 50.0]]])
    out = ov.remove_dups(pts, 5)
    assert np.isnan(out[0, 1, 0]) and out[0, 0, 0] == 10.0 and out[0, 2, 0] == 100.0 and out[1, 2, 0] == 50.0
    # a series without any valid detection is the missing-point particle everywhere
    p, s = ov.viterbi_path(np.full((7, 1, 2), np.nan), np.zeros((7, 1)), 3, 25)
    assert (p == -1).all() and (s == 0.001).all()


# ---------------------------------------------------------------------------------------------
# GPU tier
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference_golden(name):
    from macaque_3d_pose_estimation_b200 import filter2d
    g = np.load(os.path.join(GOLD, name + ".npz"))
    pts, sc = filter2d.filter_pose_viterbi(_cfg(g), g["all_points"].copy())
    assert pts.shape == g["points"].shape and sc.shape == g["scores"].shape
    assert np.array_equal(pts, g["points"], equal_nan=True)
    assert np.array_equal(sc, g["scores"], equal_nan=True)
    assert np.array_equal(filter2d.wrap_points(pts, sc), g["wrapped"], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("F,J,P,n_back,scale", [(1500, 40, 1, 3, 25.0), (400, 24, 2, 3, 25.0), (300, 16, 3, 2, 12.0),
                                               (1, 5, 1, 3, 25.0), (2, 5, 2, 3, 25.0), (130, 9, 1, 5, 30.0),
                                               # k_viterbi_small with several candidates per frame (dedup inside)
                                               (300, 8, 2, 2, 25.0), (200, 6, 4, 1, 20.0), (97, 3, 3, 1, 25.0)])
def test_gpu_random_vs_oracle(F, J, P, n_back, scale):
    from macaque_3d_pose_estimation_b200 import filter2d
    allp = _series(F, J, P, 77 + F + P)
    cfg = {"filter": {"score_threshold": 0.3, "n_back": n_back, "offset_threshold": scale}}
    rp, rs, rc = ov.filter_pose_viterbi(cfg, allp.copy(), return_choice=True)
    cand = np.ascontiguousarray(allp.transpose(1, 0, 2, 3))
    pts, sc, ch = filter2d.viterbi_series(cand, n_back, scale, 0.3, return_choice=True)
    assert int((ch.T != rc).sum()) == 0
    assert np.array_equal(pts.transpose(1, 0, 2), rp, equal_nan=True)
    assert np.array_equal(sc.T, rs, equal_nan=True)


@pytest.mark.gpu
def test_gpu_all_missing_and_limits():
    from macaque_3d_pose_estimation_b200 import filter2d
    cand = np.zeros((3, 20, 1, 3))
    cand[..., 2] = 0.1                                   # every score below the threshold
    pts, sc, ch = filter2d.viterbi_series(cand, 3, 25.0, 0.3, return_choice=True)
    assert (pts == -1).all() and (sc == 0.001).all() and (ch == -1).all()
    pts, sc = filter2d.viterbi_series(np.zeros((0, 10, 1, 3)), 3, 25.0, 0.3)
    assert pts.shape == (0, 10, 2)
    with pytest.raises(RuntimeError, match="n_back"):
        filter2d.viterbi_series(np.zeros((1, 4, 9, 3)), 4, 25.0, 0.3)


@pytest.mark.gpu
def test_gpu_step4_filter_stage_layout():
    """filter_stage == the loop of step4_aniposefiltering.py:144-167 run with the oracle."""
    from macaque_3d_pose_estimation_b200 import filter2d
    A, F, C, J = 2, 120, 3, 5
    kp2d = np.stack([np.stack([_series(F, J, 1, 500 + 10 * a + c)[:, :, 0] for c in range(C)], axis=1)
                     for a in range(A)])                  # (A, F, C, J, 3)
    got = filter2d.filter_stage(kp2d)
    cfg = {"filter": dict(filter2d.STEP4_FILTER)}
    kp = kp2d.transpose((1, 3, 0, 4, 2))                  # step4:156
    ref = np.zeros(kp.shape)
    for a in range(A):
        for c in range(C):
            points = np.expand_dims(kp[:, :, a, :, c], 2).copy()
            pf, sf = ov.filter_pose_viterbi(cfg, points)
            ref[:, :, a, :, c] = np.squeeze(ov.wrap_points(pf, sf))
    assert got.shape == ref.shape == (F, J, A, 3, C)
    assert np.array_equal(got, ref, equal_nan=True)


@pytest.mark.gpu
def test_gpu_step4_files_end_to_end(tmp_path):
    """pipeline3d.run_step4: kp2d.pickle -> kp2d_f.pickle -> kp3d.pickle (step4:140-339)."""
    import pickle
    from macaque_3d_pose_estimation_b200 import filter2d, pipeline3d, synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup
    from oracle import cameragroup as og
    from oracle import fixtures
    seed, A, F, J, C = 21, 2, 60, 17, 8
    dicts = synth.make_rig(C, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    CameraGroup.from_dicts(dicts).dump(str(tmp_path / "calibration.toml"))
    X = synth.make_tracks(F, A, seed=seed)                                       # (F, A, J, 3)
    p2 = synth.corrupt(og.project(cams, X.reshape(-1, 3)), seed=seed).reshape(C, F, A, J, 2)
    sc = np.random.default_rng(seed).uniform(0.2, 1.0, size=(C, F, A, J))
    kp2d = np.concatenate([p2, sc[..., None]], axis=-1).transpose(2, 1, 0, 3, 4)   # (A, F, C, J, 3)
    with open(tmp_path / "kp2d.pickle", "wb") as f:
        pickle.dump(kp2d, f)
    cfg = {"triangulation": {"score_threshold": 0.5, "ransac": False, "optim": False}}
    data = pipeline3d.run_step4(str(tmp_path), list(range(1, C + 1)), config=cfg)
    with open(tmp_path / "kp2d_f.pickle", "rb") as f:
        kp2d_f = pickle.load(f)
    assert kp2d_f.shape == (F, J, A, 3, C)
    assert np.array_equal(kp2d_f, filter2d.filter_stage(kp2d), equal_nan=True)
    # 3D stage on the filtered detections == oracle triangulation of the same detections
    kf = kp2d_f.transpose((2, 4, 0, 1, 3))                                       # (A, C, F, J, 3)
    for a in range(A):
        pts = kf[a, :, :, :, :2].copy()
        pts[kf[a, :, :, :, 2] < 0.5] = np.nan
        ref = og.triangulate(cams, pts.reshape(C, F * J, 2)).reshape(F, J, 3)
        assert np.array_equal(np.isnan(data["kp3d"][a]), np.isnan(ref))
        assert np.nanmax(np.abs(data["kp3d"][a] - ref)) <= 1e-6
