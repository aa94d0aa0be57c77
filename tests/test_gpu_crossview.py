"""GPU tier: cross-view association kernels (ray affinity, matchSVT, inhomogeneous LS
triangulation) through the C-ABI against golden vectors produced by the unmodified
reference (step2_crossviewmatching.geometry_affinity2 / matchSVT, mct.triangulatePoints).

Tolerances: affinity 1e-9 (absolute, values in [0,1]); match matrices bit-exact (uint8);
LS 3D points 1e-6 mm."""
import numpy as np
import pytest

from macaque_3d_pose_estimation_b200 import crossview as cv
from oracle import crossview as ocv
from oracle import fixtures

pytestmark = pytest.mark.gpu
CROSS = fixtures.golden_names("crossview")


def camparam_from_golden(g):
    from oracle import camera_math as cm
    C = g["rig_model"].shape[0]
    cp = {"camera_id": [str(x) for x in g["rig_names"]], "K": [], "xi": [], "D": [], "rvecs": [], "tvecs": [],
          "pmat": []}
    for i in range(C):
        cp["K"].append(g["rig_K"][i])
        cp["xi"].append(np.zeros((1, 1)))
        cp["D"].append(np.zeros((1, 4)))
        cp["rvecs"].append(g["rig_rvec"][i].reshape(3, 1))
        cp["tvecs"].append(g["rig_tvec"][i].reshape(3, 1))
        cp["pmat"].append(np.hstack([cm.rodrigues(g["rig_rvec"][i]), g["rig_tvec"][i].reshape(3, 1)]))
    return cp


@pytest.mark.parametrize("name", CROSS)
def test_gpu_crossview_golden(name):
    g, cams = fixtures.load_golden(name)
    cp = camparam_from_golden(g)
    nf = int(g["n_frames"])
    for f in range(nf):
        kp, dim = g["f%d_kp" % f], g["f%d_dimGroup" % f]
        aff = cv.geometry_affinity2(kp.copy(), dim, "", camparam=cp)
        ref = g["f%d_aff" % f]
        assert aff.shape == ref.shape
        assert np.array_equal(np.isnan(aff), np.isnan(ref))
        assert np.nanmax(np.abs(aff - ref)) <= 1e-9
        match = cv.matchSVT(g["f%d_W" % f].copy(), dim, alpha=0.5, _lambda=50, dual_stochastic_SVT=False)
        assert match.dtype == np.uint8
        assert np.array_equal(match, g["f%d_match" % f])
        p3 = cv.triangulatePoints("", list(g["f%d_ls_xy" % f]), g["f%d_ls_use" % f], True, camparam=cp)
        ref3 = g["f%d_ls_p3d" % f]
        assert np.array_equal(np.isnan(p3), np.isnan(ref3))
        assert np.nanmax(np.abs(p3 - ref3)) <= 1e-6
    # batched, ragged: all frames in one launch, padded to the largest detection count
    M = max(g["f%d_kp" % f].shape[0] for f in range(nf))
    J = g["f0_kp"].shape[1]
    kpb = np.zeros((nf, M, J, 3))
    dimb = np.zeros((nf, len(cams) + 1), dtype=np.int32)
    Wb = np.zeros((nf, M, M))
    for f in range(nf):
        m = g["f%d_kp" % f].shape[0]
        kpb[f, :m] = g["f%d_kp" % f]
        dimb[f] = g["f%d_dimGroup" % f]
        Wb[f, :m, :m] = g["f%d_W" % f]
    cg = cv.group_from_camparam(cp)
    affb, distb = cv.geometry_affinity_batch(cg, kpb, dimb, return_dist=True)
    mb, its = cv.match_svt_batch(Wb, dimb, len(cams), alpha=0.5, _lambda=50.0, return_iters=True)
    for f in range(nf):
        m = g["f%d_kp" % f].shape[0]
        # padding changes the population of the z-score only through entries == 300 (ignored)
        assert np.nanmax(np.abs(affb[f, :m, :m] - g["f%d_aff" % f])) <= 1e-9
        D = ocv.ray_distance_matrix(cams, g["f%d_kp" % f], g["f%d_dimGroup" % f])
        assert np.nanmax(np.abs(distb[f, :m, :m] - D)) <= 1e-8
        assert np.array_equal(mb[f, :m, :m], g["f%d_match" % f])
        assert not mb[f, m:].any() and not mb[f, :, m:].any()
        _, it_ref, _ = ocv.match_svt(g["f%d_W" % f], g["f%d_dimGroup" % f], alpha=0.5, lam=50.0, return_info=True)
        assert abs(int(its[f]) - it_ref) <= 2


def test_gpu_keyframe_association_recovers_animals():
    """MultiEstimator.predict_data (step2:502-713) on a synthetic omnidir rig, 6 animals x 8 views:
    the clusters must be the animals and the 3D poses the ground truth.  (The reference's own
    predict_data needs cv2.omnidir and cannot run here — a functional check, not a parity one.)"""
    from macaque_3d_pose_estimation_b200 import synth
    from oracle import cameragroup as og
    C, A, J = 8, 6, 17
    dicts = synth.make_rig(C, "omnidir", seed=41)
    cams = fixtures.cams_from_dicts(dicts)
    cp = {"camera_id": [d["name"] for d in dicts], "K": [np.array(d["K"]) for d in dicts],
          "xi": [np.array(d["xi"]).reshape(1, 1) for d in dicts], "D": [np.array(d["D"]).reshape(1, 4) for d in dicts],
          "rvecs": [np.array(d["rotation"]).reshape(3, 1) for d in dicts],
          "tvecs": [np.array(d["translation"]).reshape(3, 1) for d in dicts], "pmat": None}
    rng = np.random.default_rng(41)
    X = synth.make_tracks(1, A, seed=41)[0] * np.array([0.6, 0.6, 0.5])          # keep every view in frame
    info = {}
    for c in range(C):
        dets = []
        for a in rng.permutation(A):
            raw = cams[c].project(X[a]) + rng.normal(0, 0.3, size=(J, 2))
            sc = rng.uniform(0.4, 1.0, size=J)
            und = cams[c].undistort(raw)
            dets.append({"pose2d": und, "pose2d_raw": np.concatenate([raw, sc[:, None]], axis=1), "bbox": [0, 0, 1, 1],
                         "bbox_id": (c, int(a)), "cid": -1})
        info[c] = [dets]
    est = cv.MultiEstimator(cfg=None)
    matched, P3d, bcomb = est.predict_data(info, camparam=cp)
    assert len(matched) == A
    seen = set()
    for idxs, p3, bc in zip(matched, P3d, bcomb):
        assert len(idxs) == C and (bc >= 0).all()
        assert len(set(bc.tolist())) == 1                        # one animal per cluster
        a = int(bc[0])
        seen.add(a)
        assert np.abs(p3 - X[a]).max() < 2.0                      # mm, 0.3 px detector noise
    assert seen == set(range(A))


def test_gpu_step3_batch_triangulation():
    """calc_3dpose_batch / calc_3dtrace (step3_crossframematching.py:254-302) against the oracle's
    per-frame mct.triangulatePoints restatement, omnidir rig, score gate 0.3."""
    from macaque_3d_pose_estimation_b200 import synth
    from oracle import cameragroup as og
    C, J, F = 8, 17, 60
    dicts = synth.make_rig(C, "omnidir", seed=43)
    cams = fixtures.cams_from_dicts(dicts)
    cp = {"camera_id": [d["name"] for d in dicts], "K": [np.array(d["K"]) for d in dicts],
          "xi": [np.array(d["xi"]).reshape(1, 1) for d in dicts], "D": [np.array(d["D"]).reshape(1, 4) for d in dicts],
          "rvecs": [np.array(d["rotation"]).reshape(3, 1) for d in dicts],
          "tvecs": [np.array(d["translation"]).reshape(3, 1) for d in dicts], "pmat": None}
    rng = np.random.default_rng(43)
    X = synth.make_tracks(F, 1, seed=43)[:, 0] * np.array([0.6, 0.6, 0.5])
    kp = np.empty((F, C, J, 3))
    for c in range(C):
        kp[:, c, :, :2] = cams[c].project(X.reshape(-1, 3)).reshape(F, J, 2) + rng.normal(0, 0.3, size=(F, J, 2))
    kp[..., 2] = rng.uniform(0.1, 1.0, size=(F, C, J))
    kp[rng.random((F, C)) < 0.3] = np.nan                         # camera does not see the animal
    got = cv.calc_3dpose_batch(kp, cp, thr_kp=0.3)
    for f in range(0, F, 7):
        und = np.stack([cams[c].undistort(kp[f, c, :, :2]) for c in range(C)])
        with np.errstate(invalid="ignore"):
            use = ~(np.isnan(kp[f, :, :, 0]) | (kp[f, :, :, 2] < 0.3))
        ref = ocv.triangulate_ls(cams, np.nan_to_num(und), use.T)
        assert np.array_equal(np.isnan(got[f]), np.isnan(ref))
        assert np.nanmax(np.abs(got[f] - ref), initial=0.0) <= 1e-6
    trace = cv.calc_3dtrace(kp, cp)
    assert trace.shape == (F, 3)
    seen = (~np.isnan(kp[:, :, :, 0]).all(axis=2)).sum(axis=1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        ref_trace = np.nanmedian(np.where((seen < 2)[:, None, None], np.nan, got), axis=1)
    assert np.array_equal(trace, ref_trace, equal_nan=True)
    ok = ~np.isnan(trace[:, 0])
    assert ok.sum() > F // 2 and np.abs(trace[ok] - np.median(X[ok], axis=1)).max() < 100.0
