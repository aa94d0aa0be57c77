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


def test_gpu_ls_rank_deficient_matches_pinv():
    """mct.triangulatePoints with two coincident cameras: np.linalg.pinv's minimum-norm solution (golden from
    the executed reference), next to well-posed points."""
    g, cams = fixtures.load_golden("ls_degenerate")
    cp = camparam_from_golden(g)
    p3 = cv.triangulatePoints("", list(g["ls_xy"]), g["ls_use"], True, camparam=cp)
    ref = g["ls_p3d"]
    assert np.isfinite(p3).all()
    assert np.abs(p3[6:] - ref[6:]).max() <= 1e-6
    # the minimum-norm solutions: compare relative to their size (the cut-off acts on a 1e-16-level quantity)
    assert np.abs(p3[:6] - ref[:6]).max() <= 1e-6 * np.abs(ref[:6]).max()


def test_gpu_match_svt_240_reference_frames():
    """matchSVT bit-exact on 240 keyframes executed by the reference (ragged detection counts, identity
    term on, every 4th frame with heavy 2D noise: frames that do not settle into clean blocks)."""
    g, _ = fixtures.load_golden("svt_ragged_f240")
    W, dim, ref = g["W"], g["dim"].astype(np.int32), g["match"]
    got, its = cv.match_svt_batch(W, dim, dim.shape[1] - 1, alpha=0.5, _lambda=50.0, return_iters=True)
    bad = [f for f in range(W.shape[0]) if not np.array_equal(got[f], ref[f])]
    assert not bad, "match matrices differ from the reference in frames %s" % bad[:10]
    assert int(its.max()) > int(np.median(its)) + 5            # the set does contain slow frames


def _association_case(seed, F, model="pinhole", dup=0.08, drop=0.12, noise=0.4):
    from macaque_3d_pose_estimation_b200 import synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup
    C, A, J = 8, 6, 17
    dicts = synth.make_rig(C, model, seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    rng = np.random.default_rng(seed)
    X = synth.make_tracks(F, A, seed=seed) * np.array([0.6, 0.6, 0.5])
    M = C * (A + 2)
    kp = np.zeros((F, M, J, 3))
    dim = np.zeros((F, C + 1), dtype=np.int32)
    cid = -np.ones((F, M), dtype=np.int32)
    bbox = -np.ones((F, M), dtype=np.int64)
    owner = -np.ones((F, M), dtype=np.int64)
    for f in range(F):
        m = 0
        for c in range(C):
            for a in rng.permutation(A):
                if rng.random() < drop:
                    continue
                reps = 2 if rng.random() < dup else 1         # the detector fires twice on one animal
                for r in range(reps):
                    if m >= dim[f, c] + A + 2:
                        break
                    raw = cams[c].project(X[f, a]) + rng.normal(0, noise * (1 + 4 * r), size=(J, 2))
                    sc = rng.uniform(0.3, 1.0, size=J)
                    sc[rng.random(J) < 0.1] = 0.0
                    kp[f, m] = np.concatenate([raw, sc[:, None]], axis=1)
                    cid[f, m] = a if rng.random() < 0.6 else -1
                    bbox[f, m] = 100 * c + m
                    owner[f, m] = a
                    m += 1
            dim[f, c + 1] = m
    return cg, cams, X, kp, dim, cid, bbox, owner


@pytest.mark.parametrize("seed,model,dup,drop,noise,F,on_device",
                         [(51, "pinhole", 0.08, 0.12, 0.4, 24, False), (52, "fisheye", 0.08, 0.12, 0.4, 24, False),
                          (53, "pinhole", 0.004, 0.12, 0.4, 72, False), (54, "pinhole", 0.0, 0.0, 0.1, 40, True),
                          (55, "fisheye", 0.003, 0.05, 0.2, 64, True)])
def test_gpu_associate_batch_equals_per_frame_oracle(seed, model, dup, drop, noise, F, on_device):
    """associate_batch (all keyframes per launch) against the loop-faithful restatement of
    MultiEstimator.predict_data (oracle/crossview.py associate_frame) frame by frame: same persons, same
    members (incl. the duplicate-detection combinations of get_best_comb), same 3D poses."""
    cg, cams, X, kp, dim, cid, bbox, owner = _association_case(seed, F, model, dup=dup, drop=drop, noise=noise)
    if on_device:                              # detections already on the GPU: same result
        import torch
        res = cv.associate_batch(cg, torch.from_numpy(kp).cuda(), dim, cid, bbox)
    else:
        res = cv.associate_batch(cg, kp, dim, cid, bbox)
    assert np.all(np.diff(res["frame"]) >= 0), "persons are not ordered by frame"
    n_dup_frames = 0
    dup_frames = set()
    for f in range(F):
        n = int(dim[f, -1])
        m_ref, p_ref, b_ref = ocv.associate_frame(cams, kp[f, :n], dim[f], cid[f, :n], bbox[f, :n])
        sel = np.nonzero(res["frame"] == f)[0]
        assert len(sel) == len(m_ref), "frame %d: %d persons, reference %d" % (f, len(sel), len(m_ref))
        for k, (mr, pr, br) in zip(sel, zip(m_ref, p_ref, b_ref)):
            mine = res["members"][k]
            assert sorted(mine[mine >= 0].tolist()) == sorted(np.asarray(mr).tolist())
            assert np.array_equal(res["bcomb"][k], br)
            assert np.array_equal(np.isnan(res["p3d"][k]), np.isnan(pr))
            assert np.nanmax(np.abs(res["p3d"][k] - pr), initial=0.0) <= 1e-6
        lab = res["label"][f, :n]
        for c in range(8):
            seg = lab[dim[f, c]:dim[f, c + 1]]
            seg = seg[seg >= 0]
            n_dup_frames += int(len(seg) != len(set(seg.tolist())))
            if len(seg) != len(set(seg.tolist())):
                dup_frames.add(f)
    # frames with a duplicate-detection cluster go through the host's combination scoring, the others
    # through the device member tables: the cases cover all-duplicate, mixed and duplicate-free recordings
    # (a cluster with two detections of one camera also arises without a double detection, when SVT merges
    # two animals)
    if dup >= 0.05:
        assert n_dup_frames > 0, "no duplicate-detection cluster in the test data"
    if drop == 0.0 and dup == 0.0:
        assert len(dup_frames) < F, "no frame takes the device member-table path"
    # the persons are the animals
    good = 0
    for k in range(len(res["frame"])):
        f = int(res["frame"][k])
        mem = res["members"][k]
        own = owner[f, mem[mem >= 0]]
        good += int(len(set(own.tolist())) == 1 and np.nanmax(np.abs(res["p3d"][k] - X[f, own[0]])) < 25.0)
    assert good >= 0.9 * len(res["frame"])


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
