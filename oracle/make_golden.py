"""Generate tests/golden/*.npz by EXECUTING the unmodified reference.

Run in the build container only (needs /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden [--only NAME]

The reference ships no golden vectors; these files pin both the
oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_parity.py)
to what the reference's own code returns:
  src/third_party/aniposelib/cameras.py  Camera/FisheyeCamera/CameraGroup
  src/pipeline/step2_crossviewmatching.py  geometry_affinity2, matchSVT
  src/utils/multicam_toolbox.py  triangulatePoints
  src/third_party/anipose/filter_pose.py  viterbi_path, wrap_points
Each .npz stores the rig (camera dict fields as arrays), the inputs and every
output, so that nothing under /root/reference is needed at test time.
"""
import argparse
import os
import sys
import time
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from macaque_3d_pose_estimation_b200 import synth  # noqa: E402


def _import_reference():
    sys.dont_write_bytecode = True
    for name in ("h5py", "imgstore", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    from src.third_party.aniposelib import cameras as ref_cameras
    # numba JIT warm-up so that ref_seconds excludes compilation
    cg = ref_cameras.CameraGroup.from_dicts(synth.make_rig(3, "pinhole", seed=1))
    X = synth.make_tracks(1, 1, seed=1).reshape(-1, 3)
    p2 = cg.project(X)
    cg.reprojection_error(cg.triangulate(p2), p2, mean=True)
    cg.triangulate_ransac(p2[:, :2])
    return ref_cameras


def rig_arrays(cams):
    """Flatten camera dicts into arrays that np.savez can hold."""
    C = len(cams)
    model = np.zeros(C, dtype=np.int32)
    K = np.zeros((C, 3, 3))
    dist = np.zeros((C, 14))
    ndist = np.zeros(C, dtype=np.int32)
    rvec = np.zeros((C, 3))
    tvec = np.zeros((C, 3))
    xi = np.zeros(C)
    for i, d in enumerate(cams):
        if d.get("fisheye"):
            model[i] = 1
        elif d.get("omnidir"):
            model[i] = 2
        K[i] = np.array(d["K"] if model[i] == 2 else d["matrix"])
        dd = np.array(d["D"] if model[i] == 2 else d["distortions"], dtype=np.float64).ravel()
        dist[i, :dd.size] = dd
        ndist[i] = dd.size
        rvec[i] = d["rotation"]
        tvec[i] = d["translation"]
        xi[i] = d.get("xi", [0.0])[0]
    return dict(rig_model=model, rig_K=K, rig_dist=dist, rig_ndist=ndist,
                rig_rvec=rvec, rig_tvec=tvec, rig_xi=xi,
                rig_names=np.array([d["name"] for d in cams]))


def observations(cg, n_frames, n_animals, seed, **corrupt_kw):
    X = synth.make_tracks(n_frames, n_animals, seed=seed).reshape(-1, 3)
    clean = cg.project(X)
    return X, synth.corrupt(clean, seed=seed, **corrupt_kw)


def case_dlt(ref, name, n_cams, model, n_frames, n_animals, seed, **kw):
    cams = synth.make_rig(n_cams, model, seed=seed)
    cg = ref.CameraGroup.from_dicts(cams)
    X, p2d = observations(cg, n_frames, n_animals, seed, **kw)
    und = np.stack([cam.undistort_points(np.copy(p2d[c])) for c, cam in enumerate(cg.cameras)])
    t0 = time.time()
    p3d = cg.triangulate(p2d)
    t_tri = time.time() - t0
    p3d_noundist = cg.triangulate(und, undistort=False)
    err_full = cg.reprojection_error(p3d, p2d, mean=False)
    t0 = time.time()
    err_mean = cg.reprojection_error(p3d, p2d, mean=True)
    t_err = time.time() - t0
    proj = cg.project(X)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rig_arrays(cams), X_true=X, p2d=p2d,
                        undistorted=und, p3d=p3d, p3d_noundist=p3d_noundist, err_full=err_full,
                        err_mean=err_mean, proj_true=proj,
                        ref_seconds=np.array([t_tri, t_err]))
    print(name, "N=%d tri %.2fs err %.2fs" % (p2d.shape[1], t_tri, t_err), flush=True)


def case_ransac(ref, name, n_cams, model, n_frames, n_animals, seed, min_cams=2, extra=None, **kw):
    cams = synth.make_rig(n_cams, model, seed=seed)
    cg = ref.CameraGroup.from_dicts(cams)
    X, p2d = observations(cg, n_frames, n_animals, seed, **kw)
    if extra is not None:
        p2d = extra(p2d, cg, X)
    # instrument: count subset evaluations per call
    calls = {"n": 0}
    orig = ref.CameraGroup.triangulate

    def counting(self, *a, **k):
        calls["n"] += 1
        return orig(self, *a, **k)
    ref.CameraGroup.triangulate = counting
    try:
        t0 = time.time()
        out, picked, pts2d, errs = cg.triangulate_ransac(np.copy(p2d), min_cams=min_cams)
        dt = time.time() - t0
    finally:
        ref.CameraGroup.triangulate = orig
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rig_arrays(cams), X_true=X, p2d=p2d,
                        min_cams=np.array(min_cams), p3d=out, picked=picked, points_2d=pts2d,
                        errors=errs, n_subsets_evaluated=np.array(calls["n"]),
                        ref_seconds=np.array([dt]))
    print(name, "N=%d ransac %.1fs  %.1f subsets/pt" % (p2d.shape[1], dt, calls["n"] / p2d.shape[1]),
          flush=True)


def edge_points(p2d, cg, X):
    """Hand-made edge cases of SURVEY.md §8c written over the first points."""
    p = p2d
    C = p.shape[0]
    rng = np.random.default_rng(7)
    p[:, 0] = np.nan                                  # all-NaN point
    p[:, 1] = np.nan
    p[2, 1] = cg.cameras[2].project(X[1]).ravel()     # exactly one valid view
    p[:, 2] = np.nan                                  # two valid views (min_cams=3 run too)
    p[[1, 4], 2] = cg.project(X[2:3])[[1, 4], 0]
    p[3, 3, 1] = np.nan                               # NaN only in y of one view
    p[:, 4] = rng.uniform([0, 0], [2048, 1536], size=(C, 2))   # all-garbage point
    p[:, 5] = cg.project(X[5:6])[:, 0]                # single gross outlier at camera 0
    p[0, 5] += np.array([80.0, -45.0])
    p[:, 6] = cg.project(X[6:7])[:, 0]                # exact, noise-free point
    p[:, 7] = cg.project(X[7:8])[:, 0]                # outlier at the last camera
    p[C - 1, 7] += np.array([-60.0, 30.0])
    p[:, 8] = np.nan                                  # two views, one of them with y = NaN
    p[[0, 5], 8] = cg.project(X[8:9])[[0, 5], 0]
    p[5, 8, 1] = np.nan
    p[:, 9] = cg.project(X[9:10])[:, 0]               # x = NaN but y finite (view dropped)
    p[2, 9, 0] = np.nan
    return p


def case_crossview(ref, name, n_frames, seed, drop=0.0):
    from src.pipeline import step2_crossviewmatching as s2
    from src.utils import multicam_toolbox as mct
    import cv2
    C, A, J = 8, 6, 17
    cams = synth.make_rig(C, "pinhole", seed=seed)
    rng = np.random.default_rng(seed + 5)
    camparam = {"camera_id": [d["name"] for d in cams], "K": [], "xi": [], "D": [],
                "rvecs": [], "tvecs": [], "pmat": []}
    for d in cams:
        R, _ = cv2.Rodrigues(np.array(d["rotation"]))
        t = np.array(d["translation"]).reshape(3, 1)
        camparam["K"].append(np.array(d["matrix"]))
        camparam["xi"].append(np.zeros((1, 1)))
        camparam["D"].append(np.zeros((1, 4)))
        camparam["rvecs"].append(np.array(d["rotation"]).reshape(3, 1))
        camparam["tvecs"].append(t)
        camparam["pmat"].append(np.hstack([R, t]))
    X = synth.make_tracks(n_frames, A, seed=seed)               # (F, A, J, 3)
    frames = []
    for f in range(n_frames):
        kps, dim, owner = [], [0], []
        for c in range(C):
            P = camparam["pmat"][c]
            for a in range(A):
                if rng.random() < drop:
                    continue
                Xc = X[f, a] @ P[:, :3].T + P[:, 3]
                xy = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, 4e-4, size=(J, 2))
                sc = rng.uniform(0.3, 1.0, size=J)
                sc[rng.random(J) < 0.1] = 0.0
                kps.append(np.concatenate([xy, sc[:, None]], axis=1))
                owner.append(a)
            dim.append(len(kps))
        kp = np.array(kps)
        dimGroup = np.array(dim)
        t0 = time.time()
        aff = s2.geometry_affinity2(kp.copy(), dimGroup, "", camparam=camparam)
        t_aff = time.time() - t0
        W = 0.8 * aff
        W *= (aff > 0)
        W = np.nan_to_num(W)
        t0 = time.time()
        match = s2.matchSVT(W.copy(), dimGroup, alpha=0.5, _lambda=50, dual_stochastic_SVT=False)
        t_svt = time.time() - t0
        # mct.triangulatePoints on animal 0 of this frame (views that saw it)
        M = kp.shape[0]
        kp2d = np.full((C, J, 3), np.nan)
        kp2d[:, :, 2] = 0.0
        for i in range(M):
            c = int(np.searchsorted(dimGroup, i, side="right") - 1)
            if owner[i] == 0:
                kp2d[c] = kp[i]
        frame_use = (~np.isnan(kp2d[:, :, 0]) & (kp2d[:, :, 2] >= 0.1)).T
        und = [np.nan_to_num(kp2d[c, :, :2]) for c in range(C)]
        p3d_ls = mct.triangulatePoints("", und, frame_use, True, camparam=camparam)
        frames.append(dict(kp=kp, dimGroup=dimGroup, owner=np.array(owner), aff=aff, W=W,
                           match=match, ls_xy=np.array(und), ls_use=frame_use, ls_p3d=p3d_ls,
                           t=np.array([t_aff, t_svt])))
        print(name, "frame", f, "M=%d aff %.2fs svt %.2fs" % (M, t_aff, t_svt), flush=True)
    arrs = rig_arrays(cams)
    for f, fr in enumerate(frames):
        for k, v in fr.items():
            arrs["f%d_%s" % (f, k)] = v
    arrs["n_frames"] = np.array(n_frames)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)


def case_svt(ref, name, n_frames, seed, drop=0.15, p_cid=0.7, p_cid_wrong=0.1, noise=4e-4):
    """matchSVT (step2_crossviewmatching.py:130-216) alone on many keyframes: ragged detection counts,
    the identity term of predict_data switched on (W = 0.2 cid + 0.8 aff, :557-575, with some wrong and
    some missing identities), 2D noise levels from clean to heavy so that part of the frames do NOT
    converge to a clean block structure.  The affinity comes from the numpy oracle (pinned against the
    reference by the crossview_* goldens); the match matrices are the executed reference's."""
    from src.pipeline import step2_crossviewmatching as s2
    from oracle import crossview as ocv
    from oracle import fixtures
    C, A, J = 8, 6, 17
    cams = synth.make_rig(C, "pinhole", seed=seed)
    specs = fixtures.cams_from_dicts(cams)
    rng = np.random.default_rng(seed + 5)
    from oracle import camera_math as cm
    P = [np.hstack([cm.rodrigues(np.array(d["rotation"])), np.array(d["translation"]).reshape(3, 1)]) for d in cams]
    X = synth.make_tracks(n_frames, A, seed=seed)
    Ws, dims, matches, iters = [], [], [], []
    t0 = time.time()
    for f in range(n_frames):
        nz = noise * (1.0 if f % 4 else 6.0) * (1.0 + 3.0 * rng.random())     # every 4th frame is much noisier
        kps, dim, owner = [], [0], []
        for c in range(C):
            for a in range(A):
                if rng.random() < drop:
                    continue
                Xc = X[f, a] @ P[c][:, :3].T + P[c][:, 3]
                xy = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, nz, size=(J, 2))
                sc = rng.uniform(0.3, 1.0, size=J)
                sc[rng.random(J) < 0.1] = 0.0
                kps.append(np.concatenate([xy, sc[:, None]], axis=1))
                owner.append(a)
            dim.append(len(kps))
        kp, dimGroup, owner = np.array(kps), np.array(dim), np.array(owner)
        M = kp.shape[0]
        aff = ocv.geometry_affinity(specs, kp, dimGroup)
        cid = np.where(rng.random(M) < p_cid, owner, -1)
        wrong = rng.random(M) < p_cid_wrong
        cid[wrong] = rng.integers(0, A, size=int(wrong.sum()))
        sub2cam = np.searchsorted(dimGroup, np.arange(M), side="right") - 1
        cid_mat = ((sub2cam[:, None] != sub2cam[None, :]) & (cid[:, None] >= 0) & (cid[:, None] == cid[None, :])).astype(float)
        W = 0.2 * cid_mat + 0.8 * aff
        W *= (aff > 0)
        W = np.nan_to_num(W)
        match = s2.matchSVT(W.copy(), dimGroup, alpha=0.5, _lambda=50, dual_stochastic_SVT=False)
        Wp = np.zeros((A * C, A * C))
        Wp[:M, :M] = W
        mp = np.zeros((A * C, A * C), dtype=np.uint8)
        mp[:M, :M] = match
        Ws.append(Wp)
        dims.append(dimGroup)
        matches.append(mp)
        if f % 20 == 0:
            print(name, "frame", f, "M=%d  %.0f s" % (M, time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rig_arrays(cams), W=np.array(Ws), dim=np.array(dims),
                        match=np.array(matches), n_frames=np.array(n_frames))


def case_ls_degenerate(ref, name, seed):
    """mct.triangulatePoints (multicam_toolbox.py:433-486) where the rays do NOT determine the point: two
    coincident cameras see it (rank-2 system -> np.linalg.pinv returns the minimum-norm solution), next to
    ordinary well-posed points."""
    from src.utils import multicam_toolbox as mct
    import cv2
    cams = synth.make_rig(4, "pinhole", seed=seed)
    cams[1] = dict(cams[0], name="2")                                   # camera 2 coincides with camera 1
    camparam = {"camera_id": [d["name"] for d in cams], "K": [], "xi": [], "D": [], "rvecs": [], "tvecs": [], "pmat": []}
    for d in cams:
        R, _ = cv2.Rodrigues(np.array(d["rotation"]))
        t = np.array(d["translation"]).reshape(3, 1)
        camparam["K"].append(np.array(d["matrix"]))
        camparam["xi"].append(np.zeros((1, 1)))
        camparam["D"].append(np.zeros((1, 4)))
        camparam["rvecs"].append(np.array(d["rotation"]).reshape(3, 1))
        camparam["tvecs"].append(t)
        camparam["pmat"].append(np.hstack([R, t]))
    rng = np.random.default_rng(seed)
    n = 12
    X = rng.uniform([-500, -500, 0], [500, 500, 900], size=(n, 3))
    und = []
    for P in camparam["pmat"]:
        Xc = X @ P[:, :3].T + P[:, 3]
        und.append(Xc[:, :2] / Xc[:, 2:3])
    use = np.ones((n, 4), dtype=bool)
    use[:6, 2:] = False                                                 # points 0..5: only the coincident pair
    p3d = mct.triangulatePoints("", und, use, True, camparam=camparam)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rig_arrays(cams), ls_xy=np.array(und), ls_use=use,
                        ls_p3d=p3d, X_true=X)
    print(name, "degenerate rows", p3d[:2], "regular", np.abs(p3d[6:] - X[6:]).max())


def case_possible(ref, name, n_cams, n_frames, seed, n_possible=2, min_cams=2, p_swap=0.3, p_missing=0.2):
    """CameraGroup.triangulate_possible with P candidates per camera (cameras.py:639-724): the
    second candidate is a distractor (N(0, 40 px) away) or missing; in 30 % of the (camera, point)
    cells the true detection is the SECOND candidate."""
    cams = synth.make_rig(n_cams, "pinhole", seed=seed)
    cg = ref.CameraGroup.from_dicts(cams)
    rng = np.random.default_rng(seed + 5)
    X = synth.make_tracks(n_frames, 1, seed=seed).reshape(-1, 3)
    clean = cg.project(X).reshape(n_cams, -1, 2)
    N = clean.shape[1]
    pts = np.full((n_cams, N, n_possible, 2), np.nan)
    good = clean + rng.normal(0, 0.3, size=clean.shape)
    for p in range(n_possible):
        pts[:, :, p] = clean + rng.normal(0, 40.0, size=clean.shape)
    first = rng.random((n_cams, N)) >= p_swap
    pts[:, :, 0][first] = good[first]
    pts[:, :, 1][~first] = good[~first]
    miss = rng.random((n_cams, N, n_possible)) < p_missing
    pts[miss] = np.nan
    pts[:, 0] = np.nan                                   # a point without any candidate
    pts[1:, 1] = np.nan                                  # a point seen by one camera only
    t0 = time.time()
    out, picked, p2d, err = cg.triangulate_possible(pts.copy(), min_cams=min_cams)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(OUT, name + ".npz"), points=pts, out=out, picked=picked, points_2d=p2d,
                        errors=err, min_cams=min_cams, ref_seconds=dt, **rig_arrays(cams))
    print("%-28s C=%d N=%d P=%d  reference %.2f s" % (name, n_cams, N, n_possible, dt))


viterbi_series = synth.make_detection_series


def case_viterbi(name, n_frames, n_joints, n_possible, seed, score_threshold=0.3, n_back=3, offset_threshold=25):
    """anipose/filter_pose.py viterbi_path executed per series exactly as filter_pose_viterbi
    :151-186 does (score threshold written into the input, one series per joint; the spawn Pool of
    the reference only distributes these independent calls)."""
    import src.third_party.aniposelib as al
    import src.third_party.aniposelib.boards  # noqa: F401
    sys.modules.setdefault("aniposelib", al)
    sys.modules.setdefault("aniposelib.boards", al.boards)
    from src.third_party.anipose import filter_pose as fp
    all_points = viterbi_series(n_frames, n_joints, n_possible, seed)
    inp = all_points.copy()
    points_full = all_points[:, :, :, :2]
    scores_full = all_points[:, :, :, 2]
    points_full[scores_full < score_threshold] = np.nan
    F, J = n_frames, n_joints
    points = np.full((F, J, 2), np.nan)
    scores = np.empty((F, J))
    t0 = time.time()
    for j in range(J):
        points[:, j], scores[:, j] = fp.viterbi_path(points_full[:, j, :], scores_full[:, j], n_back, offset_threshold)
    dt = time.time() - t0
    wrapped = fp.wrap_points(points, scores)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), all_points=inp, points=points, scores=scores,
                        wrapped=wrapped, score_threshold=score_threshold, n_back=n_back,
                        offset_threshold=offset_threshold, ref_seconds=dt)
    print("%-28s F=%d J=%d P=%d  reference %.2f s" % (name, F, J, n_possible, dt))


MACAQUE_CONSTRAINTS = [(0, 1), (0, 2), (1, 2), (0, 3), (0, 4), (1, 3), (2, 4), (3, 4), (5, 3), (6, 4), (5, 6), (5, 7),
                       (7, 9), (6, 8), (8, 10), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16)]
MACAQUE_CONSTRAINTS_WEAK = [(5, 11), (6, 12), (5, 12), (6, 11), (5, 6), (11, 12), (1, 0), (2, 0), (1, 3), (2, 4), (3, 4)]


def case_optim(ref, name, n_cams, n_frames, seed, n_deriv=2, reproj_loss="soft_l1", fix=False, with_scores=False):
    """optim_points / optim_points_jointlenfix (cameras.py:1116-1270) on one synthetic animal with the
    constraint lists and weights of configs/config_tmpl.toml:66-97: the reference's start vector, its
    residual vector there and at a perturbed point, and what least_squares(ftol=1e-3) returns."""
    cams = synth.make_rig(n_cams, "pinhole", seed=seed)
    cg = ref.CameraGroup.from_dicts(cams)
    rng = np.random.default_rng(seed)
    X = synth.make_tracks(n_frames, 1, seed=seed)[:, 0]                       # (F, J, 3)
    # a smooth track: the random walk of make_tracks + a slow limb motion
    X = X + 30.0 * np.sin(np.arange(n_frames)[:, None, None] / 7.0 + rng.uniform(0, 6, (1, X.shape[1], 3)))
    p2 = synth.corrupt(cg.project(X.reshape(-1, 3)), seed=seed, noise=0.8, p_outlier=0.03, sigma_outlier=25.0,
                       p_missing=0.15)
    pts = p2.reshape(n_cams, n_frames, X.shape[1], 2)
    pts[0, 3, 2, 1] = np.nan                                                # y missing only: one residual dropped
    p3d0 = cg.triangulate(pts.reshape(n_cams, -1, 2)).reshape(n_frames, -1, 3)
    p3d0[5, 4] = np.nan                                                     # a gap the interpolation fills
    p3d0[9:12, 7] = np.nan
    scores = rng.uniform(0.5, 1.0, size=pts.shape[:3]) if with_scores else None
    kw = dict(scale_smooth=3, scale_length=5, scale_length_weak=2, n_deriv_smooth=n_deriv, reproj_error_threshold=3,
              reproj_loss=reproj_loss)
    cons, consw = np.array(MACAQUE_CONSTRAINTS), np.array(MACAQUE_CONSTRAINTS_WEAK)
    intp = np.apply_along_axis(ref.interpolate_data, 0, p3d0)
    med = np.apply_along_axis(ref.medfilt_data, 0, intp, size=7)
    s_full = kw["scale_smooth"] * (1.0 / np.mean(np.abs(np.diff(med, axis=0))))
    x0 = cg._initialize_params_triangulation(intp, cons, consw)
    x0[~np.isfinite(x0)] = 0
    args = (pts, cons, consw, scores, s_full, kw["scale_length"], kw["scale_length_weak"],
            kw["reproj_error_threshold"], reproj_loss, n_deriv)
    r0 = np.array(cg._error_fun_triangulation(x0, *args))
    x1 = x0 + rng.normal(0, 2.0, x0.shape)
    r1 = np.array(cg._error_fun_triangulation(x1, *args))
    out = dict(rig_arrays(cams), points=pts, p3d0=p3d0, intp=intp, scale_smooth_full=s_full, x0=x0, r0=r0, x1=x1,
               r1=r1, constraints=cons, constraints_weak=consw, n_deriv=n_deriv, fix=int(fix),
               reproj_loss=np.array(reproj_loss), **{k: v for k, v in kw.items() if k != "reproj_loss"})
    if scores is not None:
        out["scores"] = scores
    t0 = time.time()
    if fix:
        jl = x0[intp.size:]
        new, jl_out = cg.optim_points_jointlenfix(pts, p3d0, jl, constraints=cons, constraints_weak=consw, scores=scores, **kw)
        rf = np.array(cg._error_fun_triangulation_jointlenfix(new.ravel(), pts, jl, cons, consw, scores, s_full,
                                                               kw["scale_length"], kw["scale_length_weak"],
                                                               kw["reproj_error_threshold"], reproj_loss, n_deriv))
        out["r0_fix"] = np.array(cg._error_fun_triangulation_jointlenfix(
            x0[:intp.size], pts, jl, cons, consw, scores, s_full, kw["scale_length"], kw["scale_length_weak"],
            kw["reproj_error_threshold"], reproj_loss, n_deriv))
    else:
        new, jl_out = cg.optim_points(pts, p3d0, constraints=cons, constraints_weak=consw, scores=scores, **kw)
        rf = np.array(cg._error_fun_triangulation(np.hstack([new.ravel(), jl_out]), *args))
    out["ref_seconds"] = time.time() - t0
    out["opt_p3d"] = new
    out["opt_joint_len"] = jl_out
    out["opt_cost"] = 0.5 * float(rf @ rf)
    out["x0_cost"] = 0.5 * float(r0 @ r0)
    out["X_true"] = X
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s F=%d  cost %.6g -> %.6g  reference %.1f s" % (name, n_frames, out["x0_cost"], out["opt_cost"],
                                                              out["ref_seconds"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    ref = _import_reference()
    S = 20261018
    cases = {
        "dlt_pinhole_c8": lambda: case_dlt(ref, "dlt_pinhole_c8", 8, "pinhole", 60, 2, S + 1, p_missing=0.1),
        "dlt_pinhole8_c8": lambda: case_dlt(ref, "dlt_pinhole8_c8", 8, "pinhole8", 20, 2, S + 2, p_missing=0.1),
        "dlt_fisheye_c8": lambda: case_dlt(ref, "dlt_fisheye_c8", 8, "fisheye", 30, 2, S + 3, p_missing=0.1),
        "dlt_pinhole_c2": lambda: case_dlt(ref, "dlt_pinhole_c2", 2, "pinhole", 15, 2, S + 4, p_missing=0.05),
        "dlt_pinhole_c3": lambda: case_dlt(ref, "dlt_pinhole_c3", 3, "pinhole", 15, 2, S + 5, p_missing=0.1),
        "dlt_pinhole_c16": lambda: case_dlt(ref, "dlt_pinhole_c16", 16, "pinhole", 15, 2, S + 6, p_missing=0.1),
        "ransac_pinhole_c8": lambda: case_ransac(ref, "ransac_pinhole_c8", 8, "pinhole", 40, 2, S + 11,
                                                  p_outlier=0.2, p_missing=0.1),
        "ransac_pinhole_c8_min3": lambda: case_ransac(ref, "ransac_pinhole_c8_min3", 8, "pinhole", 12, 2, S + 12,
                                                       min_cams=3, p_outlier=0.2, p_missing=0.1),
        "ransac_fisheye_c8": lambda: case_ransac(ref, "ransac_fisheye_c8", 8, "fisheye", 10, 2, S + 13,
                                                  p_outlier=0.2, p_missing=0.1),
        "ransac_pinhole_c3": lambda: case_ransac(ref, "ransac_pinhole_c3", 3, "pinhole", 12, 2, S + 14,
                                                  p_outlier=0.2, p_missing=0.1),
        "ransac_pinhole_c12": lambda: case_ransac(ref, "ransac_pinhole_c12", 12, "pinhole", 1, 2, S + 15,
                                                   p_outlier=0.08, p_missing=0.1),
        "ransac_edges_c8": lambda: case_ransac(ref, "ransac_edges_c8", 8, "pinhole", 1, 1, S + 16,
                                                extra=edge_points),
        "ransac_edges_c8_min3": lambda: case_ransac(ref, "ransac_edges_c8_min3", 8, "pinhole", 1, 1, S + 16,
                                                     min_cams=3, extra=edge_points),
        "ransac_noisy_c8": lambda: case_ransac(ref, "ransac_noisy_c8", 8, "pinhole", 6, 2, S + 17,
                                                noise=0.55, p_outlier=0.1, p_missing=0.1),
        "crossview_m48": lambda: case_crossview(ref, "crossview_m48", 4, S + 21),
        "crossview_ragged": lambda: case_crossview(ref, "crossview_ragged", 3, S + 22, drop=0.15),
        "possible_c4_p2": lambda: case_possible(ref, "possible_c4_p2", 4, 3, S + 41),
        "possible_c5_p2_min3": lambda: case_possible(ref, "possible_c5_p2_min3", 5, 2, S + 42, min_cams=3),
        "possible_c3_p3": lambda: case_possible(ref, "possible_c3_p3", 3, 3, S + 43, n_possible=3),
        "viterbi_p1": lambda: case_viterbi("viterbi_p1", 400, 6, 1, S + 31),
        "viterbi_p2": lambda: case_viterbi("viterbi_p2", 150, 4, 2, S + 32),
        "viterbi_p1_nb4": lambda: case_viterbi("viterbi_p1_nb4", 120, 3, 1, S + 33, n_back=4, offset_threshold=10),
        "ls_degenerate": lambda: case_ls_degenerate(ref, "ls_degenerate", S + 71),
        "svt_ragged_f240": lambda: case_svt(ref, "svt_ragged_f240", 240, S + 61),
        "optim_c8_n2": lambda: case_optim(ref, "optim_c8_n2", 8, 48, S + 51),
        "optim_c4_n1_huber": lambda: case_optim(ref, "optim_c4_n1_huber", 4, 30, S + 52, n_deriv=1, reproj_loss="huber",
                                                with_scores=True),
        "optim_c8_fix": lambda: case_optim(ref, "optim_c8_fix", 8, 36, S + 53, fix=True),
    }
    for name, fn in cases.items():
        if args.only and args.only != name:
            continue
        fn()


if __name__ == "__main__":
    main()
