"""GPU tier: the CUDA path, called through the C-ABI (libm3d.so) via the drop-in Python
classes, against (1) the golden vectors produced by the unmodified reference, (2) the
oracle on seeded random inputs, (3) size-independent properties at larger N.

Tolerances (BASELINE.json north_star): selected camera subsets and inlier masks bit-exact;
3D points within 1e-4 relative or 0.01 mm (asserted: 1e-5 mm — two-camera rigs with near-antiparallel rays put LAPACK's own
noise at ~3e-6 mm); reprojection errors within 1e-3 px (asserted: 1e-7 px)."""
import numpy as np
import pytest

from macaque_3d_pose_estimation_b200 import _lib, synth
from macaque_3d_pose_estimation_b200.cameras import Camera, CameraGroup, FisheyeCamera, OmnidirCamera
from oracle import cameragroup as og
from oracle import fixtures

pytestmark = pytest.mark.gpu

DLT = fixtures.golden_names("dlt")
RANSAC = fixtures.golden_names("ransac")
P3D_TOL_MM = 1e-5
ERR_TOL_PX = 1e-7


def group_from_golden(g):
    cams = []
    for i in range(g["rig_model"].shape[0]):
        n = int(g["rig_ndist"][i])
        kw = dict(size=(2048, 1536), rvec=g["rig_rvec"][i], tvec=g["rig_tvec"][i], name=str(g["rig_names"][i]))
        m = int(g["rig_model"][i])
        if m == 0:
            cams.append(Camera(matrix=g["rig_K"][i], dist=g["rig_dist"][i, :n], **kw))
        elif m == 1:
            cams.append(FisheyeCamera(matrix=g["rig_K"][i], dist=g["rig_dist"][i, :n], **kw))
        else:
            cams.append(OmnidirCamera(K=g["rig_K"][i], D=g["rig_dist"][i, :n], xi=[g["rig_xi"][i]], **kw))
    return CameraGroup(cams)


def _eq_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()
    assert _lib.load().m3d_device_count() > 0


@pytest.mark.parametrize("name", DLT)
def test_gpu_dlt_golden(name):
    import torch
    g, _ = fixtures.load_golden(name)
    cg = group_from_golden(g)
    before = _lib.load().m3d_launch_count()
    # per-camera Camera.undistort_points and the batched group form
    und = np.stack([cam.undistort_points(np.copy(g["p2d"][c])) for c, cam in enumerate(cg.cameras)])
    assert _eq_nan(und, g["undistorted"])
    assert np.nanmax(np.abs(und - g["undistorted"])) <= 1e-12
    und2 = cg.undistort_points(g["p2d"])
    assert np.array_equal(und, und2, equal_nan=True)
    # triangulate (host-buffer pipeline) and the device-resident torch path
    p3d = cg.triangulate(g["p2d"])
    assert p3d.shape == g["p3d"].shape and p3d.dtype == np.float64
    assert _eq_nan(p3d, g["p3d"])
    assert np.nanmax(np.abs(p3d - g["p3d"])) <= P3D_TOL_MM
    p3d_t = cg.triangulate(torch.from_numpy(g["p2d"]).cuda())
    assert p3d_t.is_cuda and np.array_equal(p3d_t.cpu().numpy(), p3d, equal_nan=True)   # same kernel
    p3d_nu = cg.triangulate(g["undistorted"], undistort=False)
    assert np.nanmax(np.abs(p3d_nu - g["p3d_noundist"])) <= P3D_TOL_MM
    # reprojection error, both forms, evaluated at the REFERENCE's 3D points
    ef = cg.reprojection_error(g["p3d"], g["p2d"])
    assert ef.shape == g["err_full"].shape and _eq_nan(ef, g["err_full"])
    assert np.nanmax(np.abs(ef - g["err_full"])) <= 1e-9
    em = cg.reprojection_error(g["p3d"], g["p2d"], mean=True)
    assert em.shape == g["err_mean"].shape and _eq_nan(em, g["err_mean"])
    assert np.nanmax(np.abs(em - g["err_mean"])) <= 1e-9
    # fused kernel: error at OUR 3D points
    p3f, emf = cg.triangulate_with_error(g["p2d"])
    # (a different template instantiation: FMA contraction may differ in the last bit)
    assert _eq_nan(p3f, p3d) and np.nanmax(np.abs(p3f - p3d)) <= 1e-9
    assert _eq_nan(emf, g["err_mean"]) and np.nanmax(np.abs(emf - g["err_mean"])) <= ERR_TOL_PX
    # projection
    proj = cg.project(g["X_true"])
    assert proj.shape == g["proj_true"].shape
    assert np.abs(proj - g["proj_true"]).max() <= 1e-9
    pc = cg.cameras[0].project(g["X_true"])
    assert pc.shape == (g["X_true"].shape[0], 1, 2)
    assert np.abs(pc.reshape(-1, 2) - g["proj_true"][0]).max() <= 1e-9
    # one-point overloads (cameras.py:603-606, 753-757)
    i = int(np.nonzero(np.isfinite(g["p3d"][:, 0]))[0][0])
    one = cg.triangulate(g["p2d"][:, i])
    assert one.shape == (3,) and np.abs(one - g["p3d"][i]).max() <= P3D_TOL_MM
    e1 = cg.reprojection_error(g["p3d"][i], g["p2d"][:, i], mean=True)
    assert isinstance(e1, float) and abs(e1 - g["err_mean"][i]) <= 1e-9
    e2 = cg.reprojection_error(g["p3d"][i], g["p2d"][:, i])
    assert e2.shape == (len(cg.cameras), 2)
    assert np.abs(cg.get_extrinsics_mats() - np.array([c.extrinsics() for c in fixtures.cams_from_arrays(g)])).max() == 0
    assert _lib.load().m3d_launch_count() > before          # the CUDA kernels really ran


@pytest.mark.parametrize("name", DLT)
def test_gpu_dlt_golden_fast_undistort(name):
    """Opt-in M3D_UNDISTORT_FAST (first three undistortion iterations in float32) against the reference
    goldens at the tolerances BASELINE.json states: 3D points 1e-4 relative or 0.01 mm, errors 1e-3 px."""
    import torch
    g, _ = fixtures.load_golden(name)
    cg = group_from_golden(g)
    cg.fast_undistort = True
    for pts in (g["p2d"], torch.from_numpy(g["p2d"]).cuda()):
        p3d, err = cg.triangulate_with_error(pts)
        if not isinstance(p3d, np.ndarray):
            p3d, err = p3d.cpu().numpy(), err.cpu().numpy()
        assert _eq_nan(p3d, g["p3d"]) and _eq_nan(err, g["err_mean"])
        d = np.abs(p3d - g["p3d"])
        tol = np.maximum(0.01, 1e-4 * np.abs(g["p3d"]))
        assert np.nanmax(d - tol, initial=-1.0) <= 0, "3D points outside the north-star tolerance"
        assert np.nanmax(np.abs(err - g["err_mean"]), initial=0.0) <= 1e-3
        if all(type(c) is Camera and len(c.dist.ravel()) <= 5 for c in cg.cameras):
            # the fast path really ran (it differs from the strict bits) and stays far inside the budget
            strict = group_from_golden(g).triangulate(g["p2d"])
            assert np.nanmax(d, initial=0.0) <= 1e-3
            assert not np.array_equal(strict, p3d, equal_nan=True)
    p1 = cg.triangulate(g["p2d"])
    assert np.nanmax(np.abs(p1 - g["p3d"]) - np.maximum(0.01, 1e-4 * np.abs(g["p3d"])), initial=-1.0) <= 0


def test_gpu_fast_undistort_corner_cases():
    """icdist < 0 bail-out (k1 = -0.9 at the image corners), NaN views, points far outside the image:
    the fast path must take the same decisions as the reference's float64 iteration."""
    dicts = synth.make_rig(8, "pinhole", seed=19)
    for d in dicts[:4]:
        d["distortions"] = [-0.9, 0.0, 0.0, 0.0, 0.0]
    cams = fixtures.cams_from_dicts(dicts)
    rng = np.random.default_rng(19)
    n = 4000
    p2d = np.stack([rng.uniform([-300, -300], [2348, 1836], size=(n, 2)) for _ in range(8)])
    p2d[rng.random((8, n)) < 0.1] = np.nan
    ref_u = og.undistort_points(cams, p2d)
    ref3 = og.triangulate(cams, p2d)
    refe = og.reprojection_error(cams, ref3, p2d, mean=True)
    cg = CameraGroup.from_dicts(dicts)
    cg.fast_undistort = True
    p3d, err = cg.triangulate_with_error(p2d)
    assert _eq_nan(p3d, ref3)
    # garbage-in points (random pixels in 8 views) are ill-conditioned: compare where the reference's own
    # answer is stable, i.e. through the reprojection error it implies
    ok = np.isfinite(refe) & np.isfinite(err)
    bail = (np.abs(ref_u[..., 0] - (p2d[..., 0] - np.array([c.K[0, 2] for c in cams])[:, None]) /
                   np.array([c.K[0, 0] for c in cams])[:, None]) == 0).any(axis=0)
    assert bail.any(), "no icdist < 0 bail-out in the test data"
    assert np.abs(err - refe)[ok].max() <= 1e-3 * np.maximum(1.0, refe[ok]).max()


@pytest.mark.parametrize("name", RANSAC)
def test_gpu_ransac_golden(name):
    import torch
    g, _ = fixtures.load_golden(name)
    cg = group_from_golden(g)
    mc = int(g["min_cams"])
    out, picked, p2d, err, sidx, nev = cg.triangulate_ransac(np.copy(g["p2d"]), min_cams=mc, return_stats=True)
    assert picked.dtype == np.bool_ and picked.shape == g["picked"].shape
    assert np.array_equal(picked, g["picked"])                     # bit-exact inlier masks
    assert np.array_equal(p2d, g["points_2d"], equal_nan=True)
    assert _eq_nan(out, g["p3d"])
    assert np.nanmax(np.abs(out - g["p3d"]), initial=0.0) <= P3D_TOL_MM
    assert np.abs(err - g["errors"]).max() <= ERR_TOL_PX
    assert int(nev.sum()) == int(g["n_subsets_evaluated"])         # identical search length
    # device-resident path returns the same bits
    t = cg.triangulate_ransac(torch.from_numpy(g["p2d"]).cuda(), min_cams=mc, return_stats=True)
    assert np.array_equal(t[1].cpu().numpy(), picked)
    assert np.array_equal(t[0].cpu().numpy(), out, equal_nan=True)
    assert np.array_equal(t[3].cpu().numpy(), err)
    assert np.array_equal(t[4].cpu().numpy(), sidx)
    # triangulate_possible with P = 1 is the same search (cameras.py:738-743)
    tp = cg.triangulate_possible(g["p2d"].reshape(g["p2d"].shape[0], -1, 1, 2), min_cams=mc)
    assert np.array_equal(tp[1], picked) and np.array_equal(tp[0], out, equal_nan=True)


@pytest.mark.parametrize("n_cams,model,kw", [
    (8, "pinhole", dict(p_outlier=0.2, p_missing=0.1)),
    (8, "pinhole", dict(noise=2.0, p_outlier=0.4, p_missing=0.1)),
    (8, "pinhole8", dict(p_outlier=0.2, p_missing=0.2)),
    (8, "fisheye", dict(p_outlier=0.2, p_missing=0.1)),
    (8, "omnidir", dict(p_outlier=0.2, p_missing=0.1)),
    (5, "pinhole", dict(p_outlier=0.3, p_missing=0.1)),
    (2, "pinhole", dict(p_outlier=0.1, p_missing=0.1)),
    (11, "pinhole", dict(p_outlier=0.1, p_missing=0.1)),
    (16, "pinhole", dict(p_outlier=0.05, p_missing=0.3)),      # all three table levels of k_ransac_search16
])
def test_gpu_ransac_random_vs_oracle(n_cams, model, kw):
    seed = 777 + n_cams
    dicts = synth.make_rig(n_cams, model, seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    n_frames = 2 if n_cams >= 16 else (8 if n_cams > 8 else 60)
    X = synth.make_tracks(n_frames, 4, seed=seed).reshape(-1, 3)
    p2 = synth.corrupt(og.project(cams, X), seed=seed, **kw)
    for mc in ((2,) if n_cams >= 16 else (2, 3)):               # the oracle needs ~20 s per 100 16-camera points
        o = og.triangulate_ransac(cams, p2, min_cams=mc, return_stats=True)
        h = cg.triangulate_ransac(p2, min_cams=mc, return_stats=True)
        assert np.array_equal(o[1], h[1])
        assert np.array_equal(o[4], h[4])
        assert np.array_equal(o[5], h[5])
        assert np.array_equal(o[2], h[2], equal_nan=True)
        assert np.nanmax(np.abs(o[0] - h[0]), initial=0.0) <= P3D_TOL_MM
        assert np.abs(o[3] - h[3]).max() <= ERR_TOL_PX


def test_gpu_empty_and_ragged_inputs():
    cg = CameraGroup.from_dicts(synth.make_rig(8, "pinhole", seed=3))
    z = np.zeros((8, 0, 2))
    assert cg.triangulate(z).shape == (0, 3)
    r = cg.triangulate_ransac(z)
    assert r[0].shape == (0, 3) and r[1].shape == (8, 0, 1) and r[2].shape == (8, 0, 2) and r[3].shape == (0,)
    assert cg.reprojection_error(np.zeros((0, 3)), z, mean=True).shape == (0,)
    assert cg.project(np.zeros((0, 3))).shape == (8, 0, 2)
    # all-NaN input: nothing triangulated, errors default to 0.0 for RANSAC, NaN for the mean
    nan = np.full((8, 37, 2), np.nan)
    assert np.isnan(cg.triangulate(nan)).all()
    out, picked, p2, err = cg.triangulate_ransac(nan)
    assert np.isnan(out).all() and not picked.any() and np.isnan(p2).all() and (err == 0.0).all()
    # N not a multiple of the warp / block size, and a 1-camera group
    cams = fixtures.cams_from_dicts(synth.make_rig(8, "pinhole", seed=3))
    X = synth.make_tracks(3, 1, seed=3).reshape(-1, 3)[:33]
    p2 = og.project(cams, X)
    assert np.nanmax(np.abs(cg.triangulate(p2) - og.triangulate(cams, p2))) <= P3D_TOL_MM
    one = cg.subset_cameras([2])
    assert np.isnan(one.triangulate(p2[2:3])).all()


def test_gpu_large_n_properties():
    """BASELINE config-2 shaped run (8 views, ~1e6 joint-instances on the GPU) checked
    through size-independent properties: agreement with the oracle on a random sample,
    chunk invariance of the host pipeline, permutation equivariance, and device == host."""
    import torch
    seed = 99
    dicts = synth.make_rig(8, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    X = synth.make_tracks(15000, 4, seed=seed).reshape(-1, 3)          # 1.02e6 points
    clean = cg.project(X)
    p2 = synth.corrupt(clean, seed=seed, p_outlier=0.2, p_missing=0.1)
    n = p2.shape[1]
    p3d, err = cg.triangulate_with_error(p2)
    rng = np.random.default_rng(1)
    idx = rng.choice(n, 4000, replace=False)
    ref = og.triangulate(cams, p2[:, idx])
    assert _eq_nan(p3d[idx], ref) and np.nanmax(np.abs(p3d[idx] - ref)) <= P3D_TOL_MM
    referr = og.reprojection_error(cams, ref, p2[:, idx], mean=True)
    # outlier-laden DLT points are ill-conditioned: errors of 1e2..1e3 px, compare relatively
    assert np.nanmax(np.abs(err[idx] - referr) / (1.0 + np.abs(referr))) <= 1e-6
    # permutation equivariance + device path == host pipeline (bitwise)
    perm = rng.permutation(n)
    t = torch.from_numpy(np.ascontiguousarray(p2[:, perm])).cuda()
    p3p, errp = cg.triangulate_with_error(t)
    assert np.array_equal(p3p.cpu().numpy(), p3d[perm], equal_nan=True)
    assert np.array_equal(errp.cpu().numpy(), err[perm], equal_nan=True)
    # RANSAC: sample vs oracle, and equivariance
    out, picked, _, rerr, sidx, nev = cg.triangulate_ransac(p2, return_stats=True)
    o = og.triangulate_ransac(cams, p2[:, idx[:2000]], return_stats=True)
    assert np.array_equal(picked[:, idx[:2000]], o[1])
    assert np.array_equal(sidx[idx[:2000]], o[4]) and np.array_equal(nev[idx[:2000]], o[5])
    assert np.nanmax(np.abs(out[idx[:2000]] - o[0])) <= P3D_TOL_MM
    tr = cg.triangulate_ransac(t, return_stats=True)
    assert np.array_equal(tr[1].cpu().numpy(), picked[:, perm])
    assert np.array_equal(tr[4].cpu().numpy(), sidx[perm]) and np.array_equal(tr[5].cpu().numpy(), nev[perm])
    # (which of the two search kernels finishes a point depends on its neighbours in the sorted
    # order, and their Gram sums associate differently: last-bit differences in p3d are allowed)
    assert _eq_nan(tr[0].cpu().numpy(), out[perm])
    assert np.nanmax(np.abs(tr[0].cpu().numpy() - out[perm])) <= 1e-9
    # selected points: reprojection error of the selection is what was reported
    sel = sidx >= 0
    chk = cg.reprojection_error(out, np.where(picked, p2, np.nan), mean=True)
    assert np.nanmax(np.abs(chk[sel] - rerr[sel])) <= 1e-9
    assert 20.0 < nev.mean() < 120.0                                  # ~1 + p (2^k - 1) subsets / point


def test_gpu_step4_stage_matches_oracle():
    """The 3D stage of step4_aniposefiltering.py:219-330 (plain and RANSAC branches)."""
    from macaque_3d_pose_estimation_b200 import pipeline3d
    seed = 5
    dicts = synth.make_rig(8, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    A, F, J = 2, 40, 17
    X = synth.make_tracks(F, A, seed=seed)                              # (F, A, J, 3)
    p2 = synth.corrupt(og.project(cams, X.reshape(-1, 3)), seed=seed, p_outlier=0.1)
    p2 = p2.reshape(8, F, A, J, 2).transpose(2, 0, 1, 3, 4)            # (A, C, F, J, 2)
    rng = np.random.default_rng(seed)
    scores = rng.uniform(0.2, 1.0, size=(A, 8, F, J))
    kp2d = np.concatenate([p2, scores[..., None]], axis=-1)
    for ransac in (False, True):
        res = pipeline3d.reconstruct(cg, kp2d.copy(), score_threshold=0.5, ransac=ransac)
        for a in range(A):
            pts = kp2d[a, :, :, :, :2].copy()
            sc = kp2d[a, :, :, :, 2].copy()
            pts[sc < 0.5] = np.nan
            flat = pts.reshape(8, F * J, 2)
            if ransac:
                o3, opick, o2d, oerr = og.triangulate_ransac(cams, flat, min_cams=3)
                ncam = opick.sum(axis=0).sum(axis=1).reshape(F, J).astype(float)
                good = ~np.isnan(o2d.reshape(8, F, J, 2)[..., 0])
            else:
                o3 = og.triangulate(cams, flat)
                oerr = og.reprojection_error(cams, o3, flat, mean=True)
                good = ~np.isnan(pts[..., 0])
                ncam = good.sum(axis=0).astype(float)
            sc[~good] = 2
            s3 = sc.min(axis=0)
            s3[ncam < 2] = np.nan
            e3 = oerr.reshape(F, J).copy()
            e3[ncam < 2] = np.nan
            assert _eq_nan(res["kp3d"][a], o3.reshape(F, J, 3))
            assert np.nanmax(np.abs(res["kp3d"][a] - o3.reshape(F, J, 3))) <= P3D_TOL_MM
            assert np.array_equal(res["kp3d_score"][a], s3, equal_nan=True)
            assert _eq_nan(res["kp3d_err"][a], e3)
            assert np.nanmax(np.abs(res["kp3d_err"][a] - e3)) <= ERR_TOL_PX


def test_gpu_distort_points_and_camera_models():
    """Camera.distort_points (cameras.py:301,366,487) = projectPoints of (x, y, 1) with identity
    extrinsics, for the three camera models, against the oracle."""
    from oracle import camera_math as cm
    rng = np.random.default_rng(5)
    xy = rng.uniform(-0.6, 0.6, size=(500, 2))
    xyz = np.concatenate([xy, np.ones((500, 1))], axis=1)
    for model in ("pinhole", "pinhole8", "fisheye", "omnidir"):
        d = synth.make_rig(1, model, seed=21)[0]
        cg = CameraGroup.from_dicts([d])
        cam = cg.cameras[0]
        spec = fixtures.cams_from_dicts([d])[0]
        z3 = np.zeros(3)
        if spec.model == cm.MODEL_PINHOLE:
            ref = cm.project_pinhole(xyz, z3, z3, spec.K, spec.dist)
        elif spec.model == cm.MODEL_FISHEYE:
            ref = cm.project_fisheye(xyz, z3, z3, spec.K, spec.dist)
        else:
            ref = cm.project_omnidir(xyz, z3, z3, spec.K, spec.xi, spec.dist)
        out = cam.distort_points(xy)
        assert out.shape == xy.shape and np.abs(out - ref).max() <= 1e-9
        # undistort(distort(x)) comes back to x up to the model's fixed iteration count
        back = cam.undistort_points(out)
        tol = 5e-3 if model.startswith("pinhole") else 1e-7
        assert np.abs(back - xy).max() <= tol
        # per-camera reprojection_error = p2d - project(p3d) (cameras.py:325-327)
        X = synth.make_tracks(5, 2, seed=3).reshape(-1, 3)
        p2 = spec.project(X) + 1.0
        assert np.abs(cam.reprojection_error(X, p2) - 1.0).max() <= 1e-9


@pytest.mark.slow
def test_gpu_baseline_full_size_properties():
    """BASELINE.json config 2 at FULL size (8 views, 4 x 17 x 1e6 = 6.8e7 joint-instances, device
    resident) and config 3 shape at 1.36e7, checked through size-independent properties: a
    random sample against the oracle, agreement of disjoint launches (slice == whole), and the
    RANSAC selection re-scored by an independent kernel."""
    import torch
    from bench import make_device_workload
    seed = 20261018 + 2
    dicts = synth.make_rig(8, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    dev = torch.device("cuda", torch.cuda.current_device())
    xy = make_device_workload(cg, 1000000, 4, 17, seed, "dlt", dev)
    n = xy.shape[1]
    assert n == 68000000
    p3d, err = cg.triangulate_with_error(xy)
    rng = np.random.default_rng(3)
    idx = np.sort(rng.choice(n, 5000, replace=False))
    tidx = torch.from_numpy(idx).to(dev)
    sample = xy[:, tidx].cpu().numpy()
    ref = og.triangulate(cams, sample)
    got = p3d[tidx].cpu().numpy()
    assert _eq_nan(got, ref) and np.nanmax(np.abs(got - ref)) <= P3D_TOL_MM
    referr = og.reprojection_error(cams, ref, sample, mean=True)
    # (two-view points with near-parallel rays: LAPACK's own noise reaches ~1e-6 mm / px)
    assert np.nanmax(np.abs(err[tidx].cpu().numpy() - referr) / (1.0 + np.abs(referr))) <= 1e-5
    # a slice launched on its own gives the same bits as the whole
    lo, hi = 31234567, 31234567 + 100003
    p3s, errs = cg.triangulate_with_error(xy[:, lo:hi].contiguous())
    assert torch.equal(p3s.nan_to_num(), p3d[lo:hi].nan_to_num())
    assert torch.equal(errs.nan_to_num(), err[lo:hi].nan_to_num())
    # NaN pattern of the output == fewer than two valid views in the input
    nvalid = (~torch.isnan(xy[:, :, 0])).sum(dim=0)
    assert torch.equal(torch.isnan(p3d[:, 0]), nvalid < 2)
    del xy, p3d, err, nvalid
    torch.cuda.empty_cache()

    xy = make_device_workload(cg, 200000, 4, 17, seed, "ransac", dev)
    n = xy.shape[1]
    out, picked, xyp, rerr, sidx, nev = cg.triangulate_ransac(xy, return_stats=True)
    idx = np.sort(rng.choice(n, 3000, replace=False))
    tidx = torch.from_numpy(idx).to(dev)
    o = og.triangulate_ransac(cams, xy[:, tidx].cpu().numpy(), return_stats=True)
    assert np.array_equal(picked[:, tidx].cpu().numpy(), o[1])
    assert np.array_equal(sidx[tidx].cpu().numpy(), o[4]) and np.array_equal(nev[tidx].cpu().numpy(), o[5])
    assert np.nanmax(np.abs(out[tidx].cpu().numpy() - o[0])) <= P3D_TOL_MM
    assert (np.abs(rerr[tidx].cpu().numpy() - o[3]) / (1.0 + np.abs(o[3]))).max() <= 1e-5
    # the reported error is the mean reprojection error of the selected views at the selected point
    chk = cg.reprojection_error(out, xyp, mean=True)
    sel = sidx >= 0
    assert float(((chk[sel] - rerr[sel]).abs() / (1.0 + rerr[sel])).max()) <= 1e-9
    assert torch.equal(torch.isnan(xyp[:, :, 0]), ~picked[:, :, 0])
    # selected => error below the acceptance bound; early exit => below the 0.5 px threshold or arg-min
    assert float(rerr[sel].max()) < 200.0


def test_gpu_step4_file_stage(tmp_path):
    """pipeline3d.run_stage: kp2d_f.pickle + calibration.toml + config -> kp3d.pickle with the
    reference's layout (step4_aniposefiltering.py:172-339), checked against the oracle."""
    import pickle
    from macaque_3d_pose_estimation_b200 import pipeline3d
    seed = 9
    dicts = synth.make_rig(8, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    CameraGroup.from_dicts(dicts).dump(str(tmp_path / "calibration.toml"))
    A, F, J = 2, 25, 17
    X = synth.make_tracks(F, A, seed=seed)
    p2 = synth.corrupt(og.project(cams, X.reshape(-1, 3)), seed=seed).reshape(8, F, A, J, 2)
    sc = np.random.default_rng(seed).uniform(0.3, 1.0, size=(8, F, A, J))
    ids = [8, 7, 6, 5, 4, 3, 2, 1]                      # camera order of calib/config.yaml, not of the toml
    order = [i - 1 for i in ids]                        # the data's camera axis follows that order (step4:196-213)
    kp2d_f = np.concatenate([p2[order], sc[order][..., None]], axis=-1).transpose(1, 3, 2, 4, 0)   # (F, J, A, 3, C)
    with open(tmp_path / "kp2d_f.pickle", "wb") as f:
        pickle.dump(kp2d_f, f)
    cfg = {"triangulation": {"score_threshold": 0.5, "ransac": False, "optim": False}}
    data = pipeline3d.run_stage(str(tmp_path), ids, config=cfg)
    with open(tmp_path / "kp3d.pickle", "rb") as f:
        disk = pickle.load(f)
    assert set(disk) == {"kp3d", "kp3d_score", "kp3d_err", "joint_len"}
    assert disk["kp3d"].shape == (A, F, J, 3) and disk["kp3d_score"].shape == (A, F, J)
    for a in range(A):
        pts = p2[order][:, :, a].copy()
        pts[sc[order][:, :, a] < 0.5] = np.nan
        ref = og.triangulate([cams[i] for i in order], pts.reshape(8, F * J, 2)).reshape(F, J, 3)
        assert _eq_nan(disk["kp3d"][a], ref) and np.nanmax(np.abs(disk["kp3d"][a] - ref)) <= P3D_TOL_MM
    assert np.array_equal(data["kp3d"], disk["kp3d"], equal_nan=True)


POSSIBLE = fixtures.golden_names("possible")


@pytest.mark.parametrize("name", POSSIBLE)
def test_gpu_triangulate_possible_golden(name):
    """CameraGroup.triangulate_possible with P > 1 candidates per camera (cameras.py:639-724)
    against golden vectors produced by executing the reference."""
    g, cams = fixtures.load_golden(name)
    cg = group_from_golden(g)
    mc = int(g["min_cams"])
    out, picked, p2d, err, idx, nev = cg.triangulate_possible(g["points"], min_cams=mc, return_stats=True)
    assert picked.shape == g["picked"].shape and picked.dtype == np.bool_
    assert np.array_equal(picked, g["picked"])
    assert np.array_equal(p2d, g["points_2d"], equal_nan=True)
    assert _eq_nan(out, g["out"]) and np.nanmax(np.abs(out - g["out"]), initial=0.0) <= P3D_TOL_MM
    assert np.abs(err - g["errors"]).max() <= ERR_TOL_PX
    o = og.triangulate_possible(cams, g["points"], min_cams=mc, return_stats=True)
    assert np.array_equal(idx, o[4]) and np.array_equal(nev, o[5])


@pytest.mark.parametrize("C,P,mc", [(8, 2, 2), (6, 3, 3), (16, 2, 2), (8, 4, 2)])
def test_gpu_triangulate_possible_random_vs_oracle(C, P, mc):
    seed = 4100 + 10 * C + P
    dicts = synth.make_rig(C, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    rng = np.random.default_rng(seed)
    X = synth.make_tracks(1, 1, seed=seed).reshape(-1, 3) * np.array([0.6, 0.6, 0.5])
    X = X[:12 if C * P <= 18 else 6]
    clean = og.project(cams, X)
    N = clean.shape[1]
    pts = np.full((C, N, P, 2), np.nan)
    pts[:, :, 0] = clean + rng.normal(0, 0.3, size=clean.shape)
    far = rng.random((C, N)) < 0.25                       # the true detection is replaced by a distractor
    pts[:, :, 0][far] += rng.normal(0, 50.0, size=(int(far.sum()), 2))
    has2 = rng.random((C, N)) < (0.35 if C <= 8 else 0.12)   # keeps the 16-camera products small
    pts[:, :, 1][has2] = (clean + rng.normal(0, 30.0, size=clean.shape))[has2]
    pts[rng.random((C, N, P)) < 0.3] = np.nan
    o = og.triangulate_possible(cams, pts, min_cams=mc, return_stats=True)
    h = cg.triangulate_possible(pts, min_cams=mc, return_stats=True)
    assert np.array_equal(o[1], h[1])
    assert np.array_equal(o[4], h[4]) and np.array_equal(o[5], h[5])
    assert np.array_equal(o[2], h[2], equal_nan=True)
    assert _eq_nan(o[0], h[0]) and np.nanmax(np.abs(o[0] - h[0]), initial=0.0) <= P3D_TOL_MM
    assert np.abs(o[3] - h[3]).max() <= ERR_TOL_PX
    with pytest.raises(RuntimeError, match="cameras \\* candidates"):
        CameraGroup.from_dicts(synth.make_rig(8, "pinhole", seed=1)).triangulate_possible(np.zeros((8, 2, 5, 2)))


@pytest.mark.parametrize("mc", [1, 4, 6, 7])
def test_gpu_ransac_min_cams_range(mc):
    """min_cams from 1 (single-camera subsets are enumerated but can never be accepted) up to
    more than the valid views (only the full set is tried, cameras.py:691)."""
    seed = 6060
    dicts = synth.make_rig(6, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    X = synth.make_tracks(40, 2, seed=seed).reshape(-1, 3)
    p2 = synth.corrupt(og.project(cams, X), seed=seed, p_outlier=0.25, p_missing=0.15)
    o = og.triangulate_ransac(cams, p2, min_cams=mc, return_stats=True)
    h = cg.triangulate_ransac(p2, min_cams=mc, return_stats=True)
    assert np.array_equal(o[1], h[1]) and np.array_equal(o[4], h[4]) and np.array_equal(o[5], h[5])
    assert np.array_equal(o[2], h[2], equal_nan=True)
    assert np.nanmax(np.abs(o[0] - h[0]), initial=0.0) <= P3D_TOL_MM and np.abs(o[3] - h[3]).max() <= ERR_TOL_PX
    # a custom threshold goes through triangulate_possible (P = 1)
    o2 = og.triangulate_ransac(cams, p2, min_cams=mc, threshold=1.5, return_stats=True)
    h2 = cg.triangulate_possible(p2[:, :, None, :], min_cams=mc, threshold=1.5, return_stats=True)
    assert np.array_equal(o2[1], h2[1]) and np.array_equal(o2[4], h2[4]) and np.array_equal(o2[5], h2[5])


def test_gpu_ransac_output_selection():
    """triangulate_ransac(outputs=...): skipping points_2d (and picked_vals) changes nothing in what is returned,
    for numpy (host pipeline) and torch (device-resident) callers; points_2d is picked_vals applied to the input
    (what step4_aniposefiltering.py:299-300 relies on)."""
    import torch
    dicts = synth.make_rig(8, "pinhole", seed=77)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    X = synth.make_tracks(60, 2, seed=77).reshape(-1, 3)
    p2 = synth.corrupt(og.project(cams, X), seed=77, p_outlier=0.2, p_missing=0.1)
    full = cg.triangulate_ransac(p2, min_cams=3)
    assert np.array_equal(~np.isnan(full[2][..., 0]), full[1][..., 0])          # NaN pattern of points_2d == picked
    assert np.array_equal(np.where(full[1], p2, np.nan), full[2], equal_nan=True)
    for src in (p2, torch.from_numpy(p2).cuda()):
        a = cg.triangulate_ransac(src, min_cams=3, outputs="picked")
        b = cg.triangulate_ransac(src, min_cams=3, outputs="points_3d", return_stats=True)
        host = lambda t: t if isinstance(t, np.ndarray) else t.cpu().numpy()
        assert a[2] is None and b[1] is None and b[2] is None and len(b) == 6
        assert np.array_equal(host(a[0]), full[0], equal_nan=True) and np.array_equal(host(a[1]), full[1])
        assert np.array_equal(host(a[3]), full[3]) and np.array_equal(host(b[0]), full[0], equal_nan=True)
        assert np.array_equal(host(b[3]), full[3])
