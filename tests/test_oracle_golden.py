"""CPU tier: the oracle (oracle/) against the golden vectors produced by executing the
unmodified reference (oracle/make_golden.py).  Tolerances: pinhole arithmetic is restated
operation-for-operation and must be bit-exact; fisheye goes through libm tan/atan
(<= 1e-9); LAPACK-backed results (p3d) are compared at 1e-8 mm."""
import numpy as np
import pytest

from oracle import camera_math as cm
from oracle import cameragroup as og
from oracle import crossview as ocv
from oracle import fixtures

DLT = fixtures.golden_names("dlt")
RANSAC = fixtures.golden_names("ransac")
CROSS = fixtures.golden_names("crossview")


def _eq_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))


@pytest.mark.parametrize("name", DLT)
def test_oracle_dlt(name):
    g, cams = fixtures.load_golden(name)
    exact = "fisheye" not in name
    tol = 0.0 if exact else 1e-9
    und = og.undistort_points(cams, g["p2d"])
    assert _eq_nan(und, g["undistorted"])
    assert np.nanmax(np.abs(und - g["undistorted"])) <= tol
    p3d = og.triangulate(cams, g["p2d"])
    assert _eq_nan(p3d, g["p3d"])
    assert np.nanmax(np.abs(p3d - g["p3d"])) <= 1e-8
    p3d2 = og.triangulate(cams, g["undistorted"], undistort=False)
    assert np.nanmax(np.abs(p3d2 - g["p3d_noundist"])) <= 1e-8
    ef = og.reprojection_error(cams, g["p3d"], g["p2d"])
    assert _eq_nan(ef, g["err_full"])
    assert np.nanmax(np.abs(ef - g["err_full"])) <= max(tol, 1e-12)
    em = og.reprojection_error(cams, g["p3d"], g["p2d"], mean=True)
    assert _eq_nan(em, g["err_mean"])
    assert np.nanmax(np.abs(em - g["err_mean"])) <= max(tol, 1e-12)
    pr = og.project(cams, g["X_true"])
    assert np.abs(pr - g["proj_true"]).max() <= max(tol, 1e-12)
    # loop-faithful port (the timed CPU baseline) agrees as well
    n = min(60, g["p2d"].shape[1])
    p3l = og.triangulate_loops(cams, g["p2d"][:, :n])
    assert np.nanmax(np.abs(p3l - g["p3d"][:n])) <= 1e-8
    eml = og.reprojection_error_loops(cams, g["p3d"][:n], g["p2d"][:, :n], mean=True)
    assert np.nanmax(np.abs(eml - g["err_mean"][:n])) <= 1e-9


@pytest.mark.parametrize("name", RANSAC)
def test_oracle_ransac(name):
    g, cams = fixtures.load_golden(name)
    out, picked, p2d, err, sidx, nev = og.triangulate_ransac(
        cams, g["p2d"], min_cams=int(g["min_cams"]), return_stats=True)
    assert np.array_equal(picked, g["picked"])                       # bit-exact subsets
    assert np.array_equal(p2d, g["points_2d"], equal_nan=True)
    assert _eq_nan(out, g["p3d"])
    assert np.nanmax(np.abs(out - g["p3d"]), initial=0.0) <= 1e-8
    assert np.abs(err - g["errors"]).max() <= 1e-9
    assert int(nev.sum()) == int(g["n_subsets_evaluated"])            # same search length
    assert ((sidx >= 0) == np.isfinite(g["p3d"][:, 0])).all()


def test_oracle_ransac_loops_port():
    g, cams = fixtures.load_golden("ransac_edges_c8")
    out, picked, p2d, err = og.triangulate_ransac_loops(cams, g["p2d"])
    assert np.array_equal(picked, g["picked"])
    assert np.array_equal(p2d, g["points_2d"], equal_nan=True)
    assert np.nanmax(np.abs(out - g["p3d"])) <= 1e-8
    assert np.abs(err - g["errors"]).max() <= 1e-9


def test_oracle_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    K = np.array([[1234.5, 0, 1030.0], [0, 1250.1, 725.0], [0, 0, 1]])
    pts = rng.uniform([0, 0], [2048, 1536], size=(2000, 2))
    pts[5] = np.nan
    X = rng.uniform(-900, 900, size=(2000, 3))
    rvec, tvec = rng.normal(size=3) * 0.7, np.array([10.0, -20.0, 2000.0])
    for dist in ([-0.25, 0.08, 1e-3, -1e-3], [-0.2, 0.05, 1e-3, -2e-3, 0.01],
                 [-0.2, 0.05, 1e-3, -2e-3, 0.01, 0.01, 0.002, 0.001],
                 [-0.2, 0.05, 1e-3, -2e-3, 0.01, 0.01, 0.002, 0.001, 1e-3, -1e-3, 2e-3, 1e-4],
                 [-0.9, 0.0, 0, 0, 0]):
        dist = np.array(dist)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist).reshape(-1, 2)
        assert np.array_equal(ref, cm.undistort_pinhole(pts, K, dist), equal_nan=True)
        ref, _ = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, tvec, K, dist)
        assert np.array_equal(ref.reshape(-1, 2), cm.project_pinhole(X, rvec, tvec, K, dist))
    D = np.array([0.05, -0.01, 0.002, -0.0005])
    ref = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), K, D).reshape(-1, 2)
    assert np.nanmax(np.abs(ref - cm.undistort_fisheye(pts, K, D))) < 1e-12
    ref, _ = cv2.fisheye.projectPoints(X.reshape(-1, 1, 3), rvec, tvec, K, D)
    assert np.abs(ref.reshape(-1, 2) - cm.project_fisheye(X, rvec, tvec, K, D)).max() < 1e-9
    assert np.abs(cv2.Rodrigues(rvec)[0] - cm.rodrigues(rvec)).max() == 0.0


def test_oracle_omnidir_roundtrip():
    """The omnidir model is parity-unpinned (no cv2.omnidir here); at least
    undistort(project(X)) must return the perspective normalised coordinates."""
    from macaque_3d_pose_estimation_b200 import synth
    cams = fixtures.cams_from_dicts(synth.make_rig(4, "omnidir", seed=9))
    X = synth.make_tracks(20, 2, seed=9).reshape(-1, 3)
    for cam in cams:
        uv = cam.project(X)
        und = cam.undistort(uv)
        R = cm.rodrigues(cam.rvec)
        Xc = X @ R.T + cam.tvec
        assert np.abs(und - Xc[:, :2] / Xc[:, 2:3]).max() < 1e-9


def test_oracle_omnidir_at_xi_zero_equals_installed_opencv():
    """A partial pin of the Mei restatement against EXECUTED OpenCV: at xi = 0 (and zero skew) the unified model is
    the pinhole model with (k1, k2, p1, p2) — projectPoints must equal cv2.projectPoints, and the 20-iteration
    undistortion must agree with cv2.undistortPointsIter run to convergence.  The xi-dependent sphere lift is what
    remains unpinned."""
    import cv2
    rng = np.random.default_rng(31)
    for trial in range(4):
        K = np.array([[900.0 + 40 * trial, 0.0, 1010.0], [0.0, 880.0 + 30 * trial, 760.0], [0.0, 0.0, 1.0]])
        D = np.array([-0.12 + 0.05 * trial, 0.03, 0.001 * trial, -0.0015])
        rvec = rng.normal(0, 0.4, 3)
        tvec = np.array([30.0, -50.0, 2500.0]) + rng.normal(0, 40, 3)
        X = rng.uniform([-600, -500, -300], [600, 500, 300], size=(200, 3))
        ref, _ = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, tvec, K, D)
        got = cm.project_omnidir(X, rvec, tvec, K, 0.0, D)
        assert np.abs(got - ref.reshape(-1, 2)).max() < 1e-9
        uv = ref.reshape(-1, 2) + rng.normal(0, 0.5, size=(200, 2))
        crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 200, 1e-14)
        und_ref = cv2.undistortPointsIter(uv.reshape(-1, 1, 2), K, D, None, None, crit).reshape(-1, 2)
        und = cm.undistort_omnidir(uv, K, D, 0.0)
        assert np.abs(und - und_ref).max() < 1e-9


@pytest.mark.parametrize("name", CROSS)
def test_oracle_crossview(name):
    g, cams = fixtures.load_golden(name)
    for f in range(int(g["n_frames"])):
        kp, dim = g["f%d_kp" % f], g["f%d_dimGroup" % f]
        aff = ocv.geometry_affinity(cams, kp, dim)
        ref = g["f%d_aff" % f]
        assert _eq_nan(aff, ref)
        assert np.nanmax(np.abs(aff - ref)) <= 1e-9
        match = ocv.match_svt(g["f%d_W" % f], dim, alpha=0.5, lam=50.0)
        assert np.array_equal(match, g["f%d_match" % f])
        p3 = ocv.triangulate_ls(cams, g["f%d_ls_xy" % f], g["f%d_ls_use" % f])
        ref3 = g["f%d_ls_p3d" % f]
        assert _eq_nan(p3, ref3)
        assert np.nanmax(np.abs(p3 - ref3)) <= 1e-7


@pytest.mark.parametrize("name", fixtures.golden_names("possible"))
def test_oracle_triangulate_possible(name):
    """triangulate_possible with P > 1 candidates per camera (cameras.py:639-724)."""
    g, cams = fixtures.load_golden(name)
    out, picked, p2d, err = og.triangulate_possible(cams, g["points"], min_cams=int(g["min_cams"]))
    assert np.array_equal(picked, g["picked"])
    assert np.array_equal(p2d, g["points_2d"], equal_nan=True)
    assert _eq_nan(out, g["out"]) and np.nanmax(np.abs(out - g["out"]), initial=0.0) <= 1e-8
    assert np.abs(err - g["errors"]).max() <= 1e-10
