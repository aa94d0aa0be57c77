"""CPU tier: the kernels' per-point arithmetic (csrc/m3d_math.cuh, m3d_point.cuh compiled
for the host by tests/harness.py) against the golden vectors and the oracle.  This is the
same source the CUDA kernels compile, so a failure here is a kernel-math failure that
does not need a GPU to reproduce.

Tolerances (BASELINE.json north_star): selected subsets / inlier masks bit-exact; 3D
points within 1e-4 relative or 0.01 mm (we assert 1e-6 mm); reprojection errors within
1e-3 px (we assert 1e-7 px)."""
import numpy as np
import pytest

from macaque_3d_pose_estimation_b200 import synth
from oracle import cameragroup as og
from oracle import fixtures
from tests import harness as hh

DLT = fixtures.golden_names("dlt")
RANSAC = fixtures.golden_names("ransac")
P3D_TOL_MM = 1e-6
ERR_TOL_PX = 1e-7


@pytest.mark.parametrize("name", DLT)
def test_math_dlt_golden(name, host_harness):
    g, cams = fixtures.load_golden(name)
    und = hh.undistort(cams, g["p2d"])
    assert np.array_equal(np.isnan(und), np.isnan(g["undistorted"]))
    assert np.nanmax(np.abs(und - g["undistorted"])) <= 1e-12
    p3d, err = hh.triangulate_error(cams, g["p2d"])
    assert np.array_equal(np.isnan(p3d), np.isnan(g["p3d"]))
    assert np.nanmax(np.abs(p3d - g["p3d"])) <= P3D_TOL_MM
    assert np.array_equal(np.isnan(err), np.isnan(g["err_mean"]))
    assert np.nanmax(np.abs(err - g["err_mean"])) <= ERR_TOL_PX
    p3d2, _ = hh.triangulate_error(cams, g["undistorted"], undistort=False)
    assert np.nanmax(np.abs(p3d2 - g["p3d_noundist"])) <= P3D_TOL_MM
    assert np.abs(hh.project(cams, g["X_true"]) - g["proj_true"]).max() <= 1e-9
    assert np.abs(hh.extrinsics(cams) - np.array([c.extrinsics() for c in cams])).max() == 0.0


@pytest.mark.parametrize("name", RANSAC)
def test_math_ransac_golden(name, host_harness):
    g, cams = fixtures.load_golden(name)
    out, picked, p2d, err, sidx, nev = hh.ransac(cams, g["p2d"], min_cams=int(g["min_cams"]))
    assert np.array_equal(picked, g["picked"])
    assert np.array_equal(p2d, g["points_2d"], equal_nan=True)
    assert np.array_equal(np.isnan(out), np.isnan(g["p3d"]))
    assert np.nanmax(np.abs(out - g["p3d"]), initial=0.0) <= P3D_TOL_MM
    assert np.abs(err - g["errors"]).max() <= ERR_TOL_PX
    assert int(nev.sum()) == int(g["n_subsets_evaluated"])


@pytest.mark.parametrize("n_cams,model,kw", [
    (8, "pinhole", dict(p_outlier=0.2, p_missing=0.1)),
    (8, "pinhole", dict(noise=2.0, p_outlier=0.4, p_missing=0.1)),     # arg-min (pass 2) heavy
    (8, "pinhole8", dict(p_outlier=0.2, p_missing=0.2)),
    (8, "fisheye", dict(p_outlier=0.2, p_missing=0.1)),
    (5, "pinhole", dict(p_outlier=0.3, p_missing=0.1)),
    (2, "pinhole", dict(p_outlier=0.1, p_missing=0.1)),
])
def test_math_ransac_random_vs_oracle(n_cams, model, kw, host_harness):
    seed = 4242 + n_cams
    cams = fixtures.cams_from_dicts(synth.make_rig(n_cams, model, seed=seed))
    X = synth.make_tracks(25, 4, seed=seed).reshape(-1, 3)
    p2 = synth.corrupt(og.project(cams, X), seed=seed, **kw)
    for min_cams in (2, 3):
        o = og.triangulate_ransac(cams, p2, min_cams=min_cams, return_stats=True)
        h = hh.ransac(cams, p2, min_cams=min_cams)
        assert np.array_equal(o[1], h[1])
        assert np.array_equal(o[4], h[4])
        assert np.array_equal(o[5], h[5])
        assert np.array_equal(o[2], h[2], equal_nan=True)
        assert np.nanmax(np.abs(o[0] - h[0]), initial=0.0) <= P3D_TOL_MM
        assert np.abs(o[3] - h[3]).max() <= ERR_TOL_PX


def test_math_omnidir_vs_oracle(host_harness):
    """Omnidir (parity unpinned upstream): kernel arithmetic == oracle restatement."""
    cams = fixtures.cams_from_dicts(synth.make_rig(6, "omnidir", seed=11))
    X = synth.make_tracks(20, 2, seed=11).reshape(-1, 3)
    proj = og.project(cams, X)
    assert np.abs(hh.project(cams, X) - proj).max() <= 1e-9
    p2 = synth.corrupt(proj, seed=11, p_missing=0.1)
    assert np.nanmax(np.abs(hh.undistort(cams, p2) - og.undistort_points(cams, p2))) <= 1e-12
    p3d, err = hh.triangulate_error(cams, p2)
    ref = og.triangulate(cams, p2)
    assert np.nanmax(np.abs(p3d - ref)) <= P3D_TOL_MM
    assert np.nanmax(np.abs(err - og.reprojection_error(cams, ref, p2, mean=True))) <= ERR_TOL_PX


def test_math_degenerate_geometry(host_harness):
    """Coincident cameras / identical rays: the Newton branch must hand over to the Jacobi
    fallback without NaN-poisoning well-posed neighbours."""
    dicts = synth.make_rig(3, "pinhole", seed=5)
    dicts[1] = dict(dicts[0], name="2")          # camera 2 == camera 1
    cams = fixtures.cams_from_dicts(dicts)
    X = synth.make_tracks(5, 1, seed=5).reshape(-1, 3)
    p2 = og.project(cams, X)
    p3d, err = hh.triangulate_error(cams, p2)
    assert np.abs(p3d - og.triangulate(cams, p2)).max() < P3D_TOL_MM   # third camera resolves the depth
    assert np.abs(p3d - X).max() < 0.05           # (5 un-converged undistort iterations)
    p2[2] = np.nan                                # only the two coincident views remain
    p3d, _ = hh.triangulate_error(cams, p2)       # rank-deficient: anything but a crash
    assert p3d.shape == X.shape
