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
