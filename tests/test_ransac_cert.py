"""Pruned subset search (csrc/m3d_cert.h, csrc/m3d_ransac_cert.cuh).

CPU tier: (1) the pair certificate is SOUND — for random 3D points anywhere in space (in front of,
behind and far from the cameras) it never calls a pair incompatible whose two residuals fit in the
budget; (2) the host build of ransac_cert_point (the function the kernel runs per thread) selects the
reference's subsets on the goldens and on seeded sweeps, with and without pruning.
GPU tier: k_ransac_cert == exhaustive kernels == oracle."""
import numpy as np
import pytest

from macaque_3d_pose_estimation_b200 import synth
from oracle import cameragroup as og
from oracle import fixtures
from tests import harness


def _distort(cam, xy):
    k1, k2, p1, p2, k3 = (list(cam.dist) + [0.0] * 5)[:5]
    x, y = xy[..., 0], xy[..., 1]
    r2 = x * x + y * y
    cd = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3
    xd = x * cd + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * cd + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([xd * cam.K[0, 0] + cam.K[0, 2], yd * cam.K[1, 1] + cam.K[1, 2]], -1)


def _bad_pairs(cams, p2d, rho):
    """numpy statement of the kernel's pair test at residual budget rho (N,): bad[n, a, b]."""
    C, N, _ = p2d.shape
    ok, inv, E = harness.cert_tables(cams)
    U = og.undistort_points(cams, p2d)
    with np.errstate(invalid="ignore"):
        delta = np.stack([np.linalg.norm(p2d[c] - _distort(cams[c], U[c]), axis=-1) for c in range(C)])
    bad = np.zeros((N, C, C), bool)
    p = 0
    for a in range(C):
        for b in range(a + 1, C):
            e = E[p]
            p += 1
            if inv[a] <= 0 or inv[b] <= 0:
                continue
            Em = e[:9].reshape(3, 3)
            xa = np.concatenate([U[a], np.ones((N, 1))], 1)
            xb = np.concatenate([U[b], np.ones((N, 1))], 1)
            Ea = xa @ Em.T
            Etb = xb @ Em
            F = np.abs(np.sum(xb * Ea, 1))
            A = (np.abs(Etb[:, 0]) + np.abs(Etb[:, 1])) * inv[a]
            B = (np.abs(Ea[:, 0]) + np.abs(Ea[:, 1])) * inv[b]
            D = rho + delta[a] + delta[b]
            rhs = (np.maximum(A, B) + 0.25 * e[9] * D) * D
            with np.errstate(invalid="ignore"):
                bd = F > rhs * (1 + 1e-9)
            bad[:, a, b] = bd
            bad[:, b, a] = bd
    return bad


@pytest.mark.parametrize("seed", [20261020, 3, 41])
def test_pair_certificate_is_sound(seed):
    C = 8
    cams = fixtures.cams_from_dicts(synth.make_rig(C, "pinhole", seed=seed))
    ok, inv, E = harness.cert_tables(cams)
    if bin(ok).count("1") < 2:
        pytest.skip("fewer than two certified cameras in this rig")
    rng = np.random.default_rng(seed)
    X = synth.make_tracks(150, 2, seed=seed).reshape(-1, 3)
    N = X.shape[0]
    p2d = synth.corrupt(og.project(cams, X), seed=seed, p_outlier=0.3, sigma_outlier=20.0)
    k = np.full(N, C)
    T = 0.5
    rho = T * (k - 1) * (1 + 1e-9) + 1e-6
    bad = _bad_pairs(cams, p2d, rho)
    assert bad.any() and not bad.all()
    # candidate 3D points: near the true point, anywhere in the room, far away, mirrored behind the rig
    trials = [X + rng.normal(0, s, X.shape) for s in (1.0, 5.0, 30.0, 300.0)]
    trials += [rng.normal(0, 3000, X.shape), X * rng.uniform(-3, 3, (N, 1)), rng.normal(0, 1e5, X.shape)]
    # and the most dangerous ones: the two-view DLT points of every pair themselves
    for a in range(C):
        for b in range(a + 1, C):
            sub = [cams[a], cams[b]]
            trials.append(og.triangulate(sub, p2d[[a, b]]))
    worst = np.inf
    for Xt in trials:
        with np.errstate(all="ignore"):
            pr = og.project(cams, Xt)
            e = np.linalg.norm(p2d - pr, axis=-1)          # (C, N)
        for a in range(C):
            for b in range(a + 1, C):
                m = bad[:, a, b] & np.isfinite(e[a]) & np.isfinite(e[b])
                if m.any():
                    gap = (e[a] + e[b])[m] - T * (k[m] - 1)
                    worst = min(worst, gap.min())
    assert worst > 0, "a certified-bad pair fits the residual budget of some 3D point (gap %g px)" % worst


def test_certify_rejects_foldback_models():
    base = synth.make_rig(2, "pinhole", seed=1)
    base[0]["distortions"] = [-0.2, 0.0, 0.0, 0.0, 0.0]      # barrel k1 alone folds back at r = 2.2
    base[1]["distortions"] = [0.1, 0.0, 1e-3, -1e-3, 0.0]    # pincushion + tangential: monotone
    ok, inv, _ = harness.cert_tables(fixtures.cams_from_dicts(base))
    assert ok == 0b10 and inv[0] == 0 and inv[1] > 0
    for model in ("fisheye", "pinhole8"):
        ok, _, _ = harness.cert_tables(fixtures.cams_from_dicts(synth.make_rig(3, model, seed=2)))
        assert ok == 0


@pytest.mark.parametrize("C,seed,min_cams", [(8, 20261020, 2), (8, 20261020, 3), (8, 77, 2), (6, 5, 2), (5, 8, 4),
                                              (4, 9, 2), (3, 10, 2), (2, 12, 2), (11, 11, 3)])
@pytest.mark.parametrize("use_cert", [True, False])
def test_host_pruned_search_matches_oracle(C, seed, min_cams, use_cert):
    cams = fixtures.cams_from_dicts(synth.make_rig(C, "pinhole", seed=seed))
    X = synth.make_tracks(40 if C <= 8 else 6, 2, seed=seed).reshape(-1, 3)
    p2d = synth.corrupt(og.project(cams, X), seed=seed, p_outlier=0.25, p_missing=0.12)
    p2d[3 % C, 5, 1] = np.nan                                  # y missing only: picked but unusable
    p2d[:, 7] = np.nan                                         # nothing visible
    p2d[1:, 9] = np.nan                                        # one camera only
    ref = og.triangulate_ransac(cams, p2d, min_cams=min_cams, return_stats=True)
    got = harness.ransac_cert(cams, p2d, min_cams=min_cams, use_cert=use_cert)
    assert np.array_equal(ref[4], got[4]) and np.array_equal(ref[5], got[5])
    assert np.array_equal(ref[1], got[1])
    assert np.array_equal(ref[2], got[2], equal_nan=True)
    assert np.array_equal(np.isnan(ref[0]), np.isnan(got[0]))
    sel = ref[4] >= 0
    assert np.abs(ref[0][sel] - got[0][sel]).max() <= 1e-5
    assert np.abs(ref[3] - got[3]).max() <= 1e-7


@pytest.mark.parametrize("thr,init_best", [(0.3, 200.0), (2.0, 200.0), (5.0, 1.0), (0.5, 0.4)])
def test_host_pruned_search_thresholds(thr, init_best):
    cams = fixtures.cams_from_dicts(synth.make_rig(8, "pinhole", seed=20261020))
    X = synth.make_tracks(30, 2, seed=4).reshape(-1, 3)
    p2d = synth.corrupt(og.project(cams, X), seed=4, noise=0.5, p_outlier=0.3, sigma_outlier=8.0, p_missing=0.1)
    ref = og.triangulate_ransac(cams, p2d, threshold=thr, init_best=init_best, return_stats=True)
    got = harness.ransac_cert(cams, p2d, threshold=thr, init_best=init_best)
    assert np.array_equal(ref[4], got[4]) and np.array_equal(ref[5], got[5])
    assert np.array_equal(ref[1], got[1])


@pytest.mark.parametrize("name", [n for n in fixtures.golden_names("ransac")])
def test_host_pruned_search_goldens(name):
    g, _ = fixtures.load_golden(name)
    cams = fixtures.cams_from_arrays(g)
    if any(c.model != 0 or c.dist.size > 5 for c in cams):
        pytest.skip("pruned search covers plain pinhole rigs")
    mc = int(g["min_cams"]) if "min_cams" in g else 2
    got = harness.ransac_cert(cams, g["p2d"], min_cams=mc)
    assert np.array_equal(got[1].reshape(g["picked"].shape), g["picked"])
    assert np.abs(got[3] - g["errors"]).max() <= 1e-7


@pytest.mark.parametrize("case", ["bench", "barrel", "sparse"])
@pytest.mark.parametrize("use_cert", [True, False])
def test_host_straight_line_form_equals_general_form_bit_for_bit(case, use_cert):
    """The compile-time-count instantiation of the headline kernels (cert_undistort / cert_pairs as straight-line
    code over the eight cameras, invalid views run on NaN and are masked, one replay branch for icdist < 0) against
    the run-time-count form with its per-camera branches: every output identical to the last bit."""
    dicts = synth.make_rig(8, "pinhole", seed=31)
    rng = np.random.default_rng(31)
    if case == "barrel":   # k1 << 0: icdist < 0 at the image corners (cv2's bail-out, the replay branch)
        for d in dicts:
            d["distortions"] = [-0.9, 0.0, 0.0, 0.0, 0.0]
    cams = fixtures.cams_from_dicts(dicts)
    X = synth.make_tracks(60, 2, seed=31).reshape(-1, 3)
    p2d = synth.corrupt(og.project(cams, X), seed=31, p_outlier=0.25, p_missing=0.6 if case == "sparse" else 0.12)
    if case == "barrel":   # observations all over the image, corners included
        n = p2d.shape[1]
        far = rng.random((8, n)) < 0.3
        p2d[far] = np.stack([rng.uniform(0, 2048, far.sum()), rng.uniform(0, 1536, far.sum())], axis=1)
    p2d[2, 3, 1] = np.nan          # y missing only
    p2d[4, 4, 0] = np.nan          # x missing only
    p2d[:, 5] = np.nan
    p2d[1:, 6] = np.nan
    a = harness.ransac_cert(cams, p2d, use_cert=use_cert, return_solved=True)
    b = harness.ransac_cert(cams, p2d, use_cert=use_cert, return_solved=True, general=True)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and x.tobytes() == y.tobytes()
    if case == "barrel":   # the replay branch is really taken: some icdist of the five iterations is negative
        K = np.array(dicts[0]["matrix"])
        x0 = (p2d[0, :, 0] - K[0, 2]) / K[0, 0]
        y0 = (p2d[0, :, 1] - K[1, 2]) / K[1, 1]
        x, y, neg = x0.copy(), y0.copy(), np.zeros(x0.shape, dtype=bool)
        with np.errstate(all="ignore"):
            for _ in range(5):
                icd = 1.0 / (1.0 - 0.9 * (x * x + y * y))
                neg |= icd < 0
                x, y = x0 * icd, y0 * icd
        assert neg.sum() >= 3


# ---------------------------------------------------------------------------------------------
# GPU tier
# ---------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("C,seed,min_cams,n_frames", [(8, 20261020, 2, 3000), (8, 20261020, 3, 1500), (8, 5, 2, 1500),
                                                       (6, 5, 2, 800), (4, 9, 2, 800), (3, 10, 2, 500),
                                                       (12, 11, 3, 200), (16, 13, 2, 60)])
def test_gpu_pruned_equals_exhaustive_and_oracle(C, seed, min_cams, n_frames):
    import torch
    import __graft_entry__ as ge
    ge.build()
    from macaque_3d_pose_estimation_b200 import _lib
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup
    dicts = synth.make_rig(C, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    cg = CameraGroup.from_dicts(dicts)
    lib = _lib.load()
    rig = cg._rig()
    mask = lib.m3d_rig_certified_mask(rig.handle)
    X = synth.make_tracks(n_frames, 2, seed=seed).reshape(-1, 3)
    p2d = synth.corrupt(og.project(cams, X), seed=seed, p_outlier=0.2, p_missing=0.1)
    p2d[3 % C, 5, 1] = np.nan
    p2d[:, 7] = np.nan
    p2d[1:, 9] = np.nan
    pts = torch.from_numpy(p2d).cuda()
    _lib.check(lib.m3d_rig_set_ransac_mode(rig.handle, 0))
    a = [t.cpu().numpy() for t in cg.triangulate_ransac(pts, min_cams=min_cams, return_stats=True)]
    _lib.check(lib.m3d_rig_set_ransac_mode(rig.handle, 1))
    b = [t.cpu().numpy() for t in cg.triangulate_ransac(pts, min_cams=min_cams, return_stats=True)]
    _lib.check(lib.m3d_rig_set_ransac_mode(rig.handle, 0))
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2], equal_nan=True)
    assert np.nanmax(np.abs(a[3] - b[3])) <= 1e-8
    if bin(mask).count("1") >= max(2, C - 1):                  # the pruned kernel really ran in mode 0
        n_or = min(p2d.shape[1], 3400 if C <= 8 else (170 if C <= 12 else 24))   # the oracle solves 2^C subsets per point
        ref = og.triangulate_ransac(cams, p2d[:, :n_or], min_cams=min_cams, return_stats=True)
        assert np.array_equal(ref[4], a[4][:n_or]) and np.array_equal(ref[5], a[5][:n_or])
        assert np.array_equal(ref[1], a[1][:, :n_or])
        assert np.abs(ref[3] - a[3][:n_or]).max() <= 1e-7
