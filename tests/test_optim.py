"""optim_points / optim_points_jointlenfix (reference cameras.py:1116-1270).

Parity bar:
  (i)   the residual vector of _error_fun_triangulation: same layout, values <= 1e-9 against goldens
        produced by EXECUTING the reference (oracle/make_golden.py case_optim);
  (ii)  the solver's exact Jacobian blocks against central finite differences of (i);
  (iii) from the reference's own x0 the GPU Levenberg-Marquardt ends at a cost <= the cost of what the
        reference's least_squares(ftol=1e-3) returned; the point-wise deviation is reported, not asserted
        to 0.01 mm: the reference stops far from the minimiser (its result is a property of scipy's
        trust-region trajectory), so two correct solvers differ by more than that.
CPU tier: the numpy oracle and the product's host-side pre-processing against the goldens."""
import numpy as np
import pytest

from macaque_3d_pose_estimation_b200 import optim as popt
from oracle import fixtures
from oracle import optim as oopt

NAMES = fixtures.golden_names("optim")


def _kw(g):
    return dict(scale_length=float(g["scale_length"]), scale_length_weak=float(g["scale_length_weak"]),
                reproj_error_threshold=float(g["reproj_error_threshold"]), reproj_loss=str(g["reproj_loss"]),
                n_deriv_smooth=int(g["n_deriv"]))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_and_host_preprocessing_match_reference(name):
    g, _ = fixtures.load_golden(name)
    cams = fixtures.cams_from_arrays(g)
    cons = [tuple(c) for c in g["constraints"]]
    consw = [tuple(c) for c in g["constraints_weak"]]
    scores = g["scores"] if "scores" in g else None
    # oracle restatement == executed reference
    s_full, intp = oopt.scale_smooth_full(g["p3d0"], float(g["scale_smooth"]))
    assert np.array_equal(intp, g["intp"]) and s_full == float(g["scale_smooth_full"])
    x0 = oopt.initialize_params(intp, cons, consw)
    x0[~np.isfinite(x0)] = 0
    assert np.array_equal(x0, g["x0"])
    for x, r in ((g["x0"], g["r0"]), (g["x1"], g["r1"])):
        mine = oopt.error_fun(cams, x, g["points"], cons, consw, scores, float(g["scale_smooth_full"]), **_kw(g))
        assert mine.shape == r.shape
        assert np.abs(mine - r).max() <= 1e-9 * max(1.0, np.abs(r).max())
    # the product's host-side pre-processing == executed reference
    assert np.array_equal(popt.interpolate_columns(g["p3d0"]), g["intp"])
    assert abs(popt.smoothness_scale(g["intp"], float(g["scale_smooth"])) - float(g["scale_smooth_full"])) <= \
        1e-12 * float(g["scale_smooth_full"])
    strong, weak = popt.initial_lengths(g["intp"], cons, consw)
    assert np.array_equal(np.hstack([strong, weak]), g["x0"][g["intp"].size:])


@pytest.mark.parametrize("name", ["optim_c8_fix", "optim_c4_n1_huber"])
def test_port_reproduces_reference_least_squares(name):
    """oracle.optim.optim_points_port (the CPU baseline of bench.py's optim line) returns what the executed
    reference returned: same objective, same sparsity pattern, same scipy call."""
    g, _ = fixtures.load_golden(name)
    cams = fixtures.cams_from_arrays(g)
    cons = [tuple(c) for c in g["constraints"]]
    consw = [tuple(c) for c in g["constraints_weak"]]
    new, jl, cost = oopt.optim_points_port(
        cams, g["points"], g["p3d0"], cons, consw, float(g["scale_smooth"]), scores=g["scores"] if "scores" in g else None,
        joint_len=(g["x0"][g["intp"].size:] if int(g["fix"]) else None), **_kw(g))
    assert np.abs(new - g["opt_p3d"]).max() <= 1e-6
    assert abs(cost - float(g["opt_cost"])) <= 1e-9 * cost


def test_median_filter_matches_scipy_reflect_form():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(40, 3, 2))
    ref = np.apply_along_axis(oopt.medfilt_data, 0, a, size=7)
    assert np.array_equal(popt.median_filter_columns(a, 7), ref)


# ---------------------------------------------------------------------------------------------
# GPU tier
# ---------------------------------------------------------------------------------------------

def _group(g):
    from tests.test_gpu_parity import group_from_golden
    import __graft_entry__ as ge
    ge.build_library()
    return group_from_golden(g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_residual_vector_matches_reference(name):
    g, _ = fixtures.load_golden(name)
    cg = _group(g)
    cons, consw = g["constraints"], g["constraints_weak"]
    scores = g["scores"] if "scores" in g else None
    for x, r in ((g["x0"], g["r0"]), (g["x1"], g["r1"])):
        mine = popt.error_fun(cg, x, g["points"], cons, consw, scores, float(g["scale_smooth_full"]), **_kw(g))
        assert mine.shape == r.shape, "residual layout differs from the reference's"
        assert np.abs(mine - r).max() <= 1e-9 * max(1.0, np.abs(r).max())
    if int(g["fix"]):
        n3 = g["intp"].size
        mine = popt.error_fun(cg, g["x0"][:n3], g["points"], cons, consw, scores, float(g["scale_smooth_full"]),
                              joint_len=g["x0"][n3:], **_kw(g))
        assert np.abs(mine - g["r0_fix"]).max() <= 1e-9 * max(1.0, np.abs(g["r0_fix"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_jacobian_blocks_match_finite_differences(name):
    g, _ = fixtures.load_golden(name)
    cg = _group(g)
    cons, consw = g["constraints"], g["constraints_weak"]
    scores = g["scores"] if "scores" in g else None
    rng = np.random.default_rng(1)
    x = g["x1"]
    args = (g["points"], cons, consw, scores, float(g["scale_smooth_full"]))
    for trial in range(3):
        v = rng.normal(size=x.shape)
        if trial == 1:
            v[:g["intp"].size] = 0                      # lengths only
        if trial == 2:
            v[g["intp"].size:] = 0                      # points only
        h = 1e-5
        fd = (popt.error_fun(cg, x + h * v, *args, **_kw(g)) - popt.error_fun(cg, x - h * v, *args, **_kw(g))) / (2 * h)
        jv = popt.jvp(cg, x, v, *args, **_kw(g))
        # |e| and the huber switch are not differentiable at isolated points: compare where the
        # residual is not within the step of a kink, and require the rest to agree tightly
        scale = np.maximum(1.0, np.abs(fd))
        ok = np.abs(fd - jv) <= 1e-5 * scale
        assert ok.mean() > 0.999, "Jacobian-vector product differs from finite differences (%.4f ok)" % ok.mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_solver_reaches_reference_cost(name, record_property):
    g, _ = fixtures.load_golden(name)
    cg = _group(g)
    cons, consw = g["constraints"], g["constraints_weak"]
    scores = g["scores"] if "scores" in g else None
    kw = dict(constraints=cons, constraints_weak=consw, scale_smooth=float(g["scale_smooth"]), scores=scores, **_kw(g))
    if int(g["fix"]):
        new, jl, info = cg.optim_points_jointlenfix(g["points"], g["p3d0"], g["x0"][g["intp"].size:], return_info=True, **kw)
    else:
        new, jl, info = cg.optim_points(g["points"], g["p3d0"], return_info=True, **kw)
    assert np.array_equal(info["x0"], g["x0"])                                  # same start vector
    assert abs(info["cost0"] - float(g["x0_cost"])) <= 1e-9 * float(g["x0_cost"])
    assert new.shape == g["opt_p3d"].shape and jl.shape == g["opt_joint_len"].shape
    assert info["cost"] <= float(g["opt_cost"]) * (1 + 1e-9), \
        "GPU solver stopped at cost %.6g, the reference's least_squares at %.6g" % (info["cost"], float(g["opt_cost"]))
    dev = np.linalg.norm(new - g["opt_p3d"], axis=-1)
    err_ref = np.linalg.norm(g["opt_p3d"] - g["X_true"], axis=-1)
    err_new = np.linalg.norm(new - g["X_true"], axis=-1)
    print("\\n%s: cost x0 %.5g | reference %.5g | GPU %.5g (%d LM steps, %d CG its); deviation from the reference's "
          "points: median %.3f mm, p95 %.3f mm, max %.3f mm; error to ground truth: reference median %.3f mm, GPU %.3f mm"
          % (name, info["cost0"], float(g["opt_cost"]), info["cost"], info["lm_steps"], info["cg_iterations"],
             np.median(dev), np.percentile(dev, 95), dev.max(), np.median(err_ref), np.median(err_new)))
    # not farther from the truth than the reference's result
    assert np.median(err_new) <= np.median(err_ref) * 1.05 + 1e-6


def _template_config():
    """The [triangulation] block of the reference's configs/config_tmpl.toml:56-97 (optim = true, the default)."""
    return {"triangulation": {
        "ransac": False, "optim": True, "score_threshold": 0.5, "scale_smooth": 3, "scale_length": 5,
        "scale_length_weak": 2, "reproj_error_threshold": 3, "n_deriv_smooth": 2,
        "constraints": [["nose", "left_eye"], ["nose", "right_eye"], ["left_eye", "right_eye"], ["nose", "left_ear"],
                        ["nose", "right_ear"], ["left_eye", "left_ear"], ["right_eye", "right_ear"],
                        ["left_ear", "right_ear"], ["left_shoulder", "left_ear"], ["right_shoulder", "right_ear"],
                        ["left_shoulder", "right_shoulder"], ["left_shoulder", "left_elbow"],
                        ["left_elbow", "left_wrist"], ["right_shoulder", "right_elbow"],
                        ["right_elbow", "right_wrist"], ["left_hip", "right_hip"], ["left_hip", "left_knee"],
                        ["left_knee", "left_ankle"], ["right_hip", "right_knee"], ["right_knee", "right_ankle"]],
        "constraints_weak": [["left_shoulder", "left_hip"], ["right_shoulder", "right_hip"],
                             ["left_shoulder", "right_hip"], ["right_shoulder", "left_hip"],
                             ["left_shoulder", "right_shoulder"], ["left_hip", "right_hip"], ["left_eye", "nose"],
                             ["right_eye", "nose"], ["left_eye", "left_ear"], ["right_eye", "right_ear"],
                             ["left_ear", "right_ear"]]}}


def test_template_constraints_resolve():
    from macaque_3d_pose_estimation_b200 import pipeline3d
    from oracle import make_golden as mg
    cfg = _template_config()
    assert [tuple(c) for c in pipeline3d.load_constraints(cfg, pipeline3d.BODYPARTS)] == mg.MACAQUE_CONSTRAINTS
    assert [tuple(c) for c in pipeline3d.load_constraints(cfg, pipeline3d.BODYPARTS, "constraints_weak")] == \
        mg.MACAQUE_CONSTRAINTS_WEAK


@pytest.mark.gpu
def test_gpu_step4_runs_template_config_with_optim(tmp_path):
    """pipeline3d.run_stage on the template's default branch (optim = true): per animal the initial
    triangulation + optim_points, the reference's score / error / joint_len bookkeeping (step4:228-291, 332-339)."""
    import pickle
    import __graft_entry__ as ge
    ge.build_library()
    from macaque_3d_pose_estimation_b200 import pipeline3d, synth
    from macaque_3d_pose_estimation_b200.cameras import CameraGroup
    from oracle import cameragroup as og
    dicts = synth.make_rig(8, "pinhole", seed=61)
    cg = CameraGroup.from_dicts(dicts)
    cams = fixtures.cams_from_dicts(dicts)
    F, A, J = 40, 2, 17
    X = synth.make_tracks(F, A, seed=61)                                         # (F, A, J, 3)
    p2 = synth.corrupt(og.project(cams, X.reshape(-1, 3)), seed=61, noise=0.6, p_missing=0.1)
    kp = np.zeros((F, J, A, 3, 8))
    kp[:, :, :, :2, :] = p2.reshape(8, F, A, J, 2).transpose(1, 3, 2, 4, 0)
    kp[:, :, :, 2, :] = 0.9
    kp[np.isnan(kp[:, :, :, 0, :])[:, :, :, None, :].repeat(3, axis=3)] = 0.0
    kp[:, :, :, 2, :][kp[:, :, :, 0, :] == 0] = 0.1                              # missing views: low score
    cg.dump(str(tmp_path / "calibration.toml"))
    with open(tmp_path / "kp2d_f.pickle", "wb") as f:
        pickle.dump(kp, f)
    cfg = _template_config()
    data = pipeline3d.run_stage(str(tmp_path), [d["name"] for d in dicts], config=cfg)
    assert data["kp3d"].shape == (A, F, J, 3) and data["kp3d_score"].shape == (A, F, J)
    assert len(data["joint_len"]) == A and data["joint_len"][0].shape == (31,)
    assert np.isfinite(data["kp3d"]).all()
    with open(tmp_path / "kp3d.pickle", "rb") as f:
        again = pickle.load(f)
    assert np.array_equal(again["kp3d"], data["kp3d"])
    # animal 0 == the direct call sequence of step4:236-258
    pts = np.ascontiguousarray(kp.transpose(2, 4, 0, 1, 3)[0][..., :2])
    pts[kp.transpose(2, 4, 0, 1, 3)[0][..., 2] < 0.5] = np.nan
    init = cg.triangulate(pts.reshape(8, -1, 2)).reshape(F, J, 3)
    tri = cfg["triangulation"]
    direct, jl = cg.optim_points(pts, init, constraints=pipeline3d.load_constraints(cfg, pipeline3d.BODYPARTS),
                                 constraints_weak=pipeline3d.load_constraints(cfg, pipeline3d.BODYPARTS, "constraints_weak"),
                                 scale_smooth=tri["scale_smooth"], scale_length=tri["scale_length"],
                                 scale_length_weak=tri["scale_length_weak"], n_deriv_smooth=tri["n_deriv_smooth"],
                                 reproj_error_threshold=tri["reproj_error_threshold"])
    # (the pipeline loads the rig back from calibration.toml: parameters agree to the printed digits)
    assert np.abs(direct - data["kp3d"][0]).max() <= 1e-6
    # sanity: the refinement stays near the truth (the smoothness prior pulls on this 15 mm / frame random walk)
    assert np.median(np.linalg.norm(data["kp3d"][0] - X[:, 0], axis=-1)) < 25.0
    # fixed limb lengths: the _jointlenfix branch and its file name (step4:176-181, 259-271, 334-336)
    d2 = pipeline3d.run_stage(str(tmp_path), [d["name"] for d in dicts], config=cfg, joint_len=np.array(data["joint_len"]))
    assert (tmp_path / "kp3d_fxdJointLen.pickle").exists() and np.isfinite(d2["kp3d"]).all()
