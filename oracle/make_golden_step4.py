#!/usr/bin/env python
"""TEST INFRASTRUCTURE (not part of the product): golden vectors of the WHOLE 3D stage, produced by
executing the unmodified reference function ``step4_aniposefiltering.proc``
(/root/reference/src/pipeline/step4_aniposefiltering.py:89-339) on a small synthetic recording.

What makes the reference runnable here:
  * ``h5py`` is not installed: a stand-in module serves ``h5py.File(path)`` from nested dicts of arrays (the
    reference only does ``f[cam_id][field][()]``, :115-134);
  * ``configs/calibration_tmpl.toml`` declares ``omnidir = true`` for every camera, and ``cv2.omnidir`` is not
    installed: the run uses a copy of the template with ``omnidir = false`` (plain pinhole cameras, the model
    every benchmarked configuration uses).  ``configs/config_tmpl.toml`` is the reference's own, with the
    ``ransac`` / ``optim`` switches of the variant;
  * ``proc`` reads both templates relative to the working directory: the run happens in a scratch directory
    that holds the two files.

Every variant stores its inputs (kp2d.pickle array, the arrays behind the two h5 files, camera ids) and the
reference's outputs (calibration.toml text, kp2d_f.pickle, kp3d.pickle, joint_len.npy) in
``tests/golden/step4_<variant>.npz``.  Run from the repo root: ``python oracle/make_golden_step4.py``.
"""
import os
import pickle
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from macaque_3d_pose_estimation_b200 import synth  # noqa: E402

H5_STORE = {}     # file name -> {cam_id: {field: array}}


class _H5File:
    def __init__(self, path, mode="r"):
        self.data = H5_STORE[os.path.basename(path)]

    def __enter__(self):
        return {k: {f: np.array(a, copy=True) for f, a in v.items()} for k, v in self.data.items()}

    def __exit__(self, *a):
        return False


def import_reference_step4():
    sys.dont_write_bytecode = True
    for name in ("imgstore", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    h5 = types.ModuleType("h5py")
    h5.File = _H5File
    sys.modules["h5py"] = h5
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    # anipose/common.py imports the vendored aniposelib under its installed name; the spawn Pool of
    # filter_pose_viterbi (:151-186) re-imports everything in its children, which inherit sys.path
    sys.path.insert(0, os.path.join(REF, "src", "third_party"))
    from src.pipeline import step4_aniposefiltering as step4
    return step4


def synthetic_recording(n_cams, n_animals, n_frames, seed, return_tracks=False):
    """kp2d (A, F, C, J, 3) of a pinhole ring rig: projections + detector noise, a few gross jumps, low-score
    and missing detections; plus the arrays of cam_intrinsic.h5 / cam_extrinsic_optim.h5."""
    from oracle import cameragroup as og
    from oracle import fixtures
    rng = np.random.default_rng(seed)
    dicts = synth.make_rig(n_cams, "pinhole", seed=seed)
    cams = fixtures.cams_from_dicts(dicts)
    X = synth.make_tracks(n_frames, n_animals, seed=seed) * np.array([0.6, 0.6, 0.5])      # (F, A, J, 3)
    X = X + 25.0 * np.sin(np.arange(n_frames)[:, None, None, None] / 6.0 + rng.uniform(0, 6, (1,) + X.shape[1:]))
    F, A, J, _ = X.shape
    proj = og.project(cams, X.reshape(-1, 3)).reshape(n_cams, F, A, J, 2)
    kp = np.zeros((A, F, n_cams, J, 3))
    kp[..., :2] = proj.transpose(2, 1, 0, 3, 4) + rng.normal(0, 0.6, size=(A, F, n_cams, J, 2))
    kp[..., 2] = rng.uniform(0.55, 1.0, size=(A, F, n_cams, J))
    jump = rng.random((A, F, n_cams, J)) < 0.03
    kp[jump, :2] += rng.normal(0, 40.0, size=(int(jump.sum()), 2))
    low = rng.random((A, F, n_cams, J)) < 0.08
    kp[low, 2] = rng.uniform(0.0, 0.28, size=int(low.sum()))
    ids = [str(i + 1) for i in range(n_cams)]
    intrin, extrin = {}, {}
    for cid, d in zip(ids, dicts):
        mtx = np.array(d["matrix"], dtype=np.float64)
        mtx[:2, :] *= 2                                   # proc halves the first two rows (:119)
        intrin[cid] = {"mtx": mtx, "dist": np.array(d["distortions"], dtype=np.float64).reshape(1, -1),
                       "xi": np.array([[1.1]]), "K": mtx * 0.9, "D": np.array([[0.01, -0.02, 0.0, 0.0]])}
        extrin[cid] = {"rvec": np.array(d["rotation"], dtype=np.float64).reshape(3, 1),
                       "tvec": np.array(d["translation"], dtype=np.float64).reshape(3, 1)}
    if return_tracks:
        return kp, ids, intrin, extrin, X
    return kp, ids, intrin, extrin


def run_variant(step4, name, n_cams, n_animals, n_frames, seed, ransac, optim, fixed_lengths=False):
    import toml
    import yaml
    kp2d, ids, intrin, extrin, X_true = synthetic_recording(n_cams, n_animals, n_frames, seed, return_tracks=True)
    H5_STORE["cam_intrinsic.h5"] = intrin
    H5_STORE["cam_extrinsic_optim.h5"] = extrin
    scratch = tempfile.mkdtemp(prefix="m3d_step4_")
    cwd = os.getcwd()
    try:
        os.makedirs(os.path.join(scratch, "configs"))
        cfg = toml.load(os.path.join(REF, "configs", "config_tmpl.toml"))
        cfg["triangulation"]["ransac"] = bool(ransac)
        cfg["triangulation"]["optim"] = bool(optim)
        toml.dump(cfg, open(os.path.join(scratch, "configs", "config_tmpl.toml"), "w"))
        calib_t = toml.load(os.path.join(REF, "configs", "calibration_tmpl.toml"))
        for k in calib_t:
            if k.startswith("cam_"):
                calib_t[k]["omnidir"] = False             # cv2.omnidir is not installed (see the header)
        toml.dump(calib_t, open(os.path.join(scratch, "configs", "calibration_tmpl.toml"), "w"))
        os.makedirs(os.path.join(scratch, "calib"))
        config_path = os.path.join(scratch, "calib", "config.yaml")
        yaml.safe_dump({"camera_id": [int(i) for i in ids]}, open(config_path, "w"))
        joint_len_in = None
        if fixed_lengths:
            # joint_len.npy as an earlier run of proc leaves it (:273): one row of strong + weak limb lengths per
            # animal; here three rows around the true limb lengths of the synthetic animal (the median is used, :179-181)
            from oracle.make_golden import MACAQUE_CONSTRAINTS, MACAQUE_CONSTRAINTS_WEAK
            rng = np.random.default_rng(seed + 5)
            cons = np.array(MACAQUE_CONSTRAINTS + MACAQUE_CONSTRAINTS_WEAK)
            true_len = np.median(np.linalg.norm(X_true[:, 0, cons[:, 0]] - X_true[:, 0, cons[:, 1]], axis=-1), axis=0)
            joint_len_in = true_len[None] * (1.0 + 0.02 * rng.normal(size=(3, cons.shape[0])))
            np.save(os.path.join(scratch, "calib", "joint_len.npy"), joint_len_in)
        result_root = os.path.join(scratch, "results")
        os.makedirs(os.path.join(result_root, "clip"))
        with open(os.path.join(result_root, "clip", "kp2d.pickle"), "wb") as f:
            pickle.dump(kp2d, f)
        os.chdir(scratch)
        step4.proc("clip", result_root, config_path, kp2d.shape[3], redo=True)
        os.chdir(cwd)
        rd = os.path.join(result_root, "clip")
        kp2d_f = pickle.load(open(os.path.join(rd, "kp2d_f.pickle"), "rb"))
        out_name = "kp3d_fxdJointLen.pickle" if fixed_lengths else "kp3d.pickle"
        kp3d = pickle.load(open(os.path.join(rd, out_name), "rb"))
        save = dict(
            kp2d=kp2d, camera_ids=np.array(ids), ransac=bool(ransac), optim=bool(optim),
            fixed_lengths=bool(fixed_lengths),
            intrin_mtx=np.stack([intrin[i]["mtx"] for i in ids]), intrin_dist=np.stack([intrin[i]["dist"] for i in ids]),
            intrin_xi=np.stack([intrin[i]["xi"] for i in ids]), intrin_K=np.stack([intrin[i]["K"] for i in ids]),
            intrin_D=np.stack([intrin[i]["D"] for i in ids]),
            extrin_rvec=np.stack([extrin[i]["rvec"] for i in ids]), extrin_tvec=np.stack([extrin[i]["tvec"] for i in ids]),
            calibration_toml=np.array(open(os.path.join(rd, "calibration.toml")).read()),
            config_toml=np.array(open(os.path.join(rd, "config.toml")).read()),
            kp2d_f=np.asarray(kp2d_f), kp3d=np.asarray(kp3d["kp3d"]), kp3d_score=np.asarray(kp3d["kp3d_score"]),
            kp3d_err=np.asarray(kp3d["kp3d_err"]),
            joint_len=np.asarray(kp3d["joint_len"], dtype=np.float64) if len(kp3d["joint_len"]) else np.zeros((0,)),
            joint_len_in=joint_len_in if joint_len_in is not None else np.zeros((0,)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print(name, "kp3d", save["kp3d"].shape, "finite", float(np.isfinite(save["kp3d"][..., 0]).mean()),
              "mean err", float(np.nanmean(save["kp3d_err"])))
    finally:
        os.chdir(cwd)
        shutil.rmtree(scratch, ignore_errors=True)


def main():
    step4 = import_reference_step4()
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    if not only or "plain" in only:
        run_variant(step4, "step4_plain", 8, 2, 40, 601, ransac=False, optim=False)
    if not only or "ransac" in only:
        run_variant(step4, "step4_ransac", 8, 2, 40, 602, ransac=True, optim=False)
    if not only or "optim" in only:
        run_variant(step4, "step4_optim", 8, 1, 60, 603, ransac=False, optim=True)      # the template's default
    if only and "fixedlen" not in only:
        return
    run_variant(step4, "step4_optim_fixedlen", 8, 1, 60, 604, ransac=True, optim=True, fixed_lengths=True)


if __name__ == "__main__":
    main()
