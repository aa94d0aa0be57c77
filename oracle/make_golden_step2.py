#!/usr/bin/env python
"""TEST INFRASTRUCTURE (not part of the product): golden vectors of the keyframe association, produced by
executing the unmodified reference method ``MultiEstimator.predict_data``
(/root/reference/src/pipeline/step2_crossviewmatching.py:502-713) frame by frame on a synthetic recording.

``predict_data`` reaches ``cv2.omnidir.undistortPoints`` (through ``calc_3dpose`` -> ``mct.undistortPoints``,
multicam_toolbox.py:404-420) and ``cv2.omnidir.projectPoints`` (``reproject``, step2:465-489).  opencv-contrib is
not installed, so the run installs a stand-in ``cv2.omnidir`` whose two functions are the oracle's restatement of
the Mei model (oracle/camera_math.py, PARITY UNPINNED).  What the goldens therefore pin by execution is everything
ELSE of the method: the affinity / identity weighting, matchSVT, the cluster extraction, ``get_best_comb`` with
its combination order and first-minimum rule, the leftover round, the ``>= 2`` filter, the order of the returned
persons and the ``bcomb`` bookkeeping.  The omnidir arithmetic itself stays a restatement on both sides.

Writes ``tests/golden/predict_data_<variant>.npz``: inputs (rig, kp_raw (F,M,J,3), dim (F,C+1), cid, bbox id) and
the reference's return values flattened over frames (person_frame, person_members (P,C), p3d (P,J,3), bcomb (P,C)).
Run from the repo root: ``python oracle/make_golden_step2.py``.
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from macaque_3d_pose_estimation_b200 import synth  # noqa: E402
from oracle import camera_math as cm  # noqa: E402
from oracle.make_golden import rig_arrays  # noqa: E402


class _OmnidirStandIn:
    """The two cv2.omnidir entry points predict_data reaches, with OpenCV's argument order and array shapes."""

    @staticmethod
    def undistortPoints(distorted, K, D, xi, R, *a, **k):
        pts = np.asarray(distorted, dtype=np.float64)
        out = cm.undistort_omnidir(pts.reshape(-1, 2), K, np.ravel(D), float(np.ravel(xi)[0]))
        return np.asarray(out, dtype=np.float64).reshape(pts.shape)

    @staticmethod
    def projectPoints(obj, rvec, tvec, K, xi, D, *a, **k):
        p = np.asarray(obj, dtype=np.float64).reshape(-1, 3)
        out = cm.project_omnidir(p, np.ravel(rvec), np.ravel(tvec), K, xi, np.ravel(D))
        return np.asarray(out, dtype=np.float64).reshape(-1, 1, 2), None


def import_reference_step2():
    sys.dont_write_bytecode = True
    for name in ("h5py", "imgstore", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.path.insert(0, os.path.join(REF, "src", "third_party"))
    import cv2
    assert not hasattr(cv2, "omnidir"), "opencv-contrib is installed: use the real cv2.omnidir and drop the stand-in"
    cv2.omnidir = _OmnidirStandIn
    from src.pipeline import step2_crossviewmatching as s2
    return s2, cv2


def case_predict(s2, cv2, name, n_frames, seed, dup=0.08, drop=0.12, noise=0.4):
    C, A, J = 8, 6, 17
    dicts = synth.make_rig(C, "omnidir", seed=seed)
    rng = np.random.default_rng(seed)
    camparam = {"camera_id": [d["name"] for d in dicts], "K": [], "xi": [], "D": [], "rvecs": [], "tvecs": [], "pmat": []}
    for d in dicts:
        R, _ = cv2.Rodrigues(np.array(d["rotation"], dtype=np.float64))
        t = np.array(d["translation"], dtype=np.float64).reshape(3, 1)
        camparam["K"].append(np.array(d["K"], dtype=np.float64))
        camparam["xi"].append(np.array(d["xi"], dtype=np.float64).reshape(1, 1))
        camparam["D"].append(np.array(d["D"], dtype=np.float64).reshape(1, 4))
        camparam["rvecs"].append(np.array(d["rotation"], dtype=np.float64).reshape(3, 1))
        camparam["tvecs"].append(t)
        camparam["pmat"].append(np.hstack([R, t]))
    X = synth.make_tracks(n_frames, A, seed=seed) * np.array([0.6, 0.6, 0.5])
    M = C * (A + 2)
    kp_raw = np.zeros((n_frames, M, J, 3))
    dim = np.zeros((n_frames, C + 1), dtype=np.int32)
    cid = -np.ones((n_frames, M), dtype=np.int32)
    bbox = -np.ones((n_frames, M), dtype=np.int64)
    owner = -np.ones((n_frames, M), dtype=np.int64)
    est = s2.MultiEstimator("")
    pf, pm, pp, pb = [], [], [], []
    t_all = 0.0
    for f in range(n_frames):
        m = 0
        info = {}
        for c in range(C):
            dets = []
            for a in rng.permutation(A):
                if rng.random() < drop:
                    continue
                reps = 2 if rng.random() < dup else 1          # the detector fires twice on one animal
                for r in range(reps):
                    if m >= dim[f, c] + A + 2:
                        break
                    raw = cm.project_omnidir(X[f, a], np.ravel(camparam["rvecs"][c]), np.ravel(camparam["tvecs"][c]),
                                             camparam["K"][c], camparam["xi"][c], np.ravel(camparam["D"][c]))
                    raw = raw + rng.normal(0, noise * (1 + 4 * r), size=(J, 2))
                    sc = rng.uniform(0.3, 1.0, size=J)
                    sc[rng.random(J) < 0.1] = 0.0
                    kp_raw[f, m] = np.concatenate([raw, sc[:, None]], axis=1)
                    cid[f, m] = a if rng.random() < 0.6 else -1
                    bbox[f, m] = 100 * c + m
                    owner[f, m] = a
                    # pose2d: what step2.undistort_points (:306-325) hands to predict_data
                    und = cv2.omnidir.undistortPoints(kp_raw[f, m, :, :2].reshape(-1, 1, 2), camparam["K"][c],
                                                      camparam["D"][c], camparam["xi"][c], np.eye(3)).reshape(-1, 2)
                    dets.append({"pose2d": und, "pose2d_raw": kp_raw[f, m].copy(), "bbox": [0, 0, 1, 1],
                                 "bbox_id": (c, int(bbox[f, m])), "cid": int(cid[f, m])})
                    m += 1
            dim[f, c + 1] = m
            info[c] = [dets]
        t0 = time.time()
        matched, p3d, bcomb = est.predict_data(info, show=False, camparam=camparam)
        t_all += time.time() - t0
        sub2cam = np.searchsorted(dim[f], np.arange(m), side="right") - 1
        for mem, P, b in zip(matched, p3d, bcomb):
            row = -np.ones(C, dtype=np.int64)
            for s in np.asarray(mem, dtype=int):
                assert row[sub2cam[s]] < 0
                row[sub2cam[s]] = s
            pf.append(f)
            pm.append(row)
            pp.append(np.asarray(P, dtype=np.float64))
            pb.append(np.asarray(b, dtype=np.int64))
        print(name, "frame", f, "M=%d persons=%d" % (m, len(matched)), flush=True)
    arrs = rig_arrays(dicts)
    arrs.update(kp_raw=kp_raw, dim=dim, cid=cid, bbox=bbox, owner=owner, person_frame=np.array(pf, dtype=np.int64),
                person_members=np.array(pm, dtype=np.int64).reshape(-1, C), p3d=np.array(pp).reshape(-1, J, 3),
                bcomb=np.array(pb, dtype=np.int64).reshape(-1, C), ref_seconds=np.array(t_all))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(name, "persons", len(pf), "reference %.1f s" % t_all)


def case_step3(cv2, name, n_frames, seed):
    """step3_crossframematching.py calc_3dpose (:254-272, score gate 0.3), calc_3dtrace (:274-302) and
    calc_dist_pose (:304-311) executed on two synthetic tracklets in the reference's own containers:
    T[i_cam][i_frame] = list of 2D tracks (entry[0] = bbox id, entry[5] = (J,3) keypoints), trk[i_frame][i_cam] =
    the bbox id the tracklet uses there (-1 = none)."""
    from src.pipeline import step3_crossframematching as s3
    C, A, J = 8, 2, 17
    dicts = synth.make_rig(C, "omnidir", seed=seed)
    rng = np.random.default_rng(seed)
    camparam = {"camera_id": [d["name"] for d in dicts], "K": [], "xi": [], "D": [], "rvecs": [], "tvecs": [], "pmat": []}
    for d in dicts:
        R, _ = cv2.Rodrigues(np.array(d["rotation"], dtype=np.float64))
        t = np.array(d["translation"], dtype=np.float64).reshape(3, 1)
        camparam["K"].append(np.array(d["K"], dtype=np.float64))
        camparam["xi"].append(np.array(d["xi"], dtype=np.float64).reshape(1, 1))
        camparam["D"].append(np.array(d["D"], dtype=np.float64).reshape(1, 4))
        camparam["rvecs"].append(np.array(d["rotation"], dtype=np.float64).reshape(3, 1))
        camparam["tvecs"].append(t)
        camparam["pmat"].append(np.hstack([R, t]))
    X = synth.make_tracks(n_frames, A, seed=seed) * np.array([0.6, 0.6, 0.5])
    X[:, 1] = X[:, 0] + rng.normal(0, 60.0, size=(1, 1, 3)) + rng.normal(0, 8.0, size=X[:, 0].shape)   # a nearby second animal
    kp = np.full((A, n_frames, C, J, 3), np.nan)
    trk = -np.ones((A, n_frames, C), dtype=np.int64)
    T = [[[] for _ in range(n_frames)] for _ in range(C)]
    for f in range(n_frames):
        for c in range(C):
            for a in range(A):
                if rng.random() < 0.25:
                    continue                                   # this camera has no track of the animal in this frame
                raw = cm.project_omnidir(X[f, a], np.ravel(camparam["rvecs"][c]), np.ravel(camparam["tvecs"][c]),
                                         camparam["K"][c], camparam["xi"][c], np.ravel(camparam["D"][c]))
                raw = raw + rng.normal(0, 0.5, size=(J, 2))
                sc = rng.uniform(0.2, 1.0, size=J)              # some keypoints under the 0.3 gate
                k3 = np.concatenate([raw, sc[:, None]], axis=1)
                if rng.random() < 0.05:
                    k3[rng.integers(J), :2] = np.nan            # a missing keypoint
                bid = 10 * f + a
                T[c][f].append([bid, 0, 0, 0, 0, k3.tolist()])
                kp[a, f, c] = k3
                trk[a, f, c] = bid
    frames = np.arange(2, n_frames - 1)
    traces = [s3.calc_3dtrace(trk[a], T, frames, camparam, "", J) for a in range(A)]
    rmse = s3.calc_dist_pose(traces[0], traces[1])
    poses = np.stack([s3.calc_3dpose(np.nan_to_num(kp[0, f], nan=np.nan), "", camparam) for f in range(4)])
    arrs = rig_arrays(dicts)
    arrs.update(kp=kp, trk=trk, frames=frames, trace0=traces[0], trace1=traces[1], rmse=np.array(rmse), poses=poses)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(name, "rmse %.3f" % rmse, "finite trace rows", int(np.isfinite(traces[0][:, 0]).sum()))


def main():
    s2, cv2 = import_reference_step2()
    os.makedirs(OUT, exist_ok=True)
    if "step3" in sys.argv[1:]:
        case_step3(cv2, "step3_traces", 24, 821)
        return
    case_predict(s2, cv2, "predict_data_dups", 16, 811, dup=0.08, drop=0.12, noise=0.4)
    case_predict(s2, cv2, "predict_data_clean", 10, 812, dup=0.0, drop=0.0, noise=0.15)
    case_step3(cv2, "step3_traces", 24, 821)


if __name__ == "__main__":
    main()
