#!/usr/bin/env python
"""Developer tool: large randomized RANSAC parity run (GPU vs the numpy oracle on EVERY point).
Usage: python tools/stress_parity.py [n_frames]  -> gpurun_out/stress_parity.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from macaque_3d_pose_estimation_b200 import synth  # noqa: E402
from macaque_3d_pose_estimation_b200.cameras import CameraGroup  # noqa: E402
from oracle import cameragroup as og  # noqa: E402
from oracle import fixtures  # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    out = []
    for seed, model, kw, mc in ((901, "pinhole", dict(p_outlier=0.2, p_missing=0.1), 2),
                                (902, "pinhole", dict(noise=0.45, p_outlier=0.25, p_missing=0.15), 3),
                                (903, "fisheye", dict(p_outlier=0.2, p_missing=0.1), 2)):
        dicts = synth.make_rig(8, model, seed=seed)
        cams = fixtures.cams_from_dicts(dicts)
        cg = CameraGroup.from_dicts(dicts)
        X = synth.make_tracks(n_frames, 4, seed=seed).reshape(-1, 3) * np.array([0.6, 0.6, 0.5])
        p2 = synth.corrupt(og.project(cams, X), seed=seed, **kw)
        t0 = time.time()
        h = cg.triangulate_ransac(p2, min_cams=mc, return_stats=True)
        t_gpu = time.time() - t0
        t0 = time.time()
        o = og.triangulate_ransac(cams, p2, min_cams=mc, return_stats=True)
        t_cpu = time.time() - t0
        sel = o[4] >= 0
        rec = {"model": model, "min_cams": mc, "points": int(p2.shape[1]), **{k: float(v) for k, v in kw.items()},
               "picked_mismatches": int((o[1] != h[1]).any(axis=(0, 2)).sum()),
               "subset_index_mismatches": int((o[4] != h[4]).sum()),
               "search_length_mismatches": int((o[5] != h[5]).sum()),
               "max_abs_p3d_mm": float(np.nanmax(np.abs(o[0] - h[0]))),
               "max_abs_err_px": float(np.abs(o[3] - h[3]).max()),
               "min_gap_to_threshold_px": float(np.abs(o[3][sel] - 0.5).min()),
               "mean_subsets": float(o[5].mean()), "oracle_s": t_cpu, "gpu_call_s": t_gpu}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "stress_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
