#!/usr/bin/env python
"""Developer tool: pruned search (pair certificates, k_cert_*) against the exhaustive kernels on the
GPU at >= 1e7 joint-instances per case, and against the numpy oracle on a prefix of every case.
A single mismatch in subset index / picked cameras / search length is a failure of the certificate.
Usage: python tools/stress_cert.py [frames_per_case] [oracle_points] -> gpurun_out/stress_cert.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT))
import bench  # noqa: E402  (the device workload generator)
from macaque_3d_pose_estimation_b200 import _lib, synth  # noqa: E402
from macaque_3d_pose_estimation_b200.cameras import CameraGroup  # noqa: E402
from oracle import cameragroup as og  # noqa: E402
from oracle import fixtures  # noqa: E402


def run(cg, pts, mode, min_cams):
    lib = _lib.load()
    rig = cg._rig()
    _lib.check(lib.m3d_rig_set_ransac_mode(rig.handle, mode))
    r = cg.triangulate_ransac(pts, min_cams=min_cams, return_stats=True)
    _lib.check(lib.m3d_rig_set_ransac_mode(rig.handle, 0))
    return r


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
    n_or = int(sys.argv[2]) if len(sys.argv) > 2 else 6800
    dev = torch.device("cuda", 0)
    out = []
    cases = [(20261020, "bench", 2), (20261020, "bench", 3), (5, "bench", 2), (7, "wide", 2), (9, "mild", 2)]
    for seed, kind, mc in cases:
        dicts = synth.make_rig(8, "pinhole", seed=seed)
        cg = CameraGroup.from_dicts(dicts)
        cg.device = 0
        lib = _lib.load()
        mask = lib.m3d_rig_certified_mask(cg._rig().handle)
        if kind == "bench":
            xy = bench.make_device_workload(cg, frames, 4, 17, seed, "ransac", dev)
        else:
            # "wide": animals all over the room, many views off-image are kept (wild distortion region);
            # "mild": outliers of a few pixels, the hard case for the certificates
            g = torch.Generator(device=dev)
            g.manual_seed(seed)
            X = (torch.rand((frames * 68, 3), generator=g, device=dev, dtype=torch.float64) - 0.5) * \
                torch.tensor([2400.0, 2400.0, 1600.0], device=dev, dtype=torch.float64) + \
                torch.tensor([0.0, 0.0, 700.0], device=dev, dtype=torch.float64)
            xy = cg.project(X)
            del X
            for c in range(8):
                n = xy.shape[1]
                xy[c] += torch.randn((n, 2), generator=g, device=dev, dtype=torch.float64) * 0.3
                o = torch.rand((n,), generator=g, device=dev) < 0.25
                sig = 60.0 if kind == "wide" else 4.0
                xy[c] += o[:, None] * torch.randn((n, 2), generator=g, device=dev, dtype=torch.float64) * sig
                xy[c][torch.rand((n,), generator=g, device=dev) < 0.1] = float("nan")
        a = run(cg, xy, 0, mc)
        b = run(cg, xy, 1, mc)
        rec = {"rig_seed": seed, "case": kind, "min_cams": mc, "points": int(xy.shape[1]),
               "certified_mask": int(mask),
               "subset_index_mismatches": int((a[4] != b[4]).sum().item()),
               "search_length_mismatches": int((a[5] != b[5]).sum().item()),
               "picked_mismatches": int((a[1] != b[1]).any(dim=0).sum().item()),
               "max_abs_err_px": float((a[3] - b[3]).abs().max().item()),
               "max_abs_p3d_mm": float((a[0] - b[0]).nan_to_num().abs().max().item()),
               "mean_subsets": float(a[5].double().mean().item()),
               "selected_fraction": float((a[4] >= 0).double().mean().item())}
        p2 = xy[:, :n_or].cpu().numpy()
        o = og.triangulate_ransac(fixtures.cams_from_dicts(dicts), p2, min_cams=mc, return_stats=True)
        rec["oracle_points"] = int(n_or)
        rec["oracle_subset_mismatches"] = int((o[4] != a[4][:n_or].cpu().numpy()).sum())
        rec["oracle_search_length_mismatches"] = int((o[5] != a[5][:n_or].cpu().numpy()).sum())
        rec["oracle_max_abs_err_px"] = float(np.abs(o[3] - a[3][:n_or].cpu().numpy()).max())
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del xy, a, b
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "stress_cert.json"), "w"), indent=1)
    bad = sum(r["subset_index_mismatches"] + r["search_length_mismatches"] + r["picked_mismatches"] +
              r["oracle_subset_mismatches"] + r["oracle_search_length_mismatches"] for r in out)
    print("TOTAL MISMATCHES", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
