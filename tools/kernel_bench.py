#!/usr/bin/env python
"""Developer tool: device-resident timing of every secondary C-ABI entry point (CUDA events,
3 warm-ups, inputs larger than L2) -> gpurun_out/kernel_bench.json.  The headline kernels are
timed by bench.py; this fills the remaining rows of the DESIGN.md kernel table."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from macaque_3d_pose_estimation_b200 import crossview as cv  # noqa: E402
from macaque_3d_pose_estimation_b200 import synth  # noqa: E402
from macaque_3d_pose_estimation_b200.cameras import CameraGroup  # noqa: E402

HBM = 6542.1


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    C, N = 8, 20_000_000
    dicts = synth.make_rig(C, "pinhole", seed=11)
    cg = CameraGroup.from_dicts(dicts)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    X = (torch.rand((N, 3), generator=g, device=dev, dtype=torch.float64) - 0.5) * torch.tensor(
        [1200.0, 1200.0, 800.0], device=dev, dtype=torch.float64) + torch.tensor([0.0, 0.0, 400.0], device=dev, dtype=torch.float64)
    xy = cg.project(X)
    xy += torch.randn(xy.shape, generator=g, device=dev, dtype=torch.float64) * 0.3
    und = cg.undistort_points(xy)
    p3d = cg.triangulate(xy)
    out = {}

    def rec(name, ms, bytes_per_unit, units, unit):
        out[name] = {"ms": ms, "units_per_s": units / (ms * 1e-3), "unit": unit,
                     "algorithmic_bytes_per_unit": bytes_per_unit,
                     "GBps": bytes_per_unit * units / (ms * 1e-3) / 1e9,
                     "frac_of_measured_hbm": bytes_per_unit * units / (ms * 1e-3) / 1e9 / HBM}
        print(name, json.dumps(out[name]), flush=True)

    rec("k_undistort (8 cams)", timeit(lambda: cg.undistort_points(xy)), 32, C * N, "observations")
    rec("k_project (8 cams)", timeit(lambda: cg.project(X)), 24 + 16 * C, N, "joint-instances")
    rec("k_reproj mean", timeit(lambda: cg.reprojection_error(p3d, xy, mean=True)), 24 + 16 * C + 8, N, "joint-instances")
    rec("k_reproj full", timeit(lambda: cg.reprojection_error(p3d, xy, mean=False)), 24 + 32 * C, N, "joint-instances")
    rec("k_triangulate (no undistort, no err)", timeit(lambda: cg.triangulate(und, undistort=False)), 16 * C + 24, N, "joint-instances")
    rec("k_triangulate (undistort, no err)", timeit(lambda: cg.triangulate(xy)), 16 * C + 24, N, "joint-instances")
    use = torch.ones((C, N), dtype=torch.uint8, device=dev)
    rec("k_triangulate_ls", timeit(lambda: cv.triangulate_ls_batch(cg, und, use)), 17 * C + 24, N, "joint-instances")
    del xy, und, p3d, X, use
    torch.cuda.empty_cache()

    # cross-view association, BASELINE config 4 shape: 6 animals x 8 views, M = 48, J = 17
    F, A, J = 20000, 6, 17
    M = A * C
    Xf = synth.make_tracks(200, A, seed=3)                       # (200, A, J, 3) tiled to F frames
    Xf = torch.from_numpy(np.tile(Xf, (F // 200, 1, 1, 1))).to(dev)
    M4 = torch.from_numpy(cg.get_extrinsics_mats()).to(dev)      # (C,4,4)
    Xc = torch.einsum("cij,fakj->fcaki", M4[:, :3, :3], Xf) + M4[:, :3, 3][None, :, None, None, :]
    kp = torch.empty((F, C, A, J, 3), dtype=torch.float64, device=dev)
    kp[..., 0] = Xc[..., 0] / Xc[..., 2] + torch.randn((F, C, A, J), generator=g, device=dev, dtype=torch.float64) * 4e-4
    kp[..., 1] = Xc[..., 1] / Xc[..., 2] + torch.randn((F, C, A, J), generator=g, device=dev, dtype=torch.float64) * 4e-4
    kp[..., 2] = 0.3 + 0.7 * torch.rand((F, C, A, J), generator=g, device=dev, dtype=torch.float64)
    kp = kp.reshape(F, M, J, 3).contiguous()
    dim = torch.arange(0, M + 1, A, dtype=torch.int32, device=dev)[None].repeat(F, 1).contiguous()
    aff = cv.geometry_affinity_batch(cg, kp, dim)
    rec("k_ray_affinity (M=48, J=17)", timeit(lambda: cv.geometry_affinity_batch(cg, kp, dim)), 24 * M * J + 8 * M * M, F, "frames")
    W = (0.8 * aff * (aff > 0)).nan_to_num()
    match, its = cv.match_svt_batch(W, dim, C, alpha=0.5, _lambda=50.0, return_iters=True)
    rec("k_match_svt (M=48)", timeit(lambda: cv.match_svt_batch(W, dim, C, alpha=0.5, _lambda=50.0), reps=2), 9 * M * M, F, "frames")
    out["k_match_svt (M=48)"]["mean_admm_iterations"] = float(its.double().mean().item())
    # association quality on clean synthetic data: every animal's 8 detections form one cluster
    blocks = match.reshape(F, C, A, C, A)
    diag = torch.diagonal(blocks, dim1=2, dim2=4)               # same animal across cameras
    out["k_match_svt (M=48)"]["same_animal_link_rate"] = float(diag.double().mean().item())
    # 2D Viterbi filter (step-4 stage): 8 animals x 16 cameras x 17 joints = 2176 series
    from macaque_3d_pose_estimation_b200 import filter2d
    Sv, Fv = 2176, 20000
    det = torch.from_numpy(np.ascontiguousarray(
        synth.make_detection_series(Fv, 64, 1, 5).transpose(1, 0, 2, 3))).to(dev).repeat(Sv // 64, 1, 1, 1).contiguous()
    rec("k_viterbi (P=1, n_back=3, 2176 series)",
        timeit(lambda: filter2d.viterbi_series(det, 3, 25.0, 0.3), reps=3), 24 + 24, Sv * Fv, "series-frames")
    if "--only-viterbi" in sys.argv:
        out = {k: v for k, v in out.items() if "viterbi" in k}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
