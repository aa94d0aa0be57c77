#!/usr/bin/env python
"""Developer tool: ADMM iterations and Jacobi sweeps per keyframe of k_match_svt on the cfg-4 workload.
Run once as is (iterations, time), once with M3D_NVCC_EXTRA=-DM3D_DEBUG_SWEEPS after a forced rebuild
(`iters` then carries the sweep count)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build_library(force=bool(os.environ.get("M3D_NVCC_EXTRA")))
from macaque_3d_pose_estimation_b200 import crossview as cv, synth, _lib  # noqa: E402
from macaque_3d_pose_estimation_b200.cameras import CameraGroup  # noqa: E402
import ctypes  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
C, A, J = 8, 6, 17
cg = CameraGroup.from_dicts(synth.make_rig(C, "pinhole", seed=20261018 + 2))
rng = np.random.default_rng(404)
X = synth.make_tracks(F, A, seed=404) * np.array([0.6, 0.6, 0.5])
proj = cg.project(X.reshape(-1, 3)).reshape(C, F, A, J, 2)
M = C * A
kp = np.empty((F, M, J, 3))
kp[..., :2] = proj.transpose(1, 0, 2, 3, 4).reshape(F, M, J, 2) + rng.normal(0, 0.4, size=(F, M, J, 2))
sc = rng.uniform(0.3, 1.0, size=(F, M, J))
sc[rng.random((F, M, J)) < 0.1] = 0.0
kp[..., 2] = sc
dim = np.tile(np.arange(C + 1, dtype=np.int32) * A, (F, 1))
owner = np.tile(np.tile(np.arange(A), C), (F, 1))
cid = np.where(rng.random((F, M)) < 0.6, owner, -1).astype(np.int32)
lib = _lib.require_gpu()
dev = "cuda:0"
rig = cg._rig(0)
d_raw = torch.from_numpy(kp).to(dev)
d_dim = torch.from_numpy(dim).to(dev)
d_und = torch.empty_like(d_raw)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
_lib.check(lib.m3d_undistort_detections(rig.handle, vp(d_raw), vp(d_dim), F, M, J, vp(d_und), None), "und")
aff = cv.geometry_affinity_batch(cg, torch.nan_to_num(d_und), d_dim, 0.1)
W = torch.empty_like(aff)
d_cid = torch.from_numpy(cid).to(dev)
_lib.check(lib.m3d_association_weights(vp(aff), vp(d_cid), vp(d_dim), F, M, C, 0.2, vp(W), 0, None), "w")
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    match, its = cv.match_svt_batch(W, d_dim, C, alpha=0.5, _lambda=50.0, return_iters=True, device=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
its = its.cpu().numpy()
what = "sweeps" if "DEBUG_SWEEPS" in os.environ.get("M3D_NVCC_EXTRA", "") else "iterations"
print({"frames": F, "seconds": dt, "frames_per_s": F / dt, what + "_mean": float(its.mean()),
       what + "_max": int(its.max()), what + "_p50": float(np.median(its)), "match_sum": int(match.sum().item())})
