#!/usr/bin/env python
"""Developer tool: one launch of the step-4 Viterbi filter on 2176 series x 20,000 frames (the
workload of the k_viterbi line in tools/kernel_bench.py) — the target of the ncu capture under
profiles/; with --time it also prints CUDA-event timings of repeated calls."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaque_3d_pose_estimation_b200 import filter2d, synth  # noqa: E402

dev = torch.device("cuda", 0)
Sv, Fv = 2176, 20000
det = torch.from_numpy(np.ascontiguousarray(synth.make_detection_series(Fv, 64, 1, 5).transpose(1, 0, 2, 3))).to(dev)
det = det.repeat(Sv // 64, 1, 1, 1).contiguous()
filter2d.viterbi_series(det, 3, 25.0, 0.3)
torch.cuda.synchronize()
if "--time" in sys.argv:
    for reps in (1, 1, 3, 10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            filter2d.viterbi_series(det, 3, 25.0, 0.3)
        e1.record()
        torch.cuda.synchronize()
        print("reps", reps, "ms per call", e0.elapsed_time(e1) / reps)
