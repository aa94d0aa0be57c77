#!/usr/bin/env python
"""Developer tool: the cfg-4 line of bench.py alone (crossview.associate_batch, stage timings).
Usage: python tools/cfg4_bench.py [keyframes]"""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402
from macaque_3d_pose_estimation_b200 import synth  # noqa: E402
from macaque_3d_pose_estimation_b200.cameras import CameraGroup  # noqa: E402

ge.build_library()
args = types.SimpleNamespace(cameras=8, cfg4_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 100000, no_cpu=True)
cx = types.SimpleNamespace(cg=CameraGroup.from_dicts(synth.make_rig(8, "pinhole", seed=20261018 + 2)),
                           device=torch.device("cuda", 0))
print(json.dumps(bench.cfg4_line(cx, args)))
