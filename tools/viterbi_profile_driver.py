import sys, numpy as np, torch
sys.path.insert(0, ".")
from macaque_3d_pose_estimation_b200 import filter2d, synth
dev = torch.device("cuda", 0)
Sv, Fv = 2176, 20000
det = torch.from_numpy(np.ascontiguousarray(synth.make_detection_series(Fv, 64, 1, 5).transpose(1, 0, 2, 3))).to(dev).repeat(Sv // 64, 1, 1, 1).contiguous()
filter2d.viterbi_series(det, 3, 25.0, 0.3)
torch.cuda.synchronize()
