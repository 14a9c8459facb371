"""cProfile of bench.one_step on the host side (developer tool; run on a GPU box)."""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from bench import host_scans, local_features, one_step, NN_K, NN_R   # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(n_scans, 'os0-128')
cfg = dc.Config(nn_k=NN_K, nn_r=NN_R, pose_correction=dc.PoseCorrection.pose)
clouds = local_features(dc, [torch.from_numpy(p).to(dev) for p in pts_host], cfg)
poses = torch.as_tensor(poses_np, device=dev)
deltas = torch.zeros((n_scans, 6), dtype=torch.float64, device=dev, requires_grad=True)
model = dc.ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=dev)
for _ in range(3):
    one_step(dc, clouds, poses, deltas, model, cfg)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    one_step(dc, clouds, poses, deltas, model, cfg)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
