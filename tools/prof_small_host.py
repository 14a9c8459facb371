"""Host-side profile (cProfile) of the fixed-graph training iteration on a SMALL map (BASELINE.json configs[0] / [3]: the
regime of the reference's own runs), where the step is bound by Python and launch overhead, not by the kernels.

    python tools/prof_small_host.py
"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from bench import make_sequence                         # noqa: E402

dev = torch.device('cuda:0')
cfg = dc.Config(min_depth=1.0, max_depth=25.0, grid_res=0.1, nn_k=0, nn_r=0.25, loss='trace_loss', pose_correction=dc.PoseCorrection.pose)
scans_np, _, poses_init = make_sequence('fee', n_scans=12, pattern='os0-128', seed=5, pose_noise=(0.01, 0.005), bias_w=[-0.01], bias_exponent=[4.0])
clouds = dc.local_feature_clouds([dc.filtered_cloud(dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev)), cfg) for s in scans_np], cfg)
poses = torch.as_tensor(poses_init, device=dev)
model = dc.ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=dev)
deltas = torch.zeros((len(clouds), 6), dtype=torch.float64, device=dev, requires_grad=True)
opt = torch.optim.Adam([{'params': deltas, 'lr': 1e-3}, {'params': model.parameters(), 'lr': 1e-3}])
ns = dc.establish_neighborhoods(clouds=clouds, poses=poses, cfg=cfg)


def it():
    pc = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
    feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
    loss, _ = dc.trace_loss(feats, sqrt=False)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(10):
    it()
torch.cuda.synchronize()
n0 = L.launch_count
t0 = time.perf_counter()
for _ in range(100):
    it()
host = (time.perf_counter() - t0) / 100 * 1e3
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 100 * 1e3
print('%d points: host %.3f ms / iteration, wall %.3f ms, %d library launches / iteration' % (sum(len(c) for c in clouds), host, wall, (L.launch_count - n0) // 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    it()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(40)
