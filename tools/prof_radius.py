"""Radius-mode search + fixed-graph step on voxel-filtered scans (the reference's default setting: grid_res 0.1-0.2 m,
nn_r 0.25-0.4 m, nn_k 0).  Developer tool, GPU box."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.synthetic import make_sequence, make_poses   # noqa: E402


def main():
    dev = torch.device('cuda:0')
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    scans, _, _ = make_sequence('corridor', n_scans=n_scans, pattern='os0-128', seed=0)
    poses = torch.as_tensor(make_poses('corridor', n_scans), device=dev)
    for grid_res, r in ((0.1, 0.25), (0.1, 0.4), (0.05, 0.15), (0.2, 0.4)):
        cfg = dc.Config(nn_k=0, nn_r=r, grid_res=grid_res, min_depth=1.0, max_depth=25.0, pose_correction=dc.PoseCorrection.pose)
        t0 = time.perf_counter()
        clouds = []
        for s in scans:
            c = dc.filtered_cloud(dc.DepthCloud.from_points(torch.from_numpy(s['points']).to(dev)), cfg)
            c.inc_angles = torch.rand((len(c), 1), device=dev)
            clouds.append(c)
        torch.cuda.synchronize()
        t_filter = (time.perf_counter() - t0) * 1e3
        n = sum(len(c) for c in clouds)
        deltas = torch.zeros((n_scans, 6), dtype=torch.float64, device=dev, requires_grad=True)
        model = dc.ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=dev)
        best_s = best_f = 1e9
        for rep in range(3):
            L.profile = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ns = dc.establish_neighborhoods(clouds=clouds, poses=poses, cfg=cfg)
            torch.cuda.synchronize()
            best_s = min(best_s, (time.perf_counter() - t0) * 1e3)
            prof_s = L.collect_profile()
            L.profile = None

            def step():
                model.zero_grad(set_to_none=True)
                deltas.grad = None
                pc = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
                feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
                loss, _ = dc.min_eigval_loss(feats, normalization=True)
                loss.backward()
            step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            torch.cuda.synchronize()
            best_f = min(best_f, (time.perf_counter() - t0) * 1e3)
        deg = ns.graph.degrees().double()
        print('grid %.2f r %.2f: %d points (filter_grid of %d scans %.1f ms), neighbours mean %.1f max %d; search %.2f ms (%s), step %.2f ms'
              % (grid_res, r, n, n_scans, t_filter, deg.mean().item(), int(deg.max().item()), best_s,
                 ', '.join('%s %.2f' % (k, v['ms_total']) for k, v in sorted(prof_s.items(), key=lambda kv: -kv[1]['ms_total'])[:3]), best_f),
              flush=True)


if __name__ == '__main__':
    main()
