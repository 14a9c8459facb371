"""BASELINE.json configs[4]: sweep k = 16..64, r = 0.1..1.0 m, map size 1 M..100 M points -- neighbour-search and
fixed-graph step throughput (developer tool, GPU box).  Writes a markdown table.

    python tools/sweep.py [out.md] [max_scans]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.synthetic import make_sequence, make_poses   # noqa: E402


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return best, out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/sweep.md'
    max_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 763
    dev = torch.device('cuda:0')
    peak = 6548.5
    lines = ['# Sweep (BASELINE.json configs[4]): kNN-within-r search and fixed-graph step on corridor maps of full-resolution OS0-128 scans',
             '', 'CUDA-event times, best of 3; step = `dc_step_points` + `dc_step_forward_scatter` (or the two-kernel forms below 2^19 points) + chain on a reused graph '
             '(ScaledPolynomial[2,4], min_eigval_loss(normalization), per-scan pose corrections); roofline fraction = '
             '(206 + 8K) B/point (SURVEY.md section 8(d)) / step time / %.1f GB/s.' % peak, '',
             '| points | k | r [m] | mean valid neighbours | cell [m] | search ms | of which kNN kernel | search Mpts/s | step ms | step Mpts/s | step frac of HBM roofline |', '|---|---|---|---|---|---|---|---|---|---|---|']
    for n_scans in (8, 77, 763):
        if n_scans > max_scans:
            continue
        t0 = time.time()
        scans, _, _ = make_sequence('corridor', n_scans=n_scans, pattern='os0-128', seed=0)
        poses = torch.as_tensor(make_poses('corridor', n_scans), device=dev)
        cfg0 = dc.Config(nn_k=32, nn_r=0.4, pose_correction=dc.PoseCorrection.pose)
        clouds = []
        for s in scans:
            c = dc.DepthCloud.from_points(torch.from_numpy(s['points']).to(dev))
            c.inc_angles = torch.rand((len(c), 1), device=dev)        # constants of the step; their values do not matter for speed
            clouds.append(c)
        del scans
        n = sum(len(c) for c in clouds)
        print('N = %d (%d scans) generated in %.1f s' % (n, n_scans, time.time() - t0), flush=True)
        deltas = torch.zeros((n_scans, 6), dtype=torch.float64, device=dev, requires_grad=True)
        model = dc.ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=dev)
        for k in (16, 32, 64):
            for r in (0.1, 0.2, 0.4, 1.0):
                if n_scans == 763 and not (k == 32 and r == 0.4) and not (k == 64 and r == 1.0) and not (k == 16 and r == 0.1):
                    continue        # three corners at 100 M points
                cfg = dc.Config(nn_k=k, nn_r=r, pose_correction=dc.PoseCorrection.pose)
                ms_search, ns = timed(lambda: dc.establish_neighborhoods(clouds=clouds, poses=poses, cfg=cfg))
                deg = ns.graph.degrees().double().mean().item()
                L.profile = {}
                ns = dc.establish_neighborhoods(clouds=clouds, poses=poses, cfg=cfg)
                torch.cuda.synchronize()
                ms_knn = sum(v['ms_total'] for kk, v in L.collect_profile().items() if kk.startswith('dc_knn'))
                L.profile = None

                def step():
                    model.zero_grad(set_to_none=True)
                    deltas.grad = None
                    pc = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
                    feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
                    loss, _ = dc.min_eigval_loss(feats, normalization=True)
                    loss.backward()
                    return loss
                step()
                ms_step, _ = timed(step, reps=4)
                frac = (206 + 8 * k) * n / (ms_step * 1e-3) / 1e9 / peak
                lines.append('| %d | %d | %.1f | %.1f | %.4f | %.2f | %.2f | %.0f | %.2f | %.0f | %.3f |'
                             % (n, k, r, deg, ns.graph.map.cell, ms_search, ms_knn, n / ms_search / 1e3, ms_step, n / ms_step / 1e3, frac))
                print(lines[-1], flush=True)
                del ns
                L.release_workspace()
                torch.cuda.empty_cache()
        del clouds
        torch.cuda.empty_cache()
    open(out_path, 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
