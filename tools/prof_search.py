"""Wall-clock breakdown of the search phase (developer tool; run on a GPU box)."""
import sys
import time

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import SortedMap, search, _knn_cell_size   # noqa: E402
from bench import host_scans, local_features, NN_K, NN_R   # noqa: E402


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print('%-34s %8.2f ms' % (label, (t1 - t0) * 1e3))
    return time.perf_counter()


def main():
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device('cuda:0')
    pts_host, poses_np = host_scans(n_scans, 'os0-128')
    cfg = dc.Config(nn_k=NN_K, nn_r=NN_R, pose_correction=dc.PoseCorrection.pose)
    clouds = local_features(dc, [torch.from_numpy(p).to(dev) for p in pts_host], cfg)
    poses = torch.as_tensor(poses_np, device=dev)
    for rep in range(3):
        print('--- rep', rep)
        t = time.perf_counter()
        t00 = t
        cloud = dc.global_cloud(clouds=clouds, poses=poses)
        pts = cloud.to_points().detach()
        t = tick('global_cloud.to_points (torch ops)', t)
        bounds = SortedMap.bounds_of(pts)
        t = tick('bounds', t)
        cell = _knn_cell_size(pts, NN_K, NN_R, bounds)
        t = tick('cell size heuristic (cell=%.4f)' % cell, t)
        smap = SortedMap(pts, cell, bounds=bounds)
        t = tick('SortedMap (occupancy %.1f, cells %d)' % (smap.occupancy(), smap.n_cells), t)
        L.profile = {}
        g = search(pts, None, k=NN_K, r=NN_R, cell=cell)
        t = tick('search() total', t)
        gt = g.transposed()
        t = tick('transposed()', t)
        feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, poses=poses),
                                                 neighborhoods=dc.Neighborhoods(g), cfg=cfg)
        feats.step_state()
        t = tick('StepState pack', t)
        print('total %.2f ms' % ((time.perf_counter() - t00) * 1e3))
        for k, v in sorted(L.collect_profile().items(), key=lambda kv: -kv[1]['ms_total']):
            print('     %-26s %7.3f ms x%d' % (k, v['ms_total'], v['calls']))
        L.profile = None
        g._transposed = None


if __name__ == '__main__':
    main()
