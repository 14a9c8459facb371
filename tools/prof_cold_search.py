"""Where a cold kNN search (no remembered cell size) spends its time: wall-clock of the estimate's stages on the bench map
(developer tool).

    python tools/prof_cold_search.py [n_scans]
"""
import sys
import time

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402,F401
from depth_correction_b200 import graph                 # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(n_scans, 'os0-128')
poses = torch.as_tensor(poses_np, device=dev)
wp = torch.cat([torch.from_numpy(p).to(dev).double() @ T[:3, :3].T + T[:3, 3] for p, T in zip(pts_host, poses)])
for dtype in (torch.float64, torch.float32):
    pts = wp.to(dtype)
    for rep in range(2):
        graph.clear_cell_hints()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bounds = graph.SortedMap.bounds_of(pts, None)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        stages = []
        orig_occ, orig_sample = graph.SortedMap.occupancy_of, graph._knn_cell_from_sample

        def occ(*a, **kw):
            torch.cuda.synchronize(); s = time.perf_counter()
            out = orig_occ(*a, **kw)
            torch.cuda.synchronize(); stages.append(('occupancy_of(cell %.4f) -> %.1f' % (a[3], out), time.perf_counter() - s))
            return out

        def samp(*a, **kw):
            torch.cuda.synchronize(); s = time.perf_counter()
            out = orig_sample(*a, **kw)
            torch.cuda.synchronize(); stages.append(('sample model (c0 %.4f -> %.4f)' % (a[4], out), time.perf_counter() - s))
            return out

        graph.SortedMap.occupancy_of, graph._knn_cell_from_sample = staticmethod(occ), samp
        cell = graph._knn_cell_size(pts, NN_K, NN_R, bounds)
        graph.SortedMap.occupancy_of, graph._knn_cell_from_sample = staticmethod(orig_occ), orig_sample
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        g = graph.search(pts, k=NN_K, r=NN_R, cell=cell)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        del g
        print('%s rep %d: bounds %.2f ms, cell estimate %.2f ms (%s), search at cell %.4f: %.2f ms' % (
            dtype, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), '; '.join('%s %.2f ms' % (n, 1e3 * t) for n, t in stages), cell, 1e3 * (t3 - t2)))
