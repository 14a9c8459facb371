"""kNN kernel timing over cell-size / axis-order / code-path variants on the bench map (developer tool; GPU box).

    python tools/prof_knn.py [n_scans] [occ,occ,...]
"""
import os
import sys
import time

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import search          # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402


def main():
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    occs = [float(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0.45, 0.7, 1.0, 1.4]
    dev = torch.device('cuda:0')
    pts_host, poses_np = host_scans(n_scans, 'os0-128')
    poses = torch.as_tensor(poses_np, device=dev)
    world = []
    for p, T in zip(pts_host, poses):
        x = torch.from_numpy(p).to(dev).double()
        world.append(x @ T[:3, :3].T + T[:3, 3])
    pts = torch.cat(world)
    print('points', pts.shape[0])
    ref = None
    for path in os.environ.get('PATHS', 'thread').split(','):
        for order in os.environ.get('ORDERS', 'long,short').split(','):
            for occ in occs:
                os.environ['DC_KNN_PATH'] = path
                os.environ['DC_AXIS_ORDER'] = order
                os.environ['DC_KNN_OCC'] = str(occ)
                best = None
                for rep in range(3):
                    L.profile = {}
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    g = search(pts, None, k=NN_K, r=NN_R)
                    torch.cuda.synchronize()
                    wall = (time.perf_counter() - t0) * 1e3
                    prof = L.collect_profile()
                    L.profile = None
                    knn = prof['dc_knn']['ms_total']
                    if best is None or knn < best[0]:
                        best = (knn, wall, g.map.cell, g.map.occupancy())
                # order-independent checksum of the graph in ORIGINAL indices
                nb = g.neighbors()
                chk = int((nb.clamp(min=0) * (torch.arange(nb.shape[0], device=dev)[:, None] % 1000 + 1)).sum().item())
                if ref is None:
                    ref = chk
                print('path=%-6s order=%-5s occ=%.2f  cell=%.4f occupancy=%5.1f  dc_knn %7.3f ms  search wall %7.2f ms  %s'
                      % (path, order, occ, best[2], best[3], best[0], best[1], 'OK' if chk == ref else 'MISMATCH'))
                del g, nb


if __name__ == '__main__':
    main()
