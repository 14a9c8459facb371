"""cProfile of Graph.transposed() (developer tool; run on a GPU box)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200.graph import search          # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(n_scans, 'os0-128')
pts = torch.cat([torch.from_numpy(p).to(dev) + torch.tensor([float(i), 0, 0], device=dev) for i, p in enumerate(pts_host)])
g = search(pts, None, k=NN_K, r=NN_R)
for rep in range(3):
    g._transposed = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    g.transposed()
    torch.cuda.synchronize()
    pr.disable()
    print('rep', rep, 'wall %.2f ms' % ((time.perf_counter() - t0) * 1e3), 'allocated %.2f GB reserved %.2f GB' %
          (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
pstats.Stats(pr).sort_stats('tottime').print_stats(8)
