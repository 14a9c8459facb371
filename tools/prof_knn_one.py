"""One warm + one measured kNN search on the corridor map with the implementation named by DC_KNN, for ncu captures.

    DC_KNN=record python tools/prof_knn_one.py [n_scans]
"""
import sys

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402,F401
from depth_correction_b200.graph import search          # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(n_scans, 'os0-128')
poses = torch.as_tensor(poses_np, device=dev)
wp = torch.cat([(torch.from_numpy(p).to(dev).double() @ T[:3, :3].T + T[:3, 3]).float() for p, T in zip(pts_host, poses)])
for rep in range(2):
    g = search(wp, None, k=NN_K, r=NN_R)
    torch.cuda.synchronize()
print('n = %d, cell %.4f' % (len(wp), g.map.cell))
