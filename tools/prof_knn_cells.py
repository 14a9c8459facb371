"""One warm + one measured kNN search on the corridor map (cell path), for ncu captures of knn_cell_kernel (GPU box).

    python tools/prof_knn_cells.py [n_scans] [occ]
"""
import os
import sys

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402,F401
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import search          # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 2:
    os.environ['DC_KNN_OCC'] = sys.argv[2]
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(n_scans, 'os0-128')
poses = torch.as_tensor(poses_np, device=dev)
wp = torch.cat([(torch.from_numpy(p).to(dev).double() @ T[:3, :3].T + T[:3, 3]).float() for p, T in zip(pts_host, poses)])
for rep in range(2):
    L.profile = {}
    g = search(wp, None, k=NN_K, r=NN_R)
    torch.cuda.synchronize()
    prof = L.collect_profile()
    L.profile = None
    ws = L._workspace.get(('temp:dc_knn_cells', str(dev)))
    hdr = ws[:64].view(torch.int32).cpu()
    print('n=%d cell=%.4f dc_knn_cells %.3f ms, cells %d, fallback queries %d' % (wp.shape[0], g.map.cell, prof['dc_knn_cells']['ms_total'], int(hdr[0]), int(hdr[5])))
    del g
