"""Final ring / record length statistics of dc_knn_recorded on the corridor map (library built with -DDC_KNN_STATS)."""
import os
import sys
import torch
sys.path.insert(0, '.')
os.environ['DC_KNN'] = 'record'
import depth_correction_b200 as dc                      # noqa: E402,F401
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import search          # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(int(sys.argv[1]) if len(sys.argv) > 1 else 16, 'os0-128')
poses = torch.as_tensor(poses_np, device=dev)
wp = torch.cat([(torch.from_numpy(p).to(dev).double() @ T[:3, :3].T + T[:3, 3]).float() for p, T in zip(pts_host, poses)])
g = search(wp, None, k=NN_K, r=NN_R)
torch.cuda.synchronize()
ws = L._workspace.get(('temp:dc_knn_recorded', str(dev)))
c = ws[:64].view(torch.int32).cpu()
n = len(wp)
print('n %d cell %.4f max_ring %d' % (n, g.map.cell, -(-NN_R // g.map.cell)))
print('final ring histogram (1..8, 9+):', [round(int(c[2 + r]) / n, 4) for r in range(1, 10)])
print('mean words / query: %.1f' % (int(ws[56:64].view(torch.int64)) / n))
cyc = ws[64 + 8 * n - 128:64 + 8 * n].view(torch.int64).cpu().double()
print('share of the thread cycles by final ring (1..8, 9+):', [round(float(cyc[r] / cyc.sum()), 4) for r in range(1, 10)])
