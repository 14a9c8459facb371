"""Radius-mode search (count + fill passes) on three maps; run twice (DC_RADIUS_FILL=direct / default) to A/B the fill
kernel.  Prints per-entry-point CUDA-event times and a checksum of the lists.

    python tools/prof_radius.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import search         # noqa: E402
from depth_correction_b200.synthetic import make_sequence  # noqa: E402

dev = torch.device('cuda:0')
print('DC_RADIUS_FILL =', os.environ.get('DC_RADIUS_FILL', '(staged rows)'))


def world(scene, n_scans, grid_res, pattern='os0-128', **kw):
    scans, poses, _ = make_sequence(scene, n_scans=n_scans, pattern=pattern, seed=5, grid_res=grid_res, **kw)
    return torch.as_tensor(np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)]).astype(np.float32), device=dev)


for name, pts, r in (('fee 12 scans 0.1 m voxels, r=0.25', world('fee', 12, 0.1), 0.25),
                     ('corridor 16 full scans, r=0.15', world('corridor', 16, 0.0), 0.15),
                     ('street 8 HDL-64 scans, r=0.4', world('street', 8, 0.0, pattern='hdl-64', depth_clip=(5.0, 80.0)), 0.4)):
    for rep in range(3):
        L.profile = {} if rep == 2 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = search(pts, k=0, r=r)
        e1.record()
        torch.cuda.synchronize()
    times = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in L.profile.items()}
    L.profile = None
    idx = g.ell_idx.long()
    chk = int(((idx + 2) * (torch.arange(idx.numel(), device=dev) % 1000003 + 1)).sum() % (1 << 61))
    print('%-40s n=%8d width=%4d mean=%6.1f search %.3f ms  count %.3f  fill %.3f  checksum %d' % (
        name, len(pts), g.width, float((g.ell_idx >= 0).sum()) / len(pts), e0.elapsed_time(e1),
        times.get('dc_radius_count', 0), times.get('dc_radius_fill', 0), chk))
