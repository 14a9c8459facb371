"""Kernel time of dc_knn_recorded against the cell size, next to the distribution of the k-th neighbour distance
(developer tool; what the cell-size estimate of graph._knn_cell_size should aim at).

    python tools/knn_cell_sweep.py [n_scans]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402,F401
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import search          # noqa: E402
from depth_correction_b200.synthetic import make_sequence  # noqa: E402

dev = torch.device('cuda:0')


def world(scene, n_scans, pattern='os0-128', **kw):
    scans, poses, _ = make_sequence(scene, n_scans=n_scans, pattern=pattern, seed=0, **kw)
    return torch.as_tensor(np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)]).astype(np.float32), device=dev)


def kernel_ms(pts, k, r, cell, reps=3):
    best = None
    for _ in range(reps):
        L.profile = {}
        g = search(pts, k=k, r=r, cell=cell)
        torch.cuda.synchronize()
        prof = L.collect_profile()
        L.profile = None
        ms = prof['dc_knn_recorded']['ms_total']
        best = ms if best is None else min(best, ms)
        del g
    return best


n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_street = int(sys.argv[2]) if len(sys.argv) > 2 else 60
only = os.environ.get('SWEEP_ONLY')           # substring of the case names to run
cases = [('corridor %d scans k=32 r=0.4' % n_scans, world('corridor', n_scans), 32, 0.4),
         ('corridor 8 scans k=32 r=0.4', world('corridor', 8), 32, 0.4),
         ('corridor 8 scans k=16', world('corridor', 8), 16, None),
         ('corridor 8 scans k=64 r=0.5', world('corridor', 8), 64, 0.5),
         ('street 8 HDL-64 scans k=32 r=0.4', world('street', 8, pattern='hdl-64', depth_clip=(5.0, 80.0)), 32, 0.4),
         ('street %d HDL-64 scans k=32 r=0.4' % n_street, (lambda: world('street', n_street, pattern='hdl-64', depth_clip=(5.0, 80.0))), 32, 0.4)]
rng = np.random.default_rng(3)
c = rng.uniform(-10, 10, (200, 3))
clu = (c[rng.integers(0, 200, 300000)] + rng.normal(0, 0.3, (300000, 3))).astype(np.float32)
cases.append(('clustered k=8', torch.as_tensor(clu, device=dev), 8, None))
for name, pts, k, r in cases:
    if only and only not in name:
        continue
    if callable(pts):
        pts = pts()
    g = search(pts, k=k, r=r)
    cell0 = g.map.cell
    dk = g.distances()[:, k - 1]
    del g
    fin = torch.isfinite(dk)
    cap = float(r) if r else float(dk[fin].max())
    dkc = torch.where(fin, dk, torch.full_like(dk, cap))
    qs = torch.tensor([0.5, 0.6, 0.7, 0.8, 0.9, 0.95], dtype=torch.float64, device=dev)
    sample = dkc[torch.randint(0, len(dkc), (200000,), device=dev)]
    pct = torch.quantile(sample, qs).tolist()
    print('%s  n=%d  default cell %.4f  d_k quantiles 50/60/70/80/90/95 %%: %s  (no k-th neighbour within r: %.1f %%)' % (
        name, len(pts), cell0, ' '.join('%.4f' % v for v in pct), 100.0 * (1.0 - fin.double().mean().item())), flush=True)
    for f in (0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.25, 1.4, 1.6):
        cell = cell0 * f
        if r and cell > r:
            continue
        share = (sample < cell).double().mean().item()
        print('   cell %.4f (%.2f x)  d_k < cell for %.1f %% of the queries | dc_knn_recorded %.3f ms' % (
            cell, f, 100.0 * share, kernel_ms(pts, k, r, cell)), flush=True)
    del dk, dkc, sample
    # what the estimate of graph._knn_cell_size picks (cost model on a sample of queries), and the cold search it costs
    from depth_correction_b200.graph import clear_cell_hints
    for mode in ('occ', 'model'):
        os.environ['DC_KNN_CELL'] = mode
        clear_cell_hints()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = search(pts, k=k, r=r)
        e1.record()
        torch.cuda.synchronize()
        cell = g.map.cell
        del g
        print('   estimate %-5s -> cell %.4f  cold search %.2f ms | dc_knn_recorded %.3f ms' % (
            mode, cell, e0.elapsed_time(e1), kernel_ms(pts, k, r, cell)), flush=True)
    os.environ.pop('DC_KNN_CELL', None)
