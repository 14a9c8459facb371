"""dc_knn_cells (warp per cell) against dc_knn (thread per query) on the GPU box: identical neighbour sets, timing of
both kernels, share of the queries the cell kernel hands to the fp64 thread path (developer tool).

    python tools/check_knn_cells.py [n_scans] [occ,occ,...]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402,F401
from depth_correction_b200 import _lib as L             # noqa: E402
from depth_correction_b200.graph import search          # noqa: E402
from bench import host_scans, NN_K, NN_R                # noqa: E402


def rows_sorted(g):
    """ELL rows as an [n, k] matrix with every row sorted (the kernels emit unordered rows)."""
    k = g.k
    n = g.n_rows
    ns = (n + 31) // 32
    ell = g.ell_idx[:ns * 32 * k].view(ns, k, 32).permute(0, 2, 1).reshape(ns * 32, k)[:n]
    return ell.sort(dim=1).values


def timed_search(pts, query, k, r, path, cell=None, reps=3):
    os.environ['DC_KNN'] = path
    name = {'cells': 'dc_knn_cells', 'thread': 'dc_knn'}[path]
    best, g = None, None
    for _ in range(reps):
        del g
        L.profile = {}
        g = search(pts, query, k=k, r=r, cell=cell)
        torch.cuda.synchronize()
        prof = L.collect_profile()
        L.profile = None
        ms = prof[name]['ms_total']
        best = ms if best is None else min(best, ms)
    fb = None
    if path == 'cells':
        ws = L._workspace.get(('temp:dc_knn_cells', str(pts.device)))
        hdr = ws[:64].view(torch.int32).cpu()
        nq = g.n_rows
        off = 64 + ((max(nq, 1) * 4 + 15) // 16) * 16
        nfb = int(hdr[5])
        reasons = ws[off:off + 8 * nfb].view(torch.int32).view(-1, 2)[:, 1] >> 8
        rc = torch.bincount(reasons.long(), minlength=4).tolist() if nfb else [0, 0, 0, 0]
        fb = (int(hdr[0]), nfb, rc[1:4])
    return g, best, fb


def compare(name, pts, query, k, r, cell=None):
    gt, mt, _ = timed_search(pts, query, k, r, 'thread', cell)
    a = rows_sorted(gt)
    cellsz, occ = gt.map.cell, gt.map.occupancy()
    del gt
    gc, mc, fb = timed_search(pts, query, k, r, 'cells', cell)
    b = rows_sorted(gc)
    nq = gc.n_rows
    same = torch.equal(a, b)
    nbad = int((a != b).any(dim=1).sum().item()) if not same else 0
    print('%-28s n=%8d nq=%8d k=%3d r=%s cell=%.4f occ=%5.1f | thread %8.3f ms  cells %8.3f ms (x%.2f) | cells %d fp64 path %d (%.2f%%: ambiguous / crowded / ring %s) | %s'
          % (name, pts.shape[0], nq, k, r, cellsz, occ, mt, mc, mt / mc, fb[0], fb[1], 100.0 * fb[1] / max(nq, 1), fb[2],
             'IDENTICAL' if same else 'MISMATCH in %d rows' % nbad), flush=True)
    if not same:
        bad = torch.nonzero((a != b).any(dim=1))[:3, 0].tolist()
        for q in bad:
            print('   row', q, 'thread', a[q].tolist(), 'cells', b[q].tolist())
    return same


def main():
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    occs = [float(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0.3]
    dev = torch.device('cuda:0')
    ok = True
    rng = np.random.default_rng(7)
    # clustered random points, duplicates, isolated points
    c = rng.uniform(-20, 20, (400, 3))
    pts = (c[rng.integers(0, 400, 200000)] + rng.normal(0, 0.3, (200000, 3))).astype(np.float32)
    pts[1000:1400] = pts[:400]
    pts[2000:2040] = pts[0]
    pts[3000:3200] = rng.uniform(-40, 40, (200, 3)).astype(np.float32)
    p = torch.as_tensor(pts, device=dev)
    for kw in (dict(k=16, r=0.25), dict(k=8, r=None), dict(k=32, r=0.1), dict(k=1, r=None), dict(k=64, r=0.5)):
        ok &= compare('clustered', p, None, kw['k'], kw['r'])
    q = torch.as_tensor((pts[::7] + rng.normal(0, 0.05, pts[::7].shape)).astype(np.float32), device=dev)
    ok &= compare('clustered cross query', p, q, 4, None)
    ok &= compare('clustered cross k=1', p, q, 1, 0.2)
    # lattice: exact ties everywhere
    gx = np.stack(np.meshgrid(np.arange(40), np.arange(40), np.arange(12), indexing='ij'), -1).reshape(-1, 3).astype(np.float32) * 0.25
    lat = torch.as_tensor(gx, device=dev)
    ok &= compare('lattice (ties)', lat, None, 7, None)
    ok &= compare('lattice (ties) k=27 r', lat, None, 27, 0.5)
    # lidar corridor
    pts_host, poses_np = host_scans(n_scans, 'os0-128')
    poses = torch.as_tensor(poses_np, device=dev)
    world = []
    for ph, T in zip(pts_host, poses):
        x = torch.from_numpy(ph).to(dev).double()
        world.append((x @ T[:3, :3].T + T[:3, 3]).float())
    wp = torch.cat(world)
    for cell in (None, 0.02, 0.11):
        ok &= compare('corridor %d scans cell=%s' % (n_scans, cell), wp, None, NN_K, NN_R, cell)
    ok &= compare('corridor k=16 no r', wp, None, 16, None)
    ok &= compare('corridor k=64 r=1.0', wp, None, 64, 1.0)
    for occ in occs:
        os.environ['DC_KNN_OCC'] = str(occ)
        for pop in os.environ.get('POPS', '25').split(','):
            os.environ['DC_KNN_POP_X10'] = pop
            ok &= compare('corridor occ=%s pop=%s' % (occ, pop), wp, None, NN_K, NN_R)
    print('ALL IDENTICAL' if ok else 'MISMATCHES FOUND')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
