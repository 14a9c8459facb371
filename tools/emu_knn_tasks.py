#!/usr/bin/env python
"""Host emulation of the warp-task kNN kernel's work (candidate visits and a warp-instruction model) on a synthetic
lidar map, used to choose the cell size / ring policy before spending GPU time (numpy + cKDTree, CPU only).

A task = up to 32 consecutive queries of the cell-sorted map that lie in ONE row of cells (fastest axis = longest
extent) and span at most `span` cells; its candidates are the (2m+1)^2 rows x cells [c0_first - m, c0_last + m].
A lane is finished at ring m when its k-th neighbour distance is below m * cell (or r <= m * cell).
"""
import argparse
import sys
import os

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from depth_correction_b200.synthetic import make_sequence  # noqa: E402


def world_points(scene, n_scans, pattern):
    clip = (1.0, 25.0) if scene == 'corridor' else (5.0, 80.0)
    scans, poses, _ = make_sequence(scene, n_scans=n_scans, pattern=pattern, seed=0, depth_clip=clip)
    out = []
    for s, T in zip(scans, poses):
        p = s['points'].astype(np.float64)
        out.append((p @ T[:3, :3].T + T[:3, 3]).astype(np.float32).astype(np.float64))
    return np.concatenate(out)


def emulate(x, dk, k, r, cell, span_cap, instr_per_cand=22.0, overhead=1100.0, alpha=2.0, verbose=True):
    lo = x.min(0) - 1e-3 * cell
    ext = x.max(0) - x.min(0)
    axes = np.argsort(-ext, kind='stable')
    c = np.floor((x[:, axes] - lo[axes]) / cell).astype(np.int64)
    dims = c.max(0) + 1
    key = (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]
    order = np.argsort(key, kind='stable')
    key, c, dks = key[order], c[order], dk[order]
    n = len(key)
    n_cells = int(dims.prod())
    # dense cell table
    start = np.searchsorted(key, np.arange(n_cells + 1))
    need = np.where(np.isfinite(dks), np.floor(dks / (cell * (1 - 1e-9))).astype(np.int64) + 1, int(np.ceil(r / cell)))
    need = np.minimum(need, int(np.ceil(r / cell)))
    # tasks: slices of 32, split by row and span cap
    row = key // dims[0]
    total_instr = 0.0
    total_cand = 0
    total_tasks = 0
    rounds_hist = {}
    sl = np.arange(0, n, 32)
    # vectorised over segment starts: iteratively peel segments off every slice
    seg_first = sl.copy()
    seg_end = np.minimum(sl + 32, n)
    active = np.ones(len(sl), bool)
    lanes_used = 0
    while active.any():
        f = seg_first[active]
        e = seg_end[active]
        # lanes f..e-1 ; segment = prefix of lanes with same row and c0 - c0[f] <= span_cap
        L = np.zeros(len(f), np.int64)
        maxneed = np.zeros(len(f), np.int64)
        minneed = np.full(len(f), 1 << 30, np.int64)
        c0l = c[f, 0].copy()
        ok = np.ones(len(f), bool)
        for t in range(32):
            j = f + t
            valid = ok & (j < e)
            jj = np.minimum(j, n - 1)
            valid &= (row[jj] == row[f]) & (c[jj, 0] - c[f, 0] <= span_cap)
            ok = valid
            L += valid
            maxneed = np.where(valid, np.maximum(maxneed, need[jj]), maxneed)
            minneed = np.where(valid, np.minimum(minneed, need[jj]), minneed)
            c0l = np.where(valid, c[jj, 0], c0l)
        c0f = c[f, 0]
        c1 = c[f, 1]
        c2 = c[f, 2]
        # candidate count per m for m in 1..max(maxneed)
        mmax = int(maxneed.max())
        cand_m = {}
        for m in range(1, mmax + 1):
            tot = np.zeros(len(f), np.int64)
            a = np.clip(c0f - m, 0, dims[0] - 1)
            b = np.clip(c0l + m, 0, dims[0] - 1)
            for e2 in range(-m, m + 1):
                for e1 in range(-m, m + 1):
                    y, z = c1 + e1, c2 + e2
                    inside = (y >= 0) & (y < dims[1]) & (z >= 0) & (z < dims[2])
                    base = (np.clip(z, 0, dims[2] - 1) * dims[1] + np.clip(y, 0, dims[1] - 1)) * dims[0]
                    tot += np.where(inside, start[base + b + 1] - start[base + a], 0)
            cand_m[m] = tot
        # policy: start at the first m whose population >= alpha * k (or m needed by r), then grow by one until all lanes done
        m_start = np.ones(len(f), np.int64)
        for m in range(1, mmax + 1):
            m_start = np.where((m_start == m) & (cand_m[m] < alpha * k) & (m < maxneed), m + 1, m_start)
        cost = np.zeros(len(f))
        cands = np.zeros(len(f), np.int64)
        rounds = np.zeros(len(f), np.int64)
        for m in range(1, mmax + 1):
            run = (m >= m_start) & (m <= maxneed)
            cost += np.where(run, instr_per_cand * cand_m[m] + overhead, 0.0)
            cands += np.where(run, cand_m[m], 0)
            rounds += run
        total_instr += cost.sum()
        total_cand += cands.sum()
        total_tasks += len(f)
        lanes_used += L.sum()
        for rr, cnt in zip(*np.unique(rounds, return_counts=True)):
            rounds_hist[int(rr)] = rounds_hist.get(int(rr), 0) + int(cnt)
        # advance
        nf = f + L
        seg_first[active] = nf
        still = nf < e
        idx = np.where(active)[0]
        active[idx[~still]] = False
    occ = n / len(np.unique(key))
    if verbose:
        print('cell %.4f occ %.1f span %d alpha %.1f: tasks/slice %.2f  cand/query %.0f  warp-instr/32q %.0f  total %.2f G (scaled to 8.37M: %.2f G)  rounds %s'
              % (cell, occ, span_cap, alpha, total_tasks / len(sl), total_cand / n * 32 / 32, total_instr / len(sl),
                 total_instr / 1e9, total_instr / n * 8.37e6 / 1e9, rounds_hist))
    return total_instr / n


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--scans', type=int, default=16)
    ap.add_argument('--scene', default='corridor')
    ap.add_argument('--pattern', default='os0-128')
    ap.add_argument('--k', type=int, default=32)
    ap.add_argument('--r', type=float, default=0.4)
    ap.add_argument('--cells', default='0.035,0.042,0.05,0.06,0.07')
    ap.add_argument('--spans', default='2,4')
    ap.add_argument('--alphas', default='2.0')
    a = ap.parse_args()
    x = world_points(a.scene, a.scans, a.pattern)
    print('points', len(x))
    cache = '/tmp/emu_dk_%s_%d_%s_%d_%g.npy' % (a.scene, a.scans, a.pattern, a.k, a.r)
    if os.path.exists(cache):
        dk = np.load(cache)
    else:
        t = cKDTree(x)
        d, _ = t.query(x, k=a.k, distance_upper_bound=a.r, workers=-1)
        dk = d[:, -1]
        np.save(cache, dk)
    print('d_k percentiles 1/50/95/99:', np.percentile(dk[np.isfinite(dk)], [1, 50, 95, 99]), 'inf frac', np.mean(~np.isfinite(dk)))
    for cell in [float(v) for v in a.cells.split(',')]:
        for sp in [int(v) for v in a.spans.split(',')]:
            for al in [float(v) for v in a.alphas.split(',')]:
                emulate(x, dk, a.k, a.r, cell, sp, alpha=al)
