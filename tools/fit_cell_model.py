"""Offline fit of the constants of graph._knn_cell_of_distances against MEASURED kernel times (developer tool, CPU only).

The sweeps of profiles/r2_knn_cell_sweep.log give, for seven seeded maps, the time of the kNN kernel against the cell edge.
This script regenerates the same maps, takes the sorted neighbour distances of 8192 sample points with cKDTree (what the
model sees at run time), and reports for each candidate form of the model the REGRET per map: measured time at the cell the
model picks / best time of the sweep.

    python tools/fit_cell_model.py            # ~2 minutes, 6 GB of host memory for the 57 M point map
"""
import itertools
import math
import os
import re
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from depth_correction_b200.synthetic import make_sequence   # noqa: E402

DENSE = 1 << 30


def world(scene, n_scans, pattern='os0-128', **kw):
    scans, poses, _ = make_sequence(scene, n_scans=n_scans, pattern=pattern, seed=0, **kw)
    return np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)]).astype(np.float32).astype(np.float64)


def street(n):
    return lambda: world('street', n, pattern='hdl-64', depth_clip=(5.0, 80.0))


MAPS = {'corridor 64 scans k=32 r=0.4': (lambda: world('corridor', 64), 32, 0.4),
        'corridor 8 scans k=32 r=0.4': (lambda: world('corridor', 8), 32, 0.4),
        'corridor 8 scans k=64 r=0.5': (lambda: world('corridor', 8), 64, 0.5),
        'street 8 HDL-64 scans k=32 r=0.4': (street(8), 32, 0.4),
        'street 60 HDL-64 scans k=32 r=0.4': (street(60), 32, 0.4),
        'street 75 HDL-64 scans k=32 r=0.4': (street(75), 32, 0.4),
        'street 600 HDL-64 scans k=32 r=0.4': (street(600), 32, 0.4)}
# the 75-scan sweep (one slab's worth of points) was run by hand: SWEEP_ONLY="street 75" python tools/knn_cell_sweep.py 8 75
CURVES = {'street 75 HDL-64 scans k=32 r=0.4': [(.0328, 4.228), (.0382, 3.874), (.0437, 4.26), (.0492, 4.373), (.0546, 4.666),
                                                (.0601, 4.89), (.0683, 5.581), (.0765, 7.278), (.0874, 10.244)]}


def read_curves(path):
    name = None
    for ln in open(path):
        m = re.match(r'^(\S.*?)\s+n=\d+', ln)
        if m:
            name = m.group(1).strip()
            continue
        m = re.match(r'\s+cell ([\d.]+) ', ln)
        t = re.search(r'rec\[7\] ([\d.]+) ms', ln) or re.search(r'dc_knn_recorded ([\d.]+) ms', ln)
        if m and t and name in MAPS:
            CURVES.setdefault(name, []).append((float(m.group(1)), float(t.group(1))))
    # below 0.0337 m the 600-scan map loses its dense cell table (a cliff the model treats as a hard limit)
    CURVES['street 600 HDL-64 scans k=32 r=0.4'] = [p for p in CURVES['street 600 HDL-64 scans k=32 r=0.4'] if p[0] >= 0.0345]


def grid_cells(lo, hi, c):
    n = 1
    for a, b in zip(lo, hi):
        n *= int(math.floor((b - (a - 1e-3 * c)) / c)) + 1
    return n


def ring_seq(max_ring):
    seq, r = [], 1
    while True:
        seq.append(min(r, max_ring))
        if r >= max_ring:
            return seq
        r = r + 1 if r < 4 else r * 2


def model_cost(D, c, beta, face, tcost, soft=False):
    d, r, n = D['d'], D['r'], D['n']
    fin = np.isfinite(d)
    nv = np.maximum(fin.sum(1), 1)
    dk, have = d[:, -1], fin[:, -1]
    sigma = nv / (np.pi * np.maximum(np.where(have, dk, r), 1e-12) ** 2)
    max_ring = int(math.ceil(r / c))
    cost, alive = np.zeros(len(dk)), np.ones(len(dk))
    for rho in ring_seq(64):
        rc = min(rho, max_ring)
        cost += alive * (2 * rc + 1) ** 2 * (sigma * c * c + beta)
        delta = (np.where(have, dk, np.inf) - rc * c) / c
        if soft:      # probability over the position of the query in its cell that the reach rc c + face covers d_k
            p = np.where(delta < 0, 1.0, np.where(delta < 0.5, (1 - 2 * np.clip(delta, 0, 0.5)) ** 3, 0.0))
        else:
            p = (delta < face).astype(float)
        if rc >= max_ring:
            p = np.ones(len(dk))
        alive = alive * (1 - p)
        if alive.max() <= 0:
            break
    nc = grid_cells(D['lo'], D['hi'], c)
    return cost.mean() + (tcost * nc / n if nc <= DENSE else 1e18)


def measured(curve, c):
    cs, ts = np.array([x for x, _ in curve]), np.array([y for _, y in curve])
    return float(np.interp(math.log(min(max(c, cs[0]), cs[-1])), np.log(cs), ts))


def main():
    read_curves(os.path.join(ROOT, 'profiles', 'r2_knn_cell_sweep.log'))
    data = {}
    for name, (make, k, r) in MAPS.items():
        P = make()
        idx = np.random.default_rng(0).choice(len(P), 8192, replace=False)
        d, _ = cKDTree(P).query(P[idx], k=k, distance_upper_bound=r, workers=-1)
        data[name] = dict(d=d, n=len(P), r=r, lo=P.min(0).tolist(), hi=P.max(0).tolist())
        print('%-36s n = %9d  median d_k %.4f' % (name, len(P), np.median(d[:, -1])), flush=True)
    names = list(MAPS)
    rows = []
    for beta, face, tcost, soft in itertools.product((2, 3, 4, 6, 8), (0.0, 0.125, 0.25), (0.0, 0.55), (False, True)):
        if soft and face != 0.0:
            continue
        reg = []
        for name in names:
            cs = np.array([x for x, _ in CURVES[name]])
            cand = np.geomspace(cs[0], cs[-1], 40)
            c = cand[int(np.argmin([model_cost(data[name], cc, beta, face, tcost, soft) for cc in cand]))]
            reg.append(measured(CURVES[name], c) / min(t for _, t in CURVES[name]))
        rows.append((np.mean(reg), max(reg), beta, 'expectation' if soft else face, tcost, reg))
    rows.sort(key=lambda x: (x[0], x[1]))
    print('maps: ' + '; '.join(names))
    print('mean / max regret | row cost, face credit (cells), table cost | regret per map')
    for row in rows[:10] + [x for x in rows if x[2:5] == (4, 0.125, 0.55)]:
        print('%.3f %.3f | %s %s %s | %s' % (row[0], row[1], row[2], row[3], row[4], ' '.join('%.3f' % v for v in row[5])))


if __name__ == '__main__':
    main()
