"""The BENCH MAP ITSELF against the CPU oracle (the largest full check was 1.18 M points at k = 16).

64 full-resolution OS0-128 corridor scans (8.37 M points), kNN k = 32 within r = 0.4 m, ScaledPolynomial[2,4],
min_eigval_loss(normalization=True), per-scan pose corrections -- the configuration of the bench line.

  * neighbour indices of ALL rows: the GPU graph exported in the reference layout == cKDTree.query(k, distance_upper_bound)
  * loss, dL/dw, dL/dposes: GPU fused step (default float32 scatter form, and the fp64 gather form) against the oracle's
    torch fp64 autograd, evaluated in chunks of rows (the [N,K,3,3] intermediates of utils.py:109-149 do not fit in one
    piece at this size): per-chunk sums of the per-point loss, dL/dp accumulated over the chunks, then one backward
    through model + poses.

Takes a few minutes of CPU time; run on a GPU box:  python tools/full_map_parity.py [--scans 64] > profiles/...
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

ap = argparse.ArgumentParser()
ap.add_argument('--scans', type=int, default=64)
ap.add_argument('--k', type=int, default=32)
ap.add_argument('--r', type=float, default=0.4)
ap.add_argument('--chunk', type=int, default=1 << 19)
args = ap.parse_args()

import depth_correction_b200 as dc                              # noqa: E402
from depth_correction_b200 import fused                         # noqa: E402
from depth_correction_b200.synthetic import make_sequence       # noqa: E402
from oracle import oracle                                       # noqa: E402

dev = torch.device('cuda:0')
torch.set_num_threads(os.cpu_count())
scans_np, _, poses = make_sequence('corridor', n_scans=args.scans, pattern='os0-128', seed=0)
rng = np.random.default_rng(2)
cfg = dc.Config(nn_k=args.k, nn_r=args.r, pose_correction=dc.PoseCorrection.pose)
clouds, oscans = [], []
for s in scans_np:
    inc = rng.uniform(0.05, 1.3, (len(s['points']), 1)).astype(np.float32)
    msk = rng.random(len(s['points'])) < 0.9
    c = dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev))
    c.inc_angles = torch.as_tensor(inc, device=dev)
    c.mask = torch.as_tensor(msk, device=dev)
    clouds.append(c)
    # the oracle consumes the float32 records the kernels see, up-cast to float64
    oscans.append({'vps': c.vps.double().cpu(), 'dirs': c.dirs.double().cpu(), 'depth': c.depth.double().cpu(),
                   'inc_angles': torch.as_tensor(inc.astype(np.float64)), 'mask': torch.as_tensor(msk)})
n = sum(len(c) for c in clouds)
poses_t = torch.as_tensor(poses, device=dev)
d0 = torch.as_tensor(rng.normal(0, 2e-3, (len(clouds), 6)), device=dev)
w0 = [0.004, -0.003]
out = {'workload': 'bench map: corridor, %d OS0-128 scans, %d points, kNN k=%d within r=%g' % (args.scans, n, args.k, args.r),
       'cpu_threads': os.cpu_count()}

# ---- search: every row -----------------------------------------------------------------------------------------
ns = dc.establish_neighborhoods(clouds=clouds, poses=poses_t, cfg=cfg)
nb_gpu = ns[0].cpu()
t0 = time.perf_counter()
pts0, _ = oracle.global_points(oscans, torch.as_tensor(poses))
_, nb = oracle.nearest_neighbors(pts0, k=args.k, r=args.r)
out['cpu_search_s'] = time.perf_counter() - t0
diff_rows = int((nb_gpu != nb).any(dim=1).sum())
if diff_rows:      # exact distance ties are ordered by original index here, by traversal order in cKDTree: compare as sets
    rows = (nb_gpu != nb).any(dim=1).nonzero().ravel()
    set_diff = int((nb_gpu[rows].sort(dim=1).values != nb[rows].sort(dim=1).values).any(dim=1).sum())
else:
    set_diff = 0
out.update({'rows': int(nb.shape[0]), 'rows_with_a_different_order': diff_rows, 'rows_with_a_different_set': set_diff,
            'valid_neighbors': int((nb >= 0).sum())})
del nb_gpu

# ---- step: chunked oracle ------------------------------------------------------------------------------------
t0 = time.perf_counter()
w = torch.tensor([w0], dtype=torch.float64, requires_grad=True)
exponent = torch.tensor([[2.0, 4.0]], dtype=torch.float64)
deltas_c = d0.cpu().clone().requires_grad_(True)
poses_c = oracle.create_corrected_poses(torch.as_tensor(poses), deltas_c)
points, _ = oracle.global_points(oscans, poses_c, w, exponent, True)
leaf = points.detach().clone().requires_grad_(True)
total = 0.0
for a in range(0, n, args.chunk):
    feats = oracle.neighborhood_features(leaf, nb[a:a + args.chunk], eigvecs=False)
    val, _ = oracle.min_eigval_loss(feats['eigvals'], None, normalization=True, reduction='sum')
    val.backward()
    total += float(val)
    del feats, val
loss_ref = total / n
points.backward(leaf.grad / n)
out['cpu_step_s'] = time.perf_counter() - t0
del leaf

rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
for form in ('auto', 'gather'):
    fused.set_backward_form(form)
    model = dc.ScaledPolynomial(w=w0, exponent=[2, 4], device=dev)
    deltas = d0.clone().requires_grad_(True)
    pc = torch.stack(dc.create_corrected_poses(poses_t, deltas, cfg))
    feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
    loss, _ = dc.min_eigval_loss(feats, normalization=True)
    loss.backward()
    out['backward_form_' + form] = {'loss_rel_err': abs(loss.item() - loss_ref) / abs(loss_ref),
                                    'w_grad_rel_err': rel(model.w.grad.cpu(), w.grad),
                                    'pose_grad_rel_err': rel(deltas.grad.cpu(), deltas_c.grad)}
out['loss'] = loss_ref
tol = {'auto': 1e-5, 'gather': 1e-8}
out['ok'] = bool(set_diff == 0 and all(max(v for v in out['backward_form_' + f].values()) < tol[f] for f in tol))
print(json.dumps(out, indent=1))
sys.exit(0 if out['ok'] else 1)
