"""Multi-GPU execution of the map-consistency step: one process per GPU, spatial slabs, one-time halo
exchange over NCCL (NVLink), one small all-reduce per iteration.  (the reference
is single-process and has no counterpart.)

The loss is a sum of per-point terms l_i that depend only on points within the neighbour radius r of
point i, so the map shards by loss term:

  1. every rank ingests a subset of the scans (any split) and computes its points in the initial map frame;
  2. `plan()` agrees on the split axis (longest extent of the global bounding box) and on G slab boundaries
     with equal point counts (all-reduce of a 1-D histogram);
  3. `exchange()` routes every point record to the rank that owns its slab and, as a read-only HALO copy, to
     every rank whose slab lies within `halo` (= r) of it -- one all_to_all_single per record tensor.
     Model weights and pose corrections are replicated, so halo points are re-corrected and re-posed locally
     every iteration: there is NO per-iteration point exchange;
  4. each rank searches and runs the fused step on (owned + halo) points with the loss mask restricted to
     owned points; gradients w.r.t. halo points flow into the local dw / dpose accumulators;
  5. `reduce_step()` all-reduces {loss_sum, count, dL/dw, dL/dpose_deltas} in ONE buffer and normalises.

`exchange()` runs on CPU tensors with the gloo backend as well (that is how tests/test_parallel.py covers it
with world_size 2); the search and the fused step need the GPU.
"""
import torch
import torch.distributed as dist

from .depth_cloud import DepthCloud

__all__ = ['LocalMap', 'SlabPartitioner', 'reduce_step']

N_HIST_BINS = 1 << 14


class LocalMap(object):
    """What one rank holds after the exchange."""

    def __init__(self, clouds, scan_ids, owned, global_ids, axis, bounds):
        self.clouds = clouds            # DepthCloud per locally present scan (owned + halo points)
        self.scan_ids = scan_ids        # int64 [S_local] global scan id of every local cloud
        self.owned = owned              # bool [N_local] in concatenated order: this rank owns the loss term
        self.global_ids = global_ids    # int64 [N_local, 2] = (global scan id, row inside that scan)
        self.axis = axis
        self.bounds = bounds            # (lo, hi) of this rank's slab along `axis`

    def __len__(self):
        return int(self.owned.numel())


class SlabPartitioner(object):
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    # ---- collectives that degrade to no-ops in a single process -------------------------------------
    def _all_reduce(self, t, op=None):
        if self.world > 1:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=self.group)
        return t

    def _all_to_all(self, send, send_counts, recv_counts):
        out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
        if self.world > 1:
            dist.all_to_all_single(out, send.contiguous(), list(recv_counts), list(send_counts), group=self.group)
        else:
            out.copy_(send)
        return out

    # ---- 2. slab boundaries -----------------------------------------------------------------------
    def plan(self, world_points):
        """world_points: list of [n,3] tensors (this rank's scans in the initial map frame).
        Returns (axis, boundaries float64 [G+1]) with boundaries[0] = -inf, boundaries[G] = +inf."""
        dev = world_points[0].device if world_points else torch.device('cpu')
        pts = torch.cat([p.detach().reshape(-1, 3) for p in world_points]).double() if world_points else \
            torch.zeros((0, 3), dtype=torch.float64, device=dev)
        big = 1e300
        lo = pts.min(dim=0).values if len(pts) else torch.full((3,), big, dtype=torch.float64, device=dev)
        hi = pts.max(dim=0).values if len(pts) else torch.full((3,), -big, dtype=torch.float64, device=dev)
        lo = self._all_reduce(lo.clone(), dist.ReduceOp.MIN)
        hi = self._all_reduce(hi.clone(), dist.ReduceOp.MAX)
        axis = int(torch.argmax(hi - lo).item())
        a0, a1 = float(lo[axis]), float(hi[axis])
        width = max(a1 - a0, 1e-12)
        bins = ((pts[:, axis] - a0) * (N_HIST_BINS / width)).long().clamp_(0, N_HIST_BINS - 1)
        hist = torch.bincount(bins, minlength=N_HIST_BINS).double()
        hist = self._all_reduce(hist)
        cum = torch.cumsum(hist, dim=0)
        total = cum[-1]
        targets = total * torch.arange(1, self.world, dtype=torch.float64, device=dev) / self.world
        cut = torch.searchsorted(cum, targets)                       # first bin reaching the target count
        inner = a0 + (cut.double() + 1.0) * (width / N_HIST_BINS)
        inf = torch.tensor([float('inf')], dtype=torch.float64, device=dev)
        return axis, torch.cat([-inf, inner, inf])

    # ---- 3. exchange --------------------------------------------------------------------------------
    def exchange(self, clouds, scan_ids, world_points, axis, boundaries, halo):
        """Route point records to slab owners (+ halo copies).

        clouds: this rank's per-scan DepthClouds (vps, dirs, depth, inc_angles, mask); scan_ids: their global
        ids; world_points: their points in the initial map frame.  Returns a LocalMap.
        """
        assert len(clouds) == len(scan_ids) == len(world_points)
        G = self.world
        dev = clouds[0].depth.device if clouds else torch.device('cpu')
        dt = clouds[0].depth.dtype if clouds else torch.float32
        frecs, irecs, coords = [], [], []
        for c, sid, wp in zip(clouds, scan_ids, world_points):
            n = len(c)
            inc = c.inc_angles if c.inc_angles is not None else torch.zeros((n, 1), dtype=dt, device=dev)
            mask = c.mask if c.mask is not None else torch.ones(n, dtype=torch.bool, device=dev)
            frecs.append(torch.cat([c.vps.expand(n, 3), c.dirs, c.depth, inc.to(dt)], dim=1))
            irecs.append(torch.stack([torch.full((n,), int(sid), dtype=torch.int64, device=dev),
                                      torch.arange(n, dtype=torch.int64, device=dev), mask.long()], dim=1))
            coords.append(wp.detach().reshape(-1, 3)[:, axis].double())
        frec = torch.cat(frecs) if frecs else torch.zeros((0, 8), dtype=dt, device=dev)
        irec = torch.cat(irecs) if irecs else torch.zeros((0, 3), dtype=torch.int64, device=dev)
        x = torch.cat(coords) if coords else torch.zeros(0, dtype=torch.float64, device=dev)
        owner = torch.bucketize(x, boundaries[1:-1].to(dev), right=True)          # slab g: b[g] <= x < b[g+1]
        send_f, send_i, send_counts = [], [], []
        for g in range(G):
            lo, hi = float(boundaries[g]), float(boundaries[g + 1])
            member = (x >= lo - halo) & (x < hi + halo)
            sel = member.nonzero().squeeze(1)
            send_f.append(frec[sel])
            send_i.append(torch.cat([irec[sel], (owner[sel] == g).long()[:, None]], dim=1))
            send_counts.append(int(sel.numel()))
        counts = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        recv = torch.empty_like(counts)
        if G > 1:
            dist.all_to_all_single(recv, counts, group=self.group)
        else:
            recv.copy_(counts)
        recv_counts = recv.tolist()
        rf = self._all_to_all(torch.cat(send_f), send_counts, recv_counts)
        ri = self._all_to_all(torch.cat(send_i), send_counts, recv_counts)
        # group by (scan id, row): scans become contiguous and keep their original point order
        key = ri[:, 0] * (int(ri[:, 1].max().item()) + 1 if len(ri) else 1) + ri[:, 1]
        order = torch.argsort(key, stable=True)
        rf, ri = rf[order], ri[order]
        sids, sizes = torch.unique_consecutive(ri[:, 0], return_counts=True)
        local_clouds, first = [], 0
        for n in sizes.tolist():
            f = rf[first:first + n]
            local_clouds.append(DepthCloud(vps=f[:, 0:3].contiguous(), dirs=f[:, 3:6].contiguous(), depth=f[:, 6:7].contiguous(),
                                           inc_angles=f[:, 7:8].contiguous(), mask=ri[first:first + n, 2].bool()))
            first += n
        return LocalMap(local_clouds, sids, ri[:, 3].bool(), ri[:, :2].contiguous(), axis,
                        (float(boundaries[self.rank]), float(boundaries[self.rank + 1])))


def reduce_step(sum_count, params, group=None):
    """5. Backward of the local loss sum, then ONE all-reduce of {loss_sum, count, gradients}.

    sum_count: tensor [2] from the fused loss (sum over owned points, number of owned points) with autograd
    history; params: tensors whose .grad the step fills (model.w, pose deltas ...).  Returns the global mean
    loss; every rank ends with identical, globally normalised .grad tensors."""
    sum_count[0].backward()
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    buf = torch.cat([sum_count.detach().reshape(-1).double()] + [g.reshape(-1).double() for g in grads])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, group=group)
    count = buf[1]
    off = 2
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = (buf[off:off + n] / count).reshape(g.shape).to(g.dtype)
        off += n
    return buf[0] / count
