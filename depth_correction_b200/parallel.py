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

__all__ = ['LocalMap', 'SlabPartitioner', 'distributed_quantile', 'reduce_step', 'sharded_inlier_sum_count']

N_HIST_BINS = 1 << 14
SCAN_ID_CAP = 4096          # global scan ids the GPU exchange counts per destination (more scans: torch form)


class LocalMap(object):
    """What one rank holds after the exchange."""

    def __init__(self, clouds, scan_ids, owned, global_ids, axis, bounds):
        self.clouds = clouds            # DepthCloud per locally present scan (owned + halo points)
        self.scan_ids = scan_ids        # int64 [S_local] global scan id of every local cloud
        self.owned = owned              # bool [N_local] in concatenated order: this rank owns the loss term
        self.global_ids = global_ids    # int64 [N_local, 2] = (global scan id, row inside that scan)
        self.axis = axis
        self.bounds = bounds            # (lo, hi) of this rank's slab along `axis` (floats, or a device tensor [2])

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
        """world_points: list of [n,3] tensors (this rank's scans in the initial map frame) or one concatenated
        [N,3] tensor.  Returns (axis, boundaries float64 [G+1]) with boundaries[0] = -inf, boundaries[G] = +inf."""
        if isinstance(world_points, torch.Tensor):
            dev = world_points.device
            pts = world_points.detach().reshape(-1, 3).double()
        else:
            dev = world_points[0].device if world_points else torch.device('cpu')
            pts = torch.cat([p.detach().reshape(-1, 3) for p in world_points]).double() if world_points else \
                torch.zeros((0, 3), dtype=torch.float64, device=dev)
        big = 1e300
        on_gpu = pts.is_cuda and len(pts) > 0
        if on_gpu:
            # bounding box and histogram by two kernels of the library (one pass each) instead of a dozen torch passes
            from . import _lib as L
            pts = pts.contiguous()
            box = torch.empty(6, dtype=torch.float64, device=dev)
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            L.call('dc_bounds', L.ptr(pts), L.DC_F64, pts.shape[0], L.ptr(box), L.ptr(bad), L.stream())
            lo, hi = box[:3], box[3:]
        else:
            lo = pts.min(dim=0).values if len(pts) else torch.full((3,), big, dtype=torch.float64, device=dev)
            hi = pts.max(dim=0).values if len(pts) else torch.full((3,), -big, dtype=torch.float64, device=dev)
        lo = self._all_reduce(lo.clone(), dist.ReduceOp.MIN)
        hi = self._all_reduce(hi.clone(), dist.ReduceOp.MAX)
        lohi = torch.cat([lo, hi]).tolist()
        ext = [lohi[3 + a] - lohi[a] for a in range(3)]
        axis = max(range(3), key=lambda a: (ext[a], -a))                # first axis of the largest extent, like argmax
        a0, a1 = lohi[axis], lohi[3 + axis]
        width = max(a1 - a0, 1e-12)
        if on_gpu:
            hist32 = torch.empty(N_HIST_BINS, dtype=torch.int32, device=dev)
            L.call('dc_axis_histogram', L.ptr(pts), axis, pts.shape[0], a0, N_HIST_BINS / width, N_HIST_BINS, L.ptr(hist32), L.stream())
            hist = hist32.double()
        else:
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
        ids; world_points: their points in the initial map frame (list per scan, or one concatenated [N,3] tensor).
        Returns a LocalMap.
        """
        assert len(clouds) == len(scan_ids)
        G = self.world
        if clouds and clouds[0].depth.is_cuda and G <= 64 and max(int(s) for s in scan_ids) < SCAN_ID_CAP:
            return self._exchange_cuda(clouds, scan_ids, world_points, axis, boundaries, halo)
        dev = clouds[0].depth.device if clouds else torch.device('cpu')
        dt = clouds[0].depth.dtype if clouds else torch.float32
        sizes = [len(c) for c in clouds]
        n = sum(sizes)
        # one batched op per field (a training run calls this once, but with hundreds of scans per rank a loop of
        # small per-scan ops costs tens of milliseconds of launches)
        if clouds:
            vps = torch.cat([c.vps.expand(m, 3) for c, m in zip(clouds, sizes)])
            dirs = torch.cat([c.dirs for c in clouds])
            depth = torch.cat([c.depth.reshape(-1, 1) for c in clouds])
            inc = torch.cat([c.inc_angles.reshape(-1, 1).to(dt) if c.inc_angles is not None
                             else torch.zeros((m, 1), dtype=dt, device=dev) for c, m in zip(clouds, sizes)])
            frec = torch.cat([vps.to(dt), dirs, depth, inc], dim=1)
            mask = torch.cat([c.mask if c.mask is not None else torch.ones(m, dtype=torch.bool, device=dev)
                              for c, m in zip(clouds, sizes)])
            sz = torch.as_tensor(sizes, dtype=torch.int64, device=dev)        # (record columns below are int32: half the traffic)
            first = torch.cumsum(sz, 0) - sz
            sid = torch.repeat_interleave(torch.as_tensor([int(s) for s in scan_ids], dtype=torch.int64, device=dev), sz)
            row = torch.arange(n, dtype=torch.int64, device=dev) - torch.repeat_interleave(first, sz)
            irec = torch.stack([sid.int(), row.int(), mask.int()], dim=1)
            if isinstance(world_points, torch.Tensor):
                x = world_points.detach().reshape(-1, 3)[:, axis].double().contiguous()
            else:
                x = torch.cat([wp.detach().reshape(-1, 3)[:, axis] for wp in world_points]).double()
            assert x.numel() == n
        else:
            frec = torch.zeros((0, 8), dtype=dt, device=dev)
            irec = torch.zeros((0, 3), dtype=torch.int32, device=dev)
            x = torch.zeros(0, dtype=torch.float64, device=dev)
        inner = boundaries[1:-1].to(dev).contiguous()
        owner = torch.bucketize(x, inner, right=True)                               # slab g: b[g] <= x < b[g+1]
        # a point is sent to every slab within `halo` of it: g_min .. g_max (usually just its owner), i.e. the slabs
        # with b[g] - halo <= x < b[g+1] + halo.  One stable sort by destination replaces a pass per destination.
        g_max = torch.bucketize(x + halo, inner, right=True)
        g_min = torch.bucketize(x - halo, inner, right=True)
        copies = g_max - g_min + 1
        src = torch.repeat_interleave(torch.arange(n, dtype=torch.int64, device=dev), copies)
        start = torch.cumsum(copies, 0) - copies
        dest = g_min[src] + (torch.arange(src.numel(), dtype=torch.int64, device=dev) - start[src])
        by_dest = torch.argsort(dest, stable=True)                                  # keeps (scan, row) order inside a destination
        src, dest = src[by_dest], dest[by_dest]
        send_counts = torch.bincount(dest, minlength=G)
        send_f = frec[src]
        send_i = torch.cat([irec[src], (owner[src] == dest).to(irec.dtype)[:, None]], dim=1)
        recv = torch.empty_like(send_counts)
        if G > 1:
            dist.all_to_all_single(recv, send_counts, group=self.group)
        else:
            recv.copy_(send_counts)
        both = torch.stack([send_counts, recv]).tolist()                            # the one host read-back before the exchange
        send_counts, recv_counts = both[0], both[1]
        rf = self._all_to_all(send_f, send_counts, recv_counts)
        ri = self._all_to_all(send_i, send_counts, recv_counts)
        # group by (scan id, row): scans become contiguous and keep their original point order
        key = ri[:, 0].long() * (int(ri[:, 1].max().item()) + 1 if len(ri) else 1) + ri[:, 1].long()
        order = torch.argsort(key, stable=True)
        rf, ri = rf[order], ri[order]
        sids, sizes_l = torch.unique_consecutive(ri[:, 0], return_counts=True)
        # split the record matrix into per-field arrays once; the per-scan clouds are contiguous row slices of them
        f_vps, f_dirs = rf[:, 0:3].contiguous(), rf[:, 3:6].contiguous()
        f_depth, f_inc = rf[:, 6:7].contiguous(), rf[:, 7:8].contiguous()
        f_mask = ri[:, 2].bool()
        local_clouds, first = [], 0
        for m in sizes_l.tolist():
            local_clouds.append(DepthCloud(vps=f_vps[first:first + m], dirs=f_dirs[first:first + m], depth=f_depth[first:first + m],
                                           inc_angles=f_inc[first:first + m], mask=f_mask[first:first + m]))
            first += m
        return LocalMap(local_clouds, sids.long(), ri[:, 3].bool(), ri[:, :2].long().contiguous(), axis,
                        (float(boundaries[self.rank]), float(boundaries[self.rank + 1])))

    def _exchange_cuda(self, clouds, scan_ids, world_points, axis, boundaries, halo):
        """GPU form of SlabPartitioner.exchange: routing, packing and unpacking in four kernels (dc_route_*), one small
        and two payload all-to-alls, and ONE host read-back (the send / receive counts NCCL needs on the host; the same
        transfer carries the size of every scan this rank will hold, counted per (destination, scan) while routing).
        The slab boundaries stay on the device.  Same result as the torch form above (which the CPU / gloo tests
        exercise) up to the arbitrary order inside a destination, which the (scan, row) sort on the receiving side removes."""
        from . import _lib as L
        from .fused import scan_table
        G = self.world
        dev = clouds[0].depth.device
        dt = clouds[0].depth.dtype
        code = L.dtype_code(dt)
        st = L.stream()
        n = sum(len(c) for c in clouds)
        wp = world_points if isinstance(world_points, torch.Tensor) else torch.cat([w.detach().reshape(-1, 3) for w in world_points])
        wp = wp.detach().reshape(-1, 3).to(torch.float64).contiguous()
        assert wp.shape[0] == n
        tbl, first, _keep = scan_table(clouds, dt)
        ids_host = [int(s) for s in scan_ids]
        sid_t = L.upload(ids_host, torch.int32, dev)
        S_cap = SCAN_ID_CAP
        assert max(ids_host) < S_cap, 'scan ids must be below %d' % S_cap
        inner = boundaries[1:-1].to(device=dev, dtype=torch.float64).contiguous()
        if inner.numel() == 0:
            inner = torch.zeros(1, dtype=torch.float64, device=dev)
        gmin = torch.empty(n, dtype=torch.uint8, device=dev)
        gmax = torch.empty(n, dtype=torch.uint8, device=dev)
        # header row g = {records for destination g, records of every scan for destination g}
        header = torch.empty((G, 1 + S_cap), dtype=torch.int32, device=dev)
        counts = torch.empty(G, dtype=torch.int32, device=dev)
        scan_counts = torch.empty((G, S_cap), dtype=torch.int32, device=dev)
        L.call('dc_route_count', L.ptr(wp), int(axis), n, L.ptr(inner), G, float(halo), L.ptr(first), L.ptr(sid_t), len(clouds),
               L.ptr(gmin), L.ptr(gmax), L.ptr(counts), L.ptr(scan_counts), S_cap, st)
        header[:, 0] = counts
        header[:, 1:] = scan_counts
        recv_header = torch.empty_like(header)
        if G > 1:
            dist.all_to_all_single(recv_header, header, group=self.group)
        else:
            recv_header.copy_(header)
        # THE host read-back: send counts, receive counts, sizes of the scans this rank will hold
        scan_sizes = recv_header[:, 1:].sum(dim=0, dtype=torch.int64)
        host = torch.cat([counts.long(), recv_header[:, 0].long(), scan_sizes]).cpu()
        send_counts, recv_counts = host[:G].tolist(), host[G:2 * G].tolist()
        sizes_all = host[2 * G:]
        present = torch.nonzero(sizes_all)[:, 0]
        sids_host, sizes_l = present.tolist(), sizes_all[present].tolist()
        offs = [0]
        for c in send_counts[:-1]:
            offs.append(offs[-1] + c)
        m_send = sum(send_counts)
        dest_offset = L.upload(offs, torch.int64, dev)
        cursor = torch.empty(G, dtype=torch.int32, device=dev)
        send_f = torch.empty((m_send, 8), dtype=dt, device=dev)
        send_i = torch.empty((m_send, 4), dtype=torch.int32, device=dev)
        L.call('dc_route_pack', L.ptr(tbl), L.ptr(first), L.ptr(sid_t), len(clouds), n, code, L.ptr(wp), int(axis), L.ptr(inner), G, float(halo),
               L.ptr(gmin), L.ptr(gmax), L.ptr(dest_offset), L.ptr(cursor), L.ptr(send_f), L.ptr(send_i), st)
        rf = self._all_to_all(send_f, send_counts, recv_counts)
        ri = self._all_to_all(send_i, send_counts, recv_counts)
        m = rf.shape[0]
        assert m == sum(sizes_l)
        keys = torch.empty(m, dtype=torch.int64, device=dev)
        ids = torch.empty(m, dtype=torch.int32, device=dev)
        skeys = torch.empty(m, dtype=torch.int64, device=dev)
        order = torch.empty(m, dtype=torch.int32, device=dev)
        f_vps = torch.empty((m, 3), dtype=dt, device=dev)
        f_dirs = torch.empty((m, 3), dtype=dt, device=dev)
        f_depth = torch.empty((m, 1), dtype=dt, device=dev)
        f_inc = torch.empty((m, 1), dtype=dt, device=dev)
        f_mask = torch.empty(m, dtype=torch.bool, device=dev)
        f_owned = torch.empty(m, dtype=torch.bool, device=dev)
        gid = torch.empty((m, 2), dtype=torch.int64, device=dev)
        if m > 0:
            L.call('dc_route_keys', L.ptr(ri), m, L.ptr(keys), L.ptr(ids), st)
            # keys = scan id << 32 | row: only the bits that can be set are sorted
            key_bits = 32 + max(1, int(max(sids_host)).bit_length())
            L.call_with_temp('dc_sort_pairs', dev, L.ptr(keys), L.ptr(skeys), L.ptr(ids), L.ptr(order), m, key_bits, after=(st,))
            L.call('dc_route_unpack', L.ptr(rf), L.ptr(ri), L.ptr(order), m, code, L.ptr(f_vps), L.ptr(f_dirs), L.ptr(f_depth), L.ptr(f_inc),
                   L.ptr(f_mask.view(torch.uint8)), L.ptr(f_owned.view(torch.uint8)), L.ptr(gid), st)
        local_clouds, first_row = [], 0
        for sz in sizes_l:
            local_clouds.append(DepthCloud(vps=f_vps[first_row:first_row + sz], dirs=f_dirs[first_row:first_row + sz],
                                           depth=f_depth[first_row:first_row + sz], inc_angles=f_inc[first_row:first_row + sz],
                                           mask=f_mask[first_row:first_row + sz]))
            first_row += sz
        sids = L.upload(sids_host, torch.int64, dev)
        return LocalMap(local_clouds, sids, f_owned, gid, axis, boundaries[self.rank:self.rank + 2])


def reduce_step(sum_count, params, group=None):
    """5. Backward of the local loss sum, then ONE all-reduce of {loss_sum, count, gradients}.

    sum_count: tensor [2] from the fused loss (sum over owned points, number of owned points) with autograd
    history; params: tensors whose .grad the step fills (model.w, pose deltas ...).  Returns the global mean
    loss; every rank ends with identical, globally normalised .grad tensors."""
    # gradients of THIS step only (torch.autograd.grad, not .backward(): whatever a caller left in p.grad must not be
    # summed into the all-reduce and divided by the count)
    got = torch.autograd.grad(sum_count[0], list(params), allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(got, params)]
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


def distributed_quantile(x, q, group=None, max_gather=4096):
    """torch.quantile(cat_over_ranks(x), q) (linear interpolation, 1-D, NaN-free) without gathering the data:
    every rank holds a shard `x`; two all-reduced 2^16-bin histogram rounds narrow the two order statistics the
    result interpolates between down to a few candidates, which are then gathered and sorted.  Works on any
    device / backend (the CPU tests run it over gloo).  This is the global inlier threshold of
    min_eigval_loss / trace_loss (`inlier_ratio < 1`, loss.py:256-267) for a map sharded over GPUs."""
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    x = x.detach().reshape(-1).double()
    dev = x.device

    def allsum(t):
        if multi:
            dist.all_reduce(t, group=group)
        return t

    n = int(allsum(torch.tensor([x.numel()], dtype=torch.float64, device=dev)).item())
    assert n > 0, 'quantile of an empty set'
    rank = q * (n - 1)
    k_lo = int(rank)
    k_hi = min(k_lo + 1, n - 1) if rank > k_lo else k_lo
    big = torch.tensor([float('inf')], dtype=torch.float64, device=dev)
    lo = torch.min(x) if x.numel() else big[0]
    hi = torch.max(x) if x.numel() else -big[0]
    mm = torch.stack([-lo, hi])
    if multi:
        dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
    lo, hi = -mm[0], mm[1]

    def order_statistic(k):
        """k-th smallest (0-based) of the union."""
        a, b, below = lo.clone(), hi.clone(), 0           # the statistic lies in [a, b]; `below` values are < a
        cur = x
        for _ in range(6):
            cnt = int(allsum(torch.tensor([cur.numel()], dtype=torch.float64, device=dev)).item())
            if cnt <= max_gather or not (b > a):
                break
            nb = 1 << 16
            width = (b - a) / nb
            bins = ((cur - a) / width).long().clamp_(0, nb - 1)
            hist = allsum(torch.bincount(bins, minlength=nb).double())
            cum = torch.cumsum(hist, 0)
            bsel = int(torch.searchsorted(cum, torch.tensor([float(k - below) + 0.5], dtype=torch.float64, device=dev)).item())
            below += int(cum[bsel - 1].item()) if bsel > 0 else 0
            cur = cur[bins == bsel]
            a, b = a + bsel * width, a + (bsel + 1) * width
        # gather the remaining candidates (equal sizes via padding with +inf)
        m = torch.tensor([cur.numel()], dtype=torch.int64, device=dev)
        if multi:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        m = int(m.item())
        pad = torch.full((m,), float('inf'), dtype=torch.float64, device=dev)
        pad[:cur.numel()] = cur
        if multi:
            parts = [torch.empty_like(pad) for _ in range(dist.get_world_size(group))]
            dist.all_gather(parts, pad, group=group)
            pad = torch.cat(parts)
        return torch.sort(pad).values[k - below]

    v_lo = order_statistic(k_lo)
    v_hi = order_statistic(k_hi) if k_hi != k_lo else v_lo
    return torch.lerp(v_lo, v_hi, rank - k_lo)


def sharded_inlier_sum_count(cloud, owned, inlier_ratio=1.0, inlier_loss_mult=1.0, inlier_max_loss=None,
                             loss='min_eigval_loss', sqrt=False, normalization=False, group=None):
    """Multi-GPU form of the inlier-selecting losses (loss.py:256-293 / 332-369): the quantile of the raw per-point
    loss is taken over the OWNED points of all ranks (distributed_quantile), points above the threshold are dropped,
    and the local (sum, count) of the survivors is returned with autograd history for `reduce_step`."""
    from . import _lib as L
    from .fused import fused_loss
    from .loss import _kind_flags, min_eigval_loss, trace_loss
    loss_fun = trace_loss if loss in ('trace_loss', trace_loss) else min_eigval_loss
    kind, flags = _kind_flags(loss_fun, dict(sqrt=False, normalization=normalization))
    raw = fused_loss(cloud.step_state(), cloud._model, cloud.poses_tensor(), kind, flags | L.FLAG_RAW, mask=None)[owned]
    thr = None
    if inlier_ratio < 1.0:
        thr = inlier_loss_mult * distributed_quantile(raw, inlier_ratio, group=group)
    if inlier_max_loss is not None:
        cap = torch.as_tensor(inlier_max_loss, dtype=raw.dtype, device=raw.device)
        thr = cap if thr is None else torch.min(cap, thr)
    if thr is not None:
        raw = raw[raw.detach() <= thr]
    val = torch.relu(raw)
    if sqrt:
        val = torch.sqrt(val)
    return torch.stack([val.sum(), torch.as_tensor(float(val.numel()), dtype=val.dtype, device=val.device)])
