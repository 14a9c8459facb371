"""Fused fixed-graph map-consistency step (kernels 2 and 3) behind torch.autograd.

One `StepState` holds everything that is constant over the optimisation loop of
scripts/model_poses_learning:119-135 / train.py:225-312 for one global cloud: the cell-sorted
packed scan records, the neighbourhood graph (and its transpose), and the scratch buffers the
kernels write (corrected points, backward stash, per-point loss).  `fused_loss` runs

    model(cloud) -> transform(pose) -> concatenate -> update_points/mean/cov/eig -> loss   (forward)
    d loss / d {w, exponent, poses}                                                        (backward)

as three kernel launches forward+backward-prologue and one backward launch.
"""
import os
import weakref

import torch

from . import _lib as L

__all__ = ['StepState', 'fused_loss', 'model_kind_of']

TRANSPOSE_AFTER = 2                     # backward passes served by the scatter form before a kNN graph is transposed
# The scatter form accumulates dL/dp in float32 (one vector reduction per edge instead of three fp64 ones, 3x faster)
# on maps of at least this many points: there the fp32 rounding of the per-point sums (random, 6e-8 relative per
# addition) averages out over the points the chain stage adds up in fp64 (measured: gradients agree with the fp64
# gather form to ~1e-7 on the bench map and to < 1e-6 on a 0.5 M point map,
# tests/test_gpu_parity.py::test_scatter_f32_agrees_with_gather_form).
# DC_SCATTER_F32=0 / 1 forces the choice; DC_BACKWARD=gather forces the deterministic fp64 gather form (bitwise
# reproducible gradients), DC_BACKWARD=scatter the scatter form.
SCATTER_F32_MIN_POINTS = 1 << 19
# 'auto' (policy above) | 'gather' (fp64, atomic-free, bitwise reproducible gradients) | 'scatter'.  Set through
# set_backward_form() / Config.backward_form; the environment variable DC_BACKWARD only provides the initial value.
BACKWARD_FORM = os.environ.get('DC_BACKWARD', 'auto')
_form_logged = set()


def set_backward_form(form):
    """Select how dL/dp is accumulated: 'auto', 'gather' (deterministic fp64) or 'scatter' (L2 reductions; float32 on
    maps of >= 2^19 points).  The reference's autograd is deterministic: use 'gather' for bit-reproducible runs."""
    global BACKWARD_FORM
    assert form in ('auto', 'gather', 'scatter'), form
    BACKWARD_FORM = form


def _scatter_f32(n):
    force = os.environ.get('DC_SCATTER_F32')
    return (n >= SCATTER_F32_MIN_POINTS) if force is None else force == '1'


CHAIN_CHUNK = 2048                      # rows per block of the chain stage (256 threads x 8)
CHAIN_REC = 12 + 2 * L.MAX_TERMS        # doubles per partial record (dc_step.cu)


def model_kind_of(model):
    if model is None:
        return L.MODEL_NONE
    name = type(model).__name__
    if name == 'ScaledPolynomial':
        return L.MODEL_SCALED_POLYNOMIAL
    if name == 'Polynomial':
        return L.MODEL_POLYNOMIAL
    if name == 'BaseModel':
        return L.MODEL_NONE
    raise NotImplementedError('fused step supports Polynomial / ScaledPolynomial models, got %s' % name)


# key -> (tbl, first): two entries (the search and the step of one iteration use the same scans).  Only tables that
# point straight at the caller's tensors are cached, and the cache holds NO reference to them: the key is made of the
# very addresses and lengths the table contains, so a hit is valid whatever happened to the objects in between
# (a cache that kept the tensors alive made every new set of scans allocate fresh device memory: 16 -> 60 ms e2e).
_scan_table_cache = {}


def _plain(t, dt, shape_tail):
    """The tensor itself when it already is a contiguous `dt` array of the expected shape (no torch op at all)."""
    return t.dtype == dt and t.is_contiguous() and tuple(t.shape[1:]) == shape_tail


def scan_table(clouds, dt):
    """Device table of per-scan tensor addresses for the batched kernels: (tbl int64 [S,5] = {vps, dirs, depth,
    inc_angles, model mask} (0 = absent), first int64 [S+1], keep-alive list of the tensors the table points to).
    Host cost matters here (it sits in front of the search): tensors that are already in kernel layout are used as
    they are, and the table of an unchanged list of scans is reused."""
    key = (dt,) + tuple((len(c), c.depth.data_ptr(), c.dirs.data_ptr(), c.vps.data_ptr(), tuple(c.vps.shape),
                         0 if c.inc_angles is None else c.inc_angles.data_ptr(),
                         0 if c.mask is None else c.mask.data_ptr()) for c in clouds)
    hit = _scan_table_cache.get(key)
    if hit is not None:
        return hit[0], hit[1], []
    rows, keep, first = [], [], [0]
    all_plain = True
    dev = clouds[0].depth.device
    for c in clouds:
        cnt = len(c)
        assert c.depth.dtype == dt and c.dirs.is_cuda
        dirs = c.dirs
        if not _plain(dirs, dt, (3,)):
            dirs, all_plain = dirs.detach().reshape(-1, 3).contiguous(), False
        vps = c.vps
        if not (_plain(vps, dt, (3,)) and vps.shape[0] == cnt):
            vps = vps.detach()
            vps = vps.to(dt).contiguous() if vps.shape[0] == cnt else vps.to(dt).expand(cnt, 3).contiguous()
            all_plain = False
        depth = c.depth
        if not (_plain(depth, dt, (1,)) or _plain(depth, dt, ())):
            depth, all_plain = depth.detach().reshape(-1).contiguous(), False
        inc = c.inc_angles
        if inc is not None and not (_plain(inc, dt, (1,)) or _plain(inc, dt, ())):
            inc, all_plain = inc.detach().reshape(-1).to(dt).contiguous(), False
        mm = c.mask
        if mm is not None:
            # bool storage is one 0/1 byte per element: reinterpret instead of converting
            if mm.dtype == torch.bool and mm.is_contiguous():
                mm = mm.view(torch.uint8)
            else:
                mm, all_plain = mm.detach().to(torch.uint8).contiguous(), False
        keep += [dirs, vps, depth, inc, mm]
        rows.append([vps.data_ptr(), dirs.data_ptr(), depth.data_ptr(), 0 if inc is None else inc.data_ptr(),
                     0 if mm is None else mm.data_ptr()])
        first.append(first[-1] + cnt)
    both = L.upload(rows + [[f, 0, 0, 0, 0] for f in first], torch.int64, dev)      # one copy for both tables
    tbl = both[:len(rows)]
    first_t = both[len(rows):, 0].contiguous()
    if all_plain:
        if len(_scan_table_cache) >= 2:
            _scan_table_cache.pop(next(iter(_scan_table_cache)))
        _scan_table_cache[key] = (tbl, first_t)
    return tbl, first_t, keep


class StepState(object):
    def __init__(self, graph, clouds):
        """graph: Graph over the initial global cloud; clouds: per-scan local DepthClouds (in scan order)."""
        smap = graph.map
        dev = smap.device
        n = smap.n
        assert graph.self_query and graph.n_rows == n
        sizes = [len(c) for c in clouds]
        assert sum(sizes) == n, 'clouds (%i points) do not match the graph (%i points)' % (sum(sizes), n)
        # the graph owns this state (graph._step_cache); a weak reference back avoids a reference cycle, so that
        # dropping the graph frees its device memory immediately instead of at the next cyclic GC
        self._graph_ref = weakref.ref(graph)
        self.n = n
        self.n_scans = len(clouds)
        self.device = dev
        dt = clouds[0].depth.dtype
        self.dtype = dt
        self.code = L.dtype_code(dt)
        st = L.stream()
        self.rec_dir = torch.empty((n, 4), dtype=dt, device=dev)
        self.rec_vp = torch.empty((n, 4), dtype=dt, device=dev)
        self.rec_meta = torch.empty(n, dtype=torch.int32, device=dev)
        # second copy of the records in the caller's (scan-major) order for the chain stage of the backward pass
        self.rec_dir_o = torch.empty((n, 4), dtype=dt, device=dev)
        self.rec_vp_o = torch.empty((n, 4), dtype=dt, device=dev)
        self.rec_meta_o = torch.empty(n, dtype=torch.int32, device=dev)
        # one launch for all scans (temporaries are released to the caching allocator in stream order)
        tbl, first, _keep = scan_table(clouds, dt)
        L.call('dc_pack_records_batched', L.ptr(tbl), L.ptr(first), len(clouds), n, self.code, L.ptr(smap.inv_order),
               L.ptr(self.rec_dir), L.ptr(self.rec_vp), L.ptr(self.rec_meta), L.ptr(self.rec_dir_o), L.ptr(self.rec_vp_o),
               L.ptr(self.rec_meta_o), st)
        self.has_inc = all(c.inc_angles is not None for c in clouds)
        self.P = torch.empty((n, 4), dtype=torch.float64, device=dev)
        self.stash = torch.empty((n, 8), dtype=torch.float64, device=dev)
        self.loss_pp = torch.empty(n, dtype=torch.float64, device=dev)
        self.g = torch.empty((n, 3), dtype=torch.float64, device=dev)       # dL/dp per point (fp64 forms)
        self._g32 = None                                                    # float32 [n,4] accumulator of the fp32 scatter form
        self.n_blocks = ((n + 31) // 32 * 32 + 127) // 128
        self.partials = torch.zeros(2 * self.n_blocks + 2, dtype=torch.float64, device=dev)
        # block table of the chain stage: blocks are aligned to scans (CHAIN_CHUNK rows each)
        import numpy as np
        sz = np.asarray(sizes, dtype=np.int64)
        nb = (sz + CHAIN_CHUNK - 1) // CHAIN_CHUNK
        scan_first = np.concatenate([[0], np.cumsum(sz)])
        blk_first = np.concatenate([[0], np.cumsum(nb)]).astype(np.int32)
        blk_scan = np.repeat(np.arange(len(sz), dtype=np.int32), nb)
        within = np.arange(int(nb.sum()), dtype=np.int64) - np.repeat(blk_first[:-1].astype(np.int64), nb)
        blk_start = np.repeat(scan_first[:-1], nb) + within * CHAIN_CHUNK
        blk_count = np.minimum(np.repeat(sz, nb) - within * CHAIN_CHUNK, CHAIN_CHUNK).astype(np.int32)
        self.chain_blocks = int(nb.sum())
        nbk = self.chain_blocks
        tab = L.upload(np.concatenate([blk_start, blk_scan.astype(np.int64), blk_count.astype(np.int64),
                                       blk_first.astype(np.int64)]), torch.int64, dev)    # one async copy
        self.blk_start = tab[:nbk]
        self.blk_scan = tab[nbk:2 * nbk].to(torch.int32)
        self.blk_count = tab[2 * nbk:3 * nbk].to(torch.int32)
        self.scan_blk_first = tab[3 * nbk:].to(torch.int32)
        self.chain_partials = torch.empty(max(self.chain_blocks, 1) * CHAIN_REC, dtype=torch.float64, device=dev)
        self._mask_set, self._mask_ref, self._mask_ver = True, None, None      # packed with an all-true loss mask
        self._keyed = []
        self.generation = 0

    @property
    def graph(self):
        g = self._graph_ref()
        assert g is not None, 'the neighbourhood graph of this step state has been released'
        return g

    def set_loss_mask(self, mask):
        """mask: bool [N] in original (concatenated) order or None (= all points).  The upload is skipped only for the
        very same tensor object at the same version (a strong reference is kept, so its address cannot be recycled by
        a different mask of the same shape)."""
        if self._mask_set and mask is self._mask_ref and (mask is None or mask._version == self._mask_ver):
            return
        m8 = None
        if mask is not None:
            assert mask.numel() == self.n
            m8 = mask.detach().to(device=self.device).to(torch.uint8).contiguous()
        L.call('dc_set_loss_mask', L.ptr(m8), self.n, L.ptr(self.graph.map.order), L.ptr(self.rec_meta), L.stream())
        self._mask_set, self._mask_ref, self._mask_ver = True, mask, None if mask is None else mask._version


class _FusedStep(torch.autograd.Function):
    """Output: raw=False -> tensor [2] = (loss_sum, count); raw=True -> per-point raw loss in sorted space."""

    @staticmethod
    def forward(ctx, w, exponent, poses, state, model_kind, loss_kind, flags):
        dev = state.device
        st = L.stream()
        S = state.n_scans
        assert poses.shape[-2:] == (4, 4) and poses.shape[0] == S, 'poses must be [%i,4,4]' % S
        poses12 = poses.detach()[:, :3, :].to(device=dev, dtype=torch.float64).reshape(S, 12).contiguous()
        n_terms = 0
        wv = ev = None
        if model_kind != L.MODEL_NONE:
            assert state.has_inc, 'model correction needs per-scan inc_angles (model.py:251)'
            wv = w.detach().to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
            ev = exponent.detach().to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
            n_terms = wv.numel()
            assert 1 <= n_terms <= L.MAX_TERMS and ev.numel() == n_terms
        g = state.graph
        L.call('dc_step_points', L.ptr(state.rec_dir), L.ptr(state.rec_vp), L.ptr(state.rec_meta), state.code, state.n,
               L.ptr(poses12), S, model_kind, L.ptr(wv), L.ptr(ev), n_terms, L.ptr(state.P), st)
        raw = bool(flags & L.FLAG_RAW)
        loss_sum = None if raw else torch.empty(2, dtype=torch.float64, device=dev)
        # mean / sum reductions on large kNN maps: forward and the float32 backward scatter in ONE kernel (the upstream
        # gradient is a scalar applied after the chain stage); nothing is stashed.  Only when a gradient is wanted.
        form = BACKWARD_FORM
        fused_bwd = (not raw and any(ctx.needs_input_grad[:3]) and not g.symmetric and _scatter_f32(state.n)
                     and form in ('auto', 'scatter') and os.environ.get('DC_FUSE_BWD', '1') == '1')
        if fused_bwd:
            if state._g32 is None:
                state._g32 = torch.empty((state.n, 4), dtype=torch.float32, device=dev)
            L.call('dc_step_forward_scatter', L.ptr(state.P), L.ptr(state.rec_meta), state.n, L.ptr(g.slice_ptr),
                   L.ptr(g.ell_idx), loss_kind, flags, L.ptr(state.loss_pp), L.ptr(state._g32), L.ptr(loss_sum),
                   L.ptr(state.partials), state.partials.numel() * 8, st)
        else:
            L.call('dc_step_forward', L.ptr(state.P), L.ptr(state.rec_meta), state.n, L.ptr(g.slice_ptr), L.ptr(g.ell_idx),
                   loss_kind, flags, L.ptr(state.loss_pp), L.ptr(state.stash), None, L.ptr(loss_sum),
                   L.ptr(state.partials), state.partials.numel() * 8, st)
        ctx.fused_bwd = fused_bwd
        state.generation += 1
        ctx.state, ctx.generation = state, state.generation
        ctx.graph = g                 # keeps the graph (and with it the state's buffers) alive until backward
        ctx.args = (poses12, wv, ev, n_terms, model_kind, raw)
        ctx.shapes = (w.shape if w is not None else None, exponent.shape if exponent is not None else None,
                      poses.shape, poses.dtype, poses.device)
        ctx.exp_grad = exponent is not None and isinstance(exponent, torch.Tensor) and exponent.requires_grad
        if raw:
            return state.loss_pp.clone()
        return loss_sum

    @staticmethod
    def backward(ctx, *grads):
        state = ctx.state
        if state.generation != ctx.generation:
            raise RuntimeError('fused step: backward() after a newer forward() on the same cloud; '
                               'the backward stash has been overwritten (call backward before the next forward)')
        poses12, wv, ev, n_terms, model_kind, raw = ctx.args
        dev = state.device
        st = L.stream()
        S = state.n_scans
        # Transposed graph policy: radius graphs are their own transpose.  For kNN graphs building the reverse
        # lists costs about as much as three backward passes, so the first TRANSPOSE_AFTER backward passes on a
        # graph use the scatter form and the transpose is built once the graph is evidently being reused.
        graph = ctx.graph
        graph._bwd_calls = getattr(graph, '_bwd_calls', 0) + 1
        scatter_f32 = _scatter_f32(state.n)
        form = BACKWARD_FORM                                   # auto | gather | scatter
        if ctx.fused_bwd:
            gt = None                    # dL/dp already sits in state._g32 (dc_step_forward_scatter)
        elif form == 'scatter' or (form == 'auto' and scatter_f32 and not graph.symmetric):
            gt = None                    # large kNN maps: the fp32 scatter form beats the gather form outright (1.4 vs 2.0 ms
                                         # on the bench map) and needs no reverse lists
        elif graph.symmetric:
            gt = graph
        elif form == 'gather' or graph._transposed is not None or graph._bwd_calls > TRANSPOSE_AFTER:
            gt = graph.transposed()
        else:
            gt = None
        dw = torch.zeros(max(n_terms, 1), dtype=torch.float64, device=dev)
        dexp = torch.zeros(max(n_terms, 1), dtype=torch.float64, device=dev) if ctx.exp_grad else None
        dposes = torch.zeros((S, 12), dtype=torch.float64, device=dev)
        upstream = None
        if raw:
            upstream = grads[0].to(torch.float64).contiguous()
        g_index = None
        g_buf, g_code = state.g, L.DC_F64
        chosen = 'fused float32 scatter' if ctx.fused_bwd else ('fp64 gather' if gt is not None else
                                                                 ('float32 scatter' if scatter_f32 else 'fp64 scatter'))
        if chosen not in _form_logged and os.environ.get('DC_VERBOSE'):
            _form_logged.add(chosen)
            print('depth_correction_b200: backward form "%s" (n = %d, fused.set_backward_form)' % (chosen, state.n))
        if ctx.fused_bwd:
            g_buf, g_code = state._g32, L.DC_F32
            g_index = graph.map.inv_order
        elif gt is not None:
            # gather form over the transposed graph: atomic-free, deterministic
            L.call('dc_step_backward', L.ptr(state.P), state.n, L.ptr(gt.slice_ptr), L.ptr(gt.ell_idx), L.ptr(state.stash),
                   L.ptr(upstream), L.ptr(graph.map.order), L.ptr(state.g), st)
        else:
            # scatter form over the forward graph (L2 reductions): no transpose needed yet
            if scatter_f32:
                if state._g32 is None:
                    state._g32 = torch.empty((state.n, 4), dtype=torch.float32, device=dev)
                g_buf, g_code = state._g32, L.DC_F32
            g_buf.zero_()
            L.call('dc_step_backward_scatter', L.ptr(state.P), state.n, L.ptr(graph.slice_ptr), L.ptr(graph.ell_idx),
                   L.ptr(state.stash), L.ptr(upstream), L.ptr(g_buf), g_code, st)
            g_index = graph.map.inv_order
        L.call('dc_step_chain', L.ptr(g_buf), g_code, L.ptr(g_index), L.ptr(state.rec_dir_o), L.ptr(state.rec_vp_o), L.ptr(state.rec_meta_o),
               state.code, L.ptr(state.blk_scan), L.ptr(state.blk_start), L.ptr(state.blk_count), state.chain_blocks,
               L.ptr(state.scan_blk_first), L.ptr(poses12), S, model_kind, L.ptr(wv), L.ptr(ev), n_terms,
               L.ptr(state.chain_partials), L.ptr(dw), L.ptr(dexp), L.ptr(dposes), st)
        scale = None if raw else grads[0][0]
        w_shape, e_shape, p_shape, p_dtype, p_dev = ctx.shapes
        gw = ge = None
        if model_kind != L.MODEL_NONE and ctx.needs_input_grad[0]:
            gw = (dw[:n_terms] if scale is None else dw[:n_terms] * scale).reshape(w_shape)
        if dexp is not None and ctx.needs_input_grad[1]:
            ge = (dexp[:n_terms] if scale is None else dexp[:n_terms] * scale).reshape(e_shape)
        gp = None
        if ctx.needs_input_grad[2]:
            gp = torch.zeros(p_shape, dtype=torch.float64, device=dev)
            gp[:, :3, :] = (dposes if scale is None else dposes * scale).reshape(S, 3, 4)
            gp = gp.to(device=p_dev, dtype=p_dtype)
        return gw, ge, gp, None, None, None, None


def fused_loss(state, model, poses, loss_kind, flags, mask=None):
    """Returns a tensor [2] = (loss_sum, count) (fast path) or the per-point raw loss in ORIGINAL order
    when `flags & FLAG_RAW` (general path: inlier selection, offsets, custom reductions)."""
    state.set_loss_mask(mask)
    kind = model_kind_of(model)
    w = getattr(model, 'w', None) if kind != L.MODEL_NONE else None
    e = getattr(model, 'exponent', None) if kind != L.MODEL_NONE else None
    if isinstance(poses, (list, tuple)):
        poses = torch.stack(list(poses))
    out = _FusedStep.apply(w, e, poses, state, kind, loss_kind, flags)
    if flags & L.FLAG_RAW:
        return out[state.graph.map.inv_order.long()]
    return out
