"""CUDA-graph capture of the fixed-graph optimisation iteration.

The reference's own runs (BASELINE.json configs[0] / [3]: 0.08-0.2 M points) sit far below the size at which the
kernels of the step are the bound: one iteration of scripts/model_poses_learning:121-135 is ~25 kernel launches of
2-20 us each, issued by ~1 ms of Python (autograd, optimizer, wrappers).  With the neighbourhood graph fixed, every
launch of the iteration has the same arguments each time -- pose corrections and model weights are read from device
memory -- so the whole iteration (corrected poses -> model -> transform -> cov/eig -> loss -> backward -> optimizer)
is recorded once into a CUDA graph and replayed with a single launch.

    opt = torch.optim.Adam(params, lr=1e-3, capturable=True)      # the optimizer state must live on the device

    def iteration():
        poses_upd = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
        cloud = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=poses_upd),
                                                 neighborhoods=ns, cfg=cfg)
        loss, _ = dc.min_eigval_loss(cloud, mask=mask)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    step = dc.CapturedIteration(iteration)        # runs `warmup` eager iterations, then records one
    for it in range(n_opt_iters):
        loss = step()                             # one graph launch; `loss` is the same tensor every time

What may not happen inside `iteration`: anything that reads device data on the host (`.item()`, printing a tensor,
data-dependent Python branches), a neighbour search (the graph is what makes the launches static), or a change of
the set of tensors involved (new scans, a new mask object).  Those belong between replays.
"""
import torch

__all__ = ['CapturedIteration']


class CapturedIteration(object):
    def __init__(self, iteration, warmup=3):
        """iteration: callable running ONE optimisation iteration on the current stream and returning a tensor or a
        tuple of tensors (typically the loss).  `warmup` eager iterations run first (they also advance the
        optimisation): allocations, the step state of the global cloud, the transposed graph of a kNN neighbourhood
        (built on the third backward pass, fused.py) and the optimizer state all have to exist before recording."""
        assert torch.cuda.is_available(), 'CapturedIteration needs a CUDA device'
        assert warmup >= 3, 'at least 3 eager iterations are needed before the launches of the iteration are static'
        self._iteration = iteration
        self.warmup_outputs = []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = iteration()
                self.warmup_outputs.append(_detached(out))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib as L
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count
        with torch.cuda.graph(self.graph):
            self.outputs = iteration()
        self.library_launches = L.launch_count - n0         # kernels of libdcb200 inside one replay
        self.replays = 0

    def __call__(self):
        """Replay the recorded iteration on the current stream.  Returns the output tensor(s) of the recording: the
        same objects every call, overwritten by each replay (clone to keep a value across iterations)."""
        self.graph.replay()
        self.replays += 1
        return self.outputs


def _detached(out):
    if isinstance(out, torch.Tensor):
        return out.detach().clone()
    if isinstance(out, (list, tuple)):
        return type(out)(_detached(o) for o in out)
    return out
