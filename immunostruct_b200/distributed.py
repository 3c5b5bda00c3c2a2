"""Data parallelism for the hot path: one process per GPU, graph batches sharded across ranks.

The reference is single-device (no torch.distributed call anywhere).  Every graph (or cancer / wild-
type pair) is independent in the forward pass, so the path shards by contiguous ranges of the graph
list: inference scans need NO collective; training needs the gradient averaged across ranks once per step
(6.33 M fp32 = 25.3 MB for HybridModelv2) between the two lines ``loss.backward()`` / ``optimizer.step()`` of
reference procedures/train.py:27-28.

``BucketedGradientReducer`` gives every gradient a slot in one flat buffer (``optim.FlatGradients``: reverse parameter
order, so it fills front to back during the backward pass), cuts the buffer into buckets and, from a
post-accumulate-grad hook, copies a bucket's gradients into their slots with one multi-tensor launch and starts the
bucket's NCCL all-reduce as soon as its last gradient exists: the 24.5 MB of ``vae_fc4`` ... ``vae_fc1`` (the first
gradients autograd produces) are reduced under the whole GNN backward; only the EGNN bucket (0.8 MB) is exposed.
Parameters whose gradient is ``None`` (the last EGNN layer's coord_mlp) are left untouched on every rank, so
optimizers skip them exactly as in the single-GPU run.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist

from .optim import FlatGradients, flatten_gradients


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s weights and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


class BucketedGradientReducer:
    """Average gradients across ranks: bucketed, in place on the flat gradient buffer, overlapped with backward.

    Call ``step()`` between ``loss.backward()`` and ``optimizer.step()``.  The first call discovers which
    parameters receive gradients (identical on every rank: same model, same code path), builds the flat buffer
    and reduces it in one piece.  From the second step on a post-accumulate-grad hook counts the gradients of every
    bucket; when a bucket's last gradient has been written, ONE multi-tensor copy moves the bucket's gradients into
    their slots and the bucket's NCCL all-reduce is launched -- while the rest of the backward pass is still running.
    ``step()`` waits for the buckets and makes every ``p.grad`` its (now averaged) slot, so any optimiser sees the
    averaged gradients and ``FusedAdam`` reads them without another copy.  One ``backward()`` per ``step()`` (as in
    the reference's loops).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 13 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.bucket_bytes = int(bucket_bytes)
        self.fg: FlatGradients = None
        self.buckets: List[Tuple[int, int]] = []     # [lo, hi) element ranges of the flat buffer
        self.overlapped_last_step = 0                 # buckets launched from hooks during the last backward

    # ---- setup --------------------------------------------------------------------------------
    def _setup(self):
        self.fg = flatten_gradients(self.params)
        fg = self.fg
        esz = fg.flat.element_size()
        self.bucket_of, self.buckets, self.members, lo = [], [], [[]], 0
        for i, p in enumerate(fg.params):
            end = fg.offsets[i + 1] if i + 1 < len(fg.params) else fg.flat.numel()
            self.bucket_of.append(len(self.buckets))
            self.members[-1].append(i)
            if (end - lo) * esz >= self.bucket_bytes or i + 1 == len(fg.params):
                self.buckets.append((lo, end))
                lo = end
                if i + 1 < len(fg.params):
                    self.members.append([])
        self.count = [len(m) for m in self.members]
        self.pending = list(self.count)                  # mutated in place (the hooks hold a reference)
        self.works = [None] * len(self.buckets)
        self._world = dist.get_world_size(self.group)
        self._avg = dist.get_backend(self.group) == "nccl"
        for i, p in enumerate(fg.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        b, pending = self.bucket_of[i], self.pending

        def hook(p):
            pending[b] -= 1
            if pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        self.fg.gather(self.members[b])                   # the bucket's gradients -> their slots (one launch)
        lo, hi = self.buckets[b]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self.works[b] = dist.all_reduce(self.fg.flat[lo:hi], op=op, group=self.group, async_op=True)

    # ---- per step -----------------------------------------------------------------------------
    def step(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if self.fg is None:
            self._setup()
            self.pending[:] = [0] * len(self.buckets)    # first step: nothing was launched from hooks
            self.overlapped_last_step = 0
        else:
            self.overlapped_last_step = sum(w is not None for w in self.works)
        for b in range(len(self.buckets)):
            if self.works[b] is None:                      # not launched from a hook (first step / unused parameter)
                self._launch(b)
        for b, w in enumerate(self.works):
            w.wait()
            if not self._avg:
                lo, hi = self.buckets[b]
                self.fg.flat[lo:hi].div_(self._world)
        self.fg.point()                                    # p.grad = averaged slot (no copy)
        self.works = [None] * len(self.buckets)
        self.pending[:] = self.count

    def zero_grad(self) -> None:
        """Drop the gradients (autograd then assigns fresh tensors: no per-parameter accumulation kernels)."""
        for p in self.params:
            p.grad = None

    @property
    def nbytes(self) -> int:
        return 0 if self.fg is None else sum(p.numel() for p in self.fg.params) * self.fg.flat.element_size()


GradientAllReducer = BucketedGradientReducer      # previous name
