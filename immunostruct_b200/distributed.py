"""Data parallelism for the hot path: one process per GPU, graph batches sharded across ranks.

The reference is single-device (no torch.distributed call anywhere).  Every graph (or cancer / wild-
type pair) is independent in the forward pass, so the path shards by contiguous ranges of the graph
list: inference scans need NO collective; training needs one gradient all-reduce per step
(6.33 M fp32 = 25.3 MB for HybridModelv2), done here over NCCL (NVLink 5 / NVSwitch) on a flat
buffer.  Parameters whose gradient is ``None`` (the last EGNN layer's coord_mlp) are left untouched
on every rank, so optimizers skip them exactly as in the single-GPU run.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


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


class GradientAllReducer:
    """Average gradients across ranks with one all-reduce over a persistent flat buffer.

    ``step()`` is called between ``loss.backward()`` and ``optimizer.step()`` (the two lines at
    reference procedures/train.py:27-28).  The set of parameters that receive gradients is discovered
    on the first call and must be the same on every rank (it is: same model, same code path).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.flat = None
        self.live = None

    def _setup(self):
        self.live = [p for p in self.params if p.grad is not None]
        n = sum(p.grad.numel() for p in self.live)
        g0 = self.live[0].grad
        self.flat = torch.empty(n, dtype=g0.dtype, device=g0.device)
        self.views, o = [], 0
        for p in self.live:
            k = p.grad.numel()
            self.views.append(self.flat[o:o + k].view_as(p.grad))
            o += k

    def step(self) -> None:
        if not dist.is_initialized():
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        if self.flat is None:
            self._setup()
        grads = [p.grad for p in self.live]
        if any(g is None for g in grads):
            raise RuntimeError("the set of parameters with gradients changed between steps")
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        torch._foreach_copy_(grads, self.views)

    @property
    def nbytes(self) -> int:
        return 0 if self.flat is None else self.flat.numel() * self.flat.element_size()
