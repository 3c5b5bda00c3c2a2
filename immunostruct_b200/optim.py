"""Flat gradient storage and the fused Adam / AdamW step.

The reference steps ``torch.optim.Adam`` / ``AdamW`` over ~60 separate parameter tensors
(train_IEDB_wFT.py:74,97; train_Cancer_wFT.py:98,122; procedures/train.py:28,122).  Here the parameters that
receive gradients live in ONE flat fp32 buffer (each ``nn.Parameter`` is a view of it, so ``state_dict``,
``torch.save`` and ``load_state_dict`` are unchanged), their gradients are gathered into a second flat buffer, and one
launch of ``is_fused_adam`` (csrc/optim.cu) updates everything.  The flat gradient buffer is shared with
``distributed.BucketedGradientReducer``, which fills and all-reduces it bucket by bucket while the backward pass is
still running.

Parameters whose gradient is ``None`` after the first backward pass (the last EGNN layer's ``coord_mlp``,
hybrid_models.py:323-326) stay outside the flat buffers and are never touched -- exactly what torch's optimisers do
with ``grad is None`` (no moment update, no weight decay).
"""
from __future__ import annotations

import math
from typing import Iterable, List

import torch

from . import _C

__all__ = ["FlatGradients", "flatten_gradients", "FusedAdam", "FusedAdamW"]


class FlatGradients:
    """One flat buffer with a slot (view) per parameter that has a gradient, in REVERSE parameter order: autograd
    finishes the last layers first, so the buffer fills front to back during the backward pass.

    Gradients are not aliased by default: autograd keeps handing out fresh per-parameter tensors (free with
    ``zero_grad(set_to_none=True)``), and ``gather`` copies a whole group of them into their slots with ONE multi-tensor
    launch -- aliasing instead makes autograd run ``slot += grad`` once per parameter (~100 tiny kernels per step).
    ``point`` makes ``p.grad`` the slots (after an all-reduce: every consumer of ``p.grad`` then sees the averaged
    values); a gradient that already is its slot is skipped by ``gather``."""

    def __init__(self, params: List[torch.nn.Parameter], alias: bool = False):
        self.params = [p for p in reversed(params) if p.grad is not None]
        if not self.params:
            raise RuntimeError("no parameter has a gradient yet: call after the first backward()")
        g0 = self.params[0].grad
        # every slot starts on a 16-byte boundary (vector accesses of the fused optimiser, NCCL alignment)
        self.offsets, o = [], 0
        for p in self.params:
            self.offsets.append(o)
            o += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(o, dtype=g0.dtype, device=g0.device)
        self.views = [self.flat[a:a + p.numel()].view_as(p) for a, p in zip(self.offsets, self.params)]
        for p in self.params:
            p._is_flat_grad = self
        if alias:
            self.realias()

    def gather(self, indices=None) -> None:
        """Copy the current ``.grad`` of the given parameters (default: all) into their slots: one multi-tensor launch."""
        src, dst = [], []
        for i in (range(len(self.params)) if indices is None else indices):
            g, v = self.params[i].grad, self.views[i]
            if g is None:
                v.zero_()
            elif g is not v and g.data_ptr() != v.data_ptr():
                src.append(g)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)

    def point(self, indices=None) -> None:
        """``p.grad = slot`` (no copy)."""
        for i in (range(len(self.params)) if indices is None else indices):
            self.params[i].grad = self.views[i]

    def realias(self) -> None:
        """gather + point: every ``p.grad`` becomes its slot, holding the same values."""
        self.gather()
        self.point()

    def zero_(self) -> None:
        self.flat.zero_()
        self.point()

    def index(self, p) -> int:
        for i, q in enumerate(self.params):
            if q is p:
                return i
        raise KeyError("parameter not in this FlatGradients")


def flatten_gradients(params: Iterable[torch.nn.Parameter], alias: bool = False) -> FlatGradients:
    """The ``FlatGradients`` of ``params`` -- created on first use, shared by every later caller with the same
    live parameter set (the gradient reducer and the optimiser of one model)."""
    params = [p for p in params if p.requires_grad]
    live = [p for p in params if p.grad is not None]
    fg = getattr(live[0], "_is_flat_grad", None) if live else None
    if fg is not None and len(fg.params) == len(live) and all(a is b for a, b in zip(fg.params, reversed(live))):
        if alias:
            fg.realias()
        return fg
    return FlatGradients(params, alias)


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (``decoupled=False``) / ``AdamW`` (``decoupled=True``) with the whole step in one kernel.

    Same defaults, hyper-parameter names (``param_groups[i]['lr']`` is read every step, so LR schedulers work) and
    per-parameter state keys (``step``, ``exp_avg``, ``exp_avg_sq``) as torch's optimisers; ``amsgrad`` /
    ``maximize`` are not supported.  The per-parameter gradients are gathered into the flat buffer by one multi-tensor
    copy per step -- or are already there when ``distributed.BucketedGradientReducer`` has all-reduced them."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, grad_scale=1.0,
                 capturable=False):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled))
        self.grad_scale = float(grad_scale)
        self.capturable = bool(capturable)      # step count on the device: the step can be replayed from a CUDA graph
        self._flat = {}       # group index -> dict(fg, p, m, v, runs, step)

    def _setup_group(self, gi, group):
        fg = flatten_gradients(group["params"])
        mine = {id(p) for p in group["params"]}
        idx = [i for i, p in enumerate(fg.params) if id(p) in mine]
        n = sum((fg.params[i].numel() + 3) // 4 * 4 for i in idx)
        dev = fg.flat.device
        fp = torch.zeros(n, dtype=torch.float32, device=dev)
        fm, fv = torch.zeros_like(fp), torch.zeros_like(fp)
        runs, o = [], 0           # (param offset, grad offset, length) merged while contiguous in both buffers
        for i in idx:
            p = fg.params[i]
            k, kp = p.numel(), (p.numel() + 3) // 4 * 4
            with torch.no_grad():
                fp[o:o + k].copy_(p.detach().reshape(-1))
            p.data = fp[o:o + k].view_as(p)
            st = self.state[p]
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = fm[o:o + k].view_as(p)
            st["exp_avg_sq"] = fv[o:o + k].view_as(p)
            go = fg.offsets[i]
            if runs and runs[-1][0] + runs[-1][2] == o and runs[-1][1] + runs[-1][2] == go:
                runs[-1] = (runs[-1][0], runs[-1][1], runs[-1][2] + kp)
            else:
                runs.append((o, go, kp))
            o += kp
        self._flat[gi] = dict(fg=fg, p=fp, m=fm, v=fv, runs=runs, step=0, idx=idx,
                              step_dev=torch.zeros(1, dtype=torch.float32, device=dev) if self.capturable else None)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if gi not in self._flat:
                if all(p.grad is None for p in group["params"]):
                    continue
                self._setup_group(gi, group)
            fl = self._flat[gi]
            fg = fl["fg"]
            fg.gather(fl["idx"])           # one multi-tensor copy; skipped for gradients that already are their slots
                                           # (a gradient reducer has all-reduced them in the flat buffer)
            fl["step"] += 1
            t = fl["step"]
            b1, b2 = group["betas"]
            lr = float(group["lr"])
            live = [fg.params[i] for i in fl["idx"]]
            if self.capturable:
                for k, (po, go, n) in enumerate(fl["runs"]):
                    _C.fused_adam_capturable(fl["p"][po:po + n], fg.flat[go:go + n], fl["m"][po:po + n], fl["v"][po:po + n],
                                             lr, b1, b2, group["eps"], group["weight_decay"], group["decoupled"],
                                             fl["step_dev"], k == 0, self.grad_scale)
                continue                        # (no host-side bookkeeping inside a captured step)
            step_size = lr / (1.0 - b1 ** t)
            inv_bc2_sqrt = 1.0 / math.sqrt(1.0 - b2 ** t)
            for po, go, n in fl["runs"]:
                _C.fused_adam(fl["p"][po:po + n], fg.flat[go:go + n], fl["m"][po:po + n], fl["v"][po:po + n], lr, b1, b2,
                              group["eps"], group["weight_decay"], group["decoupled"], step_size, inv_bc2_sqrt,
                              self.grad_scale)
            torch.autograd.graph.increment_version(live)          # the kernel wrote through raw pointers: caches keyed on
            for p in live:                                        # the version counters must see the update
                self.state[p]["step"] = torch.tensor(float(t))
        return loss

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)


class FusedAdamW(FusedAdam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0, capturable=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=True, grad_scale=grad_scale,
                         capturable=capturable)
