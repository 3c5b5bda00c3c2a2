"""Step-level fusion: a whole step -- on-device collation, forward, loss, backward, fused optimiser --
captured ONCE into a CUDA graph and replayed.  A training step of the reference's loop (procedures/train.py:16-30) is ~300
kernel launches in ~9 ms; the idle gaps between them are 6 % of the step, which the replay removes.

    static = {"x": ..., "src": ..., ..., "seq": ..., "target": ...}        # device tensors of the step's fixed shapes
    def body(t):                                                             # everything the step does, reading only `t`
        gb = GraphBatch.from_arrays(t["x"], t["src"], ..., max_nodes=200)
        opt.zero_grad(set_to_none=True)
        recon, mu, logvar, out = model(gb, t["seq"], t["prop"])
        loss = losses.BCE_loss(recon, t["seq"], mu, logvar, out, t["target"])
        loss.backward(); opt.step()                                          # FusedAdam(capturable=True): step count on the device
        return loss
    step = CapturedStep(body, static)
    for batch in loader: loss = step(batch)                                  # copies into the static buffers, replays

Requirements (the usual ones of CUDA graphs): fixed shapes, no host synchronisation inside ``body`` (the kernels of this
package have none; ``GraphBatch`` validation is deferred), optimiser state on the device.  Random draws (dropout, the VAE's
``randn``) advance through torch's graph-safe generator offsets on every replay."""
from __future__ import annotations

from typing import Callable, Dict

import torch

__all__ = ["CapturedStep"]


class CapturedStep:
    def __init__(self, body: Callable[[Dict[str, torch.Tensor]], torch.Tensor], static_inputs: Dict[str, torch.Tensor], warmup: int = 3):
        if not static_inputs:
            raise ValueError("static_inputs must hold the step's device tensors")
        self.static = {k: v.clone() for k, v in static_inputs.items()}
        dev = next(iter(self.static.values())).device
        if dev.type != "cuda":
            raise RuntimeError("CapturedStep needs CUDA tensors (CUDA graphs)")
        self._body = body
        self._out = None
        # warm-up on a side stream (lazy initialisations, allocator pools, weight-plane caches), then capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                body(self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = body(self.static)
            self._out = out.detach() if torch.is_tensor(out) else out
        self.warmup_steps = max(1, warmup)              # REAL steps the body has already performed on the initial inputs (capture records, it does not run)

    def __call__(self, inputs: Dict[str, torch.Tensor] | None = None):
        """Copy ``inputs`` (same keys / shapes as the static buffers; host tensors are copied asynchronously if pinned) into
        the static buffers and replay the step.  Returns the captured output tensor (overwritten by the next replay)."""
        if inputs is not None:
            for k, v in inputs.items():
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self._out
