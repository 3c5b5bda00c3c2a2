"""Host -> device input pipeline for the hot path (minimal form).

``DevicePrefetcher`` wraps any iterable of host batches in the reference's own format --
``(graph_data | (graph_data, graph_data), sequence_data, target, peptide_property)`` as produced by
``collate`` (data/utils.py:160-176) -- and hands them out already resident on the GPU: the H2D copies of
batch i+1 run on a side stream while batch i is being computed; the (cheap) on-device collation runs on
the consumer's stream right before the batch is yielded.  The ``.to(device)`` calls of the reference
loops (procedures/train.py:20-21, procedures/infer.py:16-17) then return the same objects.  The batch
format does not change.

Memory safety without ``record_stream``: destination buffers come from the consumer stream's allocator
pool, and the copy stream waits for the consumer stream before it writes into them, so a block that is
being re-used is never overwritten while an earlier batch's kernels still read it.
"""
from __future__ import annotations

import torch

from .graph import GraphBatch
from .packed import PackedGraphBatch, PackedSequence


class DevicePrefetcher:
    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _h2d(self, t):
        if not torch.is_tensor(t) or t.device.type != "cpu":
            return t
        src = t if t.is_pinned() else t.pin_memory()
        dst = torch.empty(t.shape, dtype=t.dtype, device=self.device)      # consumer-stream pool
        with torch.cuda.stream(self.copy_stream):
            dst.copy_(src, non_blocking=True)
        return dst

    def _stage(self, obj):
        if isinstance(obj, GraphBatch):
            if obj.device.type != "cpu":
                return obj
            return GraphBatch(self._h2d(obj.ndata["x"]), self._h2d(obj._src_local), self._h2d(obj._dst_local),
                              self._h2d(obj.edata["edge_attr"]), self._h2d(obj._node_counts),
                              self._h2d(obj._edge_counts), obj.max_nodes or int(obj._node_counts.max()),
                              collate_now=False)
        if isinstance(obj, PackedGraphBatch):          # compact arrays over PCIe; expanded in _finish on the consumer's stream
            return obj if obj.device.type != "cpu" else obj._map(self._h2d)
        if isinstance(obj, PackedSequence):
            return PackedSequence(self._h2d(obj.tokens), obj.vocab)
        if torch.is_tensor(obj):
            return self._h2d(obj)
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._stage(o) for o in obj)
        return obj

    def _finish(self, obj):
        """Consumer-stream work on a landed batch; returns the object the caller sees."""
        if isinstance(obj, GraphBatch):
            if obj.indptr is None and obj.ndata["x"].is_cuda:
                obj._collate()                                                # on the consumer's stream
            return obj
        if isinstance(obj, (PackedGraphBatch, PackedSequence)):
            return obj.expand()                                               # csrc/unpack.cu (+ collation)
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._finish(o) for o in obj)
        return obj

    def __iter__(self):
        it = iter(self.loader)
        cur_stream = torch.cuda.current_stream(self.device)

        def stage_next():
            try:
                batch = next(it)
            except StopIteration:
                return None
            self.copy_stream.wait_stream(cur_stream)      # buffers being re-used are no longer read
            return self._stage(batch)

        nxt = stage_next()
        while nxt is not None:
            cur_stream.wait_stream(self.copy_stream)      # batch i has landed
            cur = self._finish(nxt)
            nxt = stage_next()                            # batch i+1: H2D overlaps batch i's compute
            yield cur
