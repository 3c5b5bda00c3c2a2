"""Compact host-side batch format + on-device expansion.

The reference moves fp32 one-hots over PCIe: ``x = [one_hot_20 | xyz]`` per residue (data/utils.py:75-89,
preprocess.py:40-41,181), int64 edge endpoints, an all-ones ``edge_attr`` (data/utils.py:60) and the
peptide + MHC pseudo-sequence as a ``[283, 21]`` fp32 one-hot -- 82 KB per 200-residue graph.  ``PackedGraphBatch``
and ``PackedSequence`` carry one byte per residue, fp32 coordinates, int32 graph-local endpoints and one byte per
sequence position (19 KB per graph) and expand on the device -- bit-exactly -- into the ``GraphBatch`` / dense
sequence tensor that the unchanged model API takes:

    gb_packed, seq_packed = pack_graph_batch(graph_batch), pack_sequence(sequence_data)     # data pipeline, once
    graph_data = gb_packed.to(device)            # H2D of the compact arrays + csrc/unpack.cu + on-device collation
    sequence_data = seq_packed.to(device)        # [B, 283, 21] fp32 on the device
    model(graph_data, sequence_data, peptide_property)

Feature rows that are not one-hot (the SSL "mask to one" variant, data/immmunopred_dataloader.py:111) cannot be
packed; ``pack_graph_batch`` refuses them and the dense ``GraphBatch`` path stays available.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _C
from .graph import GraphBatch

PAD_RESIDUE = 255          # all-zero feature row (the reference's zero-padded nodes, data/utils.py:13-33)


class PackedGraphBatch:
    """aa u8 [N] | xyz f32 [N,3] | src, dst i32 [E] (graph-local) | edge_attr f32 [E] or None (= ones) | counts i32 [B]."""

    def __init__(self, aa, xyz, src, dst, edge_attr, node_counts, edge_counts, max_nodes: Optional[int] = None):
        self.aa, self.xyz, self.src, self.dst, self.edge_attr = aa, xyz, src, dst, edge_attr
        self.node_counts, self.edge_counts, self.max_nodes = node_counts, edge_counts, max_nodes

    def _tensors(self):
        return [t for t in (self.aa, self.xyz, self.src, self.dst, self.edge_attr, self.node_counts, self.edge_counts)
                if t is not None]

    @property
    def device(self):
        return self.aa.device

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._tensors())

    def _map(self, fn):
        return PackedGraphBatch(fn(self.aa), fn(self.xyz), fn(self.src), fn(self.dst),
                                None if self.edge_attr is None else fn(self.edge_attr), fn(self.node_counts),
                                fn(self.edge_counts), self.max_nodes)

    def pin_memory(self):
        return self if self.device.type != "cpu" else self._map(lambda t: t.pin_memory())

    def expand(self) -> GraphBatch:
        """Device-resident packed arrays -> GraphBatch (dense features, int64 endpoints, CSR / CSC / offsets)."""
        if not self.aa.is_cuda:
            raise RuntimeError("PackedGraphBatch.expand() runs on the GPU; call .to(device) (there is no CPU path)")
        dev, n, e = self.aa.device, self.aa.numel(), self.src.numel()
        x = torch.empty(n, 23, dtype=torch.float32, device=dev)
        src64, dst64 = torch.empty(e, dtype=torch.int64, device=dev), torch.empty(e, dtype=torch.int64, device=dev)
        attr = torch.empty(e, 1, dtype=torch.float32, device=dev)
        _C.unpack_nodes(self.aa, self.xyz, x)
        _C.unpack_edges(self.src, self.dst, self.edge_attr, src64, dst64, attr)
        return GraphBatch(x, src64, dst64, attr, self.node_counts.to(torch.int64), self.edge_counts.to(torch.int64),
                          self.max_nodes)

    def to(self, device, non_blocking: bool = False) -> GraphBatch:
        """H2D of the compact arrays, expansion and collation on the device; returns the GraphBatch the models take."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("PackedGraphBatch expands on a CUDA device only")
        return self._map(lambda t: t.to(device, non_blocking=non_blocking)).expand()


class PackedSequence:
    """tokens u8 [B, L] (0 .. vocab-1; >= vocab = all-zero position) standing for a [B, L, vocab] fp32 one-hot."""

    def __init__(self, tokens, vocab: int = 21):
        self.tokens, self.vocab = tokens, vocab

    @property
    def nbytes(self) -> int:
        return self.tokens.numel()

    def pin_memory(self):
        return self if self.tokens.device.type != "cpu" else PackedSequence(self.tokens.pin_memory(), self.vocab)

    def expand(self):
        if not self.tokens.is_cuda:
            raise RuntimeError("PackedSequence.expand() runs on the GPU")
        out = torch.empty(*self.tokens.shape, self.vocab, dtype=torch.float32, device=self.tokens.device)
        _C.onehot_tokens(self.tokens.contiguous(), out, self.vocab)
        return out

    def to(self, device, non_blocking: bool = False):
        return PackedSequence(self.tokens.to(device, non_blocking=non_blocking), self.vocab).expand()


def pack_graph_batch(gb: GraphBatch) -> PackedGraphBatch:
    """Host GraphBatch (dense, as produced by ``batch`` / ``collate``) -> compact form.  Raises if a feature row is
    neither one-hot nor all-zero, or if endpoints / counts do not fit int32."""
    x = gb.ndata["x"]
    if x.device.type != "cpu":
        raise ValueError("pack_graph_batch packs host batches (the data pipeline side)")
    feat = x[:, :20]
    ones = (feat == 1.0).sum(1)
    if not bool((((ones == 1) | (ones == 0)) & ((feat != 0.0).sum(1) == ones)).all()):
        raise ValueError("feature rows must be one-hot or all-zero to be packed")
    aa = torch.where(ones == 1, feat.argmax(1), torch.full_like(ones, PAD_RESIDUE)).to(torch.uint8)
    ea = gb.edata["edge_attr"].reshape(-1)
    return PackedGraphBatch(aa, x[:, 20:].contiguous(), gb._src_local.to(torch.int32), gb._dst_local.to(torch.int32),
                            None if bool((ea == 1.0).all()) else ea.contiguous().float(),
                            gb._node_counts.to(torch.int32), gb._edge_counts.to(torch.int32),
                            gb.max_nodes or (int(gb._node_counts.max()) if gb.n_graphs else 0))


def pack_sequence(sequence_data: torch.Tensor) -> PackedSequence:
    """[B, L, V] one-hot floats -> PackedSequence (positions that are all-zero become token V)."""
    if sequence_data.dim() != 3:
        raise ValueError("expected [B, L, V] one-hot sequence data")
    v = sequence_data.shape[2]
    ones = (sequence_data == 1.0).sum(2)
    if not bool((((ones == 1) | (ones == 0)) & ((sequence_data != 0.0).sum(2) == ones)).all()):
        raise ValueError("sequence rows must be one-hot or all-zero to be packed")
    tok = torch.where(ones == 1, sequence_data.argmax(2), torch.full_like(ones, v))
    return PackedSequence(tok.to(torch.uint8), v)


class PackedGraphDataset:
    """A whole dataset in a handful of flat compact arrays.

    The reference keeps a python list of DGL graphs, ``copy.deepcopy``s each sample, concatenates them one by one in
    ``dgl.batch`` inside DataLoader workers and pickles the batch to the main process (train_IEDB_wFT.py:82-87,
    data/util_dataloader.py:25-50, data/utils.py:148-163).  Here the samples are packed once
    (``from_samples``); a batch is then a VECTORISED gather over the flat arrays -- no per-graph python work, no
    pickling, optionally pinned so that the H2D copy of the compact batch is asynchronous -- and ``loader`` yields
    ``(PackedGraphBatch, PackedSequence, target, property)`` tuples that ``DevicePrefetcher`` turns into the dense
    device batch of the unchanged model API.
    """

    def __init__(self, aa, xyz, src, dst, edge_attr, node_counts, edge_counts, tokens, targets, properties, vocab=21):
        self.aa, self.xyz, self.src, self.dst, self.edge_attr = aa, xyz, src, dst, edge_attr
        self.node_counts, self.edge_counts = node_counts.to(torch.int64), edge_counts.to(torch.int64)
        self.tokens, self.targets, self.properties, self.vocab = tokens, targets, properties, vocab
        z = torch.zeros(1, dtype=torch.int64)
        self.node_off = torch.cat([z, torch.cumsum(self.node_counts, 0)])
        self.edge_off = torch.cat([z, torch.cumsum(self.edge_counts, 0)])

    @classmethod
    def from_samples(cls, graphs, sequences, targets, properties):
        """graphs: sequence of ``Graph`` (one-hot / zero-padded features); sequences [S, L, V] one-hot floats;
        targets [S]; properties [S, 2]."""
        from .graph import batch as _batch
        pk = pack_graph_batch(_batch(list(graphs)))
        ps = pack_sequence(torch.as_tensor(sequences))
        return cls(pk.aa, pk.xyz, pk.src, pk.dst, pk.edge_attr, pk.node_counts, pk.edge_counts, ps.tokens,
                   torch.as_tensor(targets), torch.as_tensor(properties), ps.vocab)

    def __len__(self):
        return int(self.node_counts.numel())

    @staticmethod
    def _ranges(starts, counts):
        """Concatenated index ranges [starts_i, starts_i + counts_i) without a python loop."""
        total = int(counts.sum())
        if total == 0:
            return torch.zeros(0, dtype=torch.int64)
        out_off = torch.cumsum(counts, 0) - counts
        return torch.arange(total) + torch.repeat_interleave(starts - out_off, counts)

    def batch(self, indices, pin_memory: bool = False):
        """-> (PackedGraphBatch, PackedSequence, target [B], property [B,2]) for the given sample indices."""
        idx = torch.as_tensor(indices, dtype=torch.int64).reshape(-1)
        nc, ec = self.node_counts[idx], self.edge_counts[idx]
        ni, ei = self._ranges(self.node_off[idx], nc), self._ranges(self.edge_off[idx], ec)
        pk = PackedGraphBatch(self.aa[ni], self.xyz[ni], self.src[ei], self.dst[ei],
                              None if self.edge_attr is None else self.edge_attr[ei], nc.to(torch.int32), ec.to(torch.int32),
                              int(nc.max()) if idx.numel() else 0)
        ps = PackedSequence(self.tokens[idx], self.vocab)
        tgt, prop = self.targets[idx], self.properties[idx]
        if pin_memory:
            pk, ps, tgt, prop = pk.pin_memory(), ps.pin_memory(), tgt.pin_memory(), prop.pin_memory()
        return pk, ps, tgt, prop

    def loader(self, batch_size: int, shuffle: bool = False, drop_last: bool = False, generator=None, pin_memory: bool = False):
        """Batches in the reference loader's order semantics (last partial batch kept unless ``drop_last``)."""
        n = len(self)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        for lo in range(0, n, batch_size):
            sel = order[lo:lo + batch_size]
            if drop_last and sel.numel() < batch_size:
                return
            yield self.batch(sel, pin_memory=pin_memory)
