"""On-device training augmentations: the per-sample Python work of the reference's
``SplitDataset.__getitem__`` train branch (data/util_dataloader.py:25-61) applied to a whole device-resident
batch by the kernels of ``csrc/augment.cu``.

reference (per sample, host)                                     here (per batch, device)
  copy.deepcopy(graph)                                             the batch is a fresh device copy already
  x[:, -3:] = RandomRotation()(x[:, -3:])   data/utils.py:148-155  random_rotation_(batch)
  mask_single_structure(graph)              immmunopred_dataloader.py:104-115 (pair: :248-266)
                                                                   mask_single_structure_(batch[, partner])
  mask_structure(graph)                     :92-102 (pair :234-247)  mask_structure_(batch, count)
  mask_sequence(full, peptide, 'J')         :78-90 (pair :216-231)   mask_sequence_(seq, peptide_len, count)

Randomness comes from torch's device generator (the reference uses numpy / python ``random`` on the host, so
the streams differ; the distributions are the same): one 3x3 normal draw per graph, one uniform per graph for the
residue choice, one uniform key per node / position for sampling without replacement.  Every function takes an
optional ``generator`` and returns what the reference returns (the masked residue ids for the SSL head).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _C

N_AA = 20            # one-hot residue columns of x = [one_hot_20 | xyz] (data/preprocess.py:40-41,181)
PAD_TOKEN = 20       # index of PADDING_CHAR 'J' in AMINO_ACIDS + 'J' (data/utils.py:74-76)

__all__ = ["random_rotation_", "mask_single_structure_", "mask_structure_", "mask_sequence_", "TrainAugment"]


def _x(batch):
    x = batch.ndata["x"]
    if not x.is_cuda:
        raise RuntimeError("augmentations run on device-resident batches (call .to(device) first)")
    return x


def random_rotation_(batch, generator: Optional[torch.Generator] = None, return_q: bool = False):
    """Rotate every graph's coordinates (last three columns of x) by its own random orthogonal matrix, in place."""
    x = _x(batch)
    b = batch.n_graphs
    m = torch.randn(b, 9, device=x.device, dtype=torch.float32, generator=generator)
    q = torch.empty(b, 9, device=x.device, dtype=torch.float32) if return_q else None
    _C.rotate_coords(x, x.shape[1] - 3, batch.node_off, m, q)
    return q.view(b, 3, 3) if return_q else None


def mask_single_structure_(batch, partner=None, generator: Optional[torch.Generator] = None):
    """SSL target: overwrite one valid residue per graph with an all-ones one-hot; returns its residue id [B]
    (0 for a graph without a valid residue, as the reference).  ``partner`` (the wild-type batch of a pair) gets a
    residue of the same type masked as well."""
    x = _x(batch)
    b = batch.n_graphs
    dev = x.device
    u = torch.rand(b, device=dev, generator=generator)
    aa = torch.empty(b, dtype=torch.int64, device=dev)
    _C.mask_single_residue(x, N_AA, batch.node_off, u, None, aa)
    if partner is not None:
        u2 = torch.rand(b, device=dev, generator=generator)
        aa2 = torch.empty(b, dtype=torch.int64, device=dev)
        _C.mask_single_residue(_x(partner), N_AA, partner.node_off, u2, aa, aa2)
    return aa


def mask_structure_(batch, count: int, generator: Optional[torch.Generator] = None):
    """Zero the one-hot of ``count`` distinct random nodes per graph (nodes already SSL-masked to ones are skipped)."""
    if count <= 0:
        return
    x = _x(batch)
    keys = torch.rand(x.shape[0], device=x.device, generator=generator)
    _C.mask_rows(x, N_AA, batch.node_off, None, keys, int(count), -1, int(batch.max_nodes))


def mask_sequence_(seq, peptide_len, count: int, pad_token: int = PAD_TOKEN, generator: Optional[torch.Generator] = None,
                   keys=None):
    """Replace ``count`` distinct random positions among the first ``L - peptide_len`` rows of every [L, 21] sequence by
    the padding token's one-hot, in place.  ``peptide_len``: int or int64 tensor [B].  ``keys`` [B * L] lets a
    cancer / wild-type pair share the positions (reference :216-231 masks both at the same indices)."""
    if count <= 0:
        return keys
    if seq.dim() != 3 or not seq.is_contiguous():
        raise ValueError("mask_sequence_: expected a contiguous [B, L, V] tensor")
    b, l, v = seq.shape
    dev = seq.device
    if keys is None:
        keys = torch.rand(b * l, device=dev, generator=generator)
    seg = torch.arange(b + 1, device=dev, dtype=torch.int64) * l
    if torch.is_tensor(peptide_len):
        limit = (l - peptide_len.to(dev, torch.int64)).contiguous()
    else:
        limit = torch.full((b,), l - int(peptide_len), dtype=torch.int64, device=dev)
    _C.mask_rows(seq.view(b * l, v), v, seg, limit, keys, int(count), int(pad_token), l)
    return keys


class TrainAugment:
    """The train-split branch of ``SplitDataset.__getitem__`` for a whole batch, in the reference's order: rotation,
    then (SSL) single-residue masking, then structure masking; sequence masking on the dense sequence tensor.
    ``__call__(graph_data, sequence_data)`` -> (graph_data, sequence_data, amino_acid or None); pairs are tuples."""

    def __init__(self, structure_pad_count: int = 0, sequence_pad_count: int = 0, return_amino_acid: bool = False,
                 peptide_len: int = 11, generator: Optional[torch.Generator] = None):
        self.structure_pad_count, self.sequence_pad_count = structure_pad_count, sequence_pad_count
        self.return_amino_acid, self.peptide_len, self.generator = return_amino_acid, peptide_len, generator

    def __call__(self, graph_data, sequence_data=None):
        g = self.generator
        pair = isinstance(graph_data, (tuple, list))
        graphs = list(graph_data) if pair else [graph_data]
        for gb in graphs:
            random_rotation_(gb, g)
        aa = None
        if self.return_amino_acid:
            aa = mask_single_structure_(graphs[0], graphs[1] if pair else None, g)
        for gb in graphs:
            mask_structure_(gb, self.structure_pad_count, g)
        if sequence_data is not None and self.sequence_pad_count > 0:
            seqs = list(sequence_data) if pair else [sequence_data]
            keys = None
            for s in seqs:
                keys = mask_sequence_(s, self.peptide_len, self.sequence_pad_count, generator=g, keys=keys)
        return graph_data, sequence_data, aa
