"""Seeded synthetic pMHC inputs of the shape the reference's datasets produce.

There is no network and the reference's AlphaFold graph ``.pt`` files are absent, so every test
and benchmark uses this generator: per graph ``n_nodes`` residues with
``x = [one_hot_20(residue type) | xyz]`` (layout of ``data/preprocess.py:40-41,181``), directed
k-nearest-neighbour contact edges ``neighbour -> node`` (in-degree exactly ``k``), ``edge_attr = 1``
(``data/utils.py:60``), an optional tail of all-zero isolated padding nodes (``pad_graph``,
``data/utils.py:13-33``), a 283x21 one-hot sequence, two peptide properties and a target.
Everything is produced with torch ops on the requested device from an explicit generator.
"""
from __future__ import annotations

import torch

SEQ_LEN, SEQ_ALPHABET = 283, 21


def synthetic_graph_arrays(n_graphs: int, n_nodes: int = 200, k: int = 10, seed: int = 1,
                           device="cpu", n_pad: int = 0, shuffle_edges: bool = True,
                           coord_scale: float = 10.0) -> dict:
    """Concatenated per-graph arrays with GRAPH-LOCAL edge endpoints (what ``batch`` consumes).

    Returns ``x`` [B*N,23] f32, ``src``/``dst`` int64 [B*E] (local ids), ``edge_attr`` [B*E,1] f32,
    ``node_counts``/``edge_counts`` int64 [B].  ``n_pad`` trailing nodes per graph are all-zero and
    isolated; the k-NN graph is built over the first ``n_nodes - n_pad`` residues only.
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    n_real = n_nodes - n_pad
    assert n_real > k, "need more real residues than neighbours"
    xyz = torch.randn(n_graphs, n_real, 3, generator=gen, device=dev) * coord_scale
    aa = torch.randint(0, 20, (n_graphs, n_real), generator=gen, device=dev)
    x = torch.zeros(n_graphs, n_nodes, 23, device=dev)
    x[:, :n_real, :20] = torch.nn.functional.one_hot(aa, 20).float()
    x[:, :n_real, 20:] = xyz
    d = torch.cdist(xyz, xyz)
    d.diagonal(dim1=1, dim2=2).fill_(float("inf"))
    nbr = d.topk(k, dim=2, largest=False).indices                      # [B, n_real, k]
    dst = torch.arange(n_real, device=dev).view(1, n_real, 1).expand(n_graphs, n_real, k)
    src = nbr.reshape(n_graphs, -1)
    dst = dst.reshape(n_graphs, -1)
    e = src.shape[1]
    if shuffle_edges:
        perm = torch.rand(n_graphs, e, generator=gen, device=dev).argsort(dim=1)
        src, dst = src.gather(1, perm), dst.gather(1, perm)
    return {
        "x": x.reshape(-1, 23).contiguous(),
        "src": src.reshape(-1).contiguous().to(torch.int64),
        "dst": dst.reshape(-1).contiguous().to(torch.int64),
        "edge_attr": torch.ones(n_graphs * e, 1, device=dev),
        "node_counts": torch.full((n_graphs,), n_nodes, dtype=torch.int64, device=dev),
        "edge_counts": torch.full((n_graphs,), e, dtype=torch.int64, device=dev),
    }


def synthetic_dense(n_graphs: int, seed: int = 1, device="cpu", positive_rate: float = 0.19) -> dict:
    """Sequence one-hots [B,283,21], peptide properties [B,2], BCE targets and regression targets."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed + 7919)
    tok = torch.randint(0, SEQ_ALPHABET, (n_graphs, SEQ_LEN), generator=gen, device=dev)
    seq = torch.nn.functional.one_hot(tok, SEQ_ALPHABET).float()
    prop = torch.rand(n_graphs, 2, generator=gen, device=dev)
    y_bin = (torch.rand(n_graphs, generator=gen, device=dev) < positive_rate).float()
    y_reg = torch.rand(n_graphs, generator=gen, device=dev) * 2 - 1
    return {"seq": seq, "prop": prop, "target": y_bin, "target_reg": y_reg}


def split_graphs(arrays: dict) -> list:
    """Per-sample graph dicts (the items a Dataset would hand to ``collate``)."""
    out, n0, e0 = [], 0, 0
    for n, e in zip(arrays["node_counts"].tolist(), arrays["edge_counts"].tolist()):
        out.append({"src": arrays["src"][e0:e0 + e], "dst": arrays["dst"][e0:e0 + e], "num_nodes": n,
                    "x": arrays["x"][n0:n0 + n], "edge_attr": arrays["edge_attr"][e0:e0 + e]})
        n0, e0 = n0 + n, e0 + e
    return out
