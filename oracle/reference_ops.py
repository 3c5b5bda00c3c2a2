"""CPU oracle for the ImmunoStruct hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain-PyTorch (CPU, fp32 or fp64) restatement of the arithmetic of the
reference's forward/backward path.  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it, and only as the checker / the timed CPU baseline.  Nothing under
``immunostruct_b200/`` imports it.

Parity status: the restatements of the reference's *own* code (attention blocks, model
forward passes, losses, contrastive loss) are pinned against the unmodified reference files
imported from ``/root/reference`` (see ``oracle/shim.py`` + ``tests/golden/make_golden.py``;
the committed vectors live in ``tests/golden/*.npz``).  The arithmetic that lives in
un-vendored third-party packages -- ``dgl.nn.EGNNConv``, ``dgl.batch`` (DGL, version
unpinned by the reference, README.md:107,137) and ``torch_geometric.nn.global_mean_pool`` /
``global_max_pool`` (torch_geometric 2.5.3, README.md:108,143) -- is restated here from the
published algorithm.  Neither package is installable in this
environment and the reference ships no tests or golden vectors, so for those operators
**parity is unpinned** by upstream outputs; they are pinned against hand-derived
known-answer cases instead (tests/test_oracle.py).

All functions are written as pure functions over a ``state_dict``-style mapping of
parameter name -> tensor so that the very same weights can be fed to the oracle and to
the CUDA product path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# Batching / CSR construction  (reference: immunostruct/data/utils.py:160-176 -> dgl.batch)
# --------------------------------------------------------------------------------------
def dgl_batch(graphs: Sequence[dict]) -> dict:
    """Restatement of ``dgl.batch`` as used by ``collate`` (data/utils.py:163,169-170).

    Each graph is a dict with ``src``/``dst`` int64 [E_g] (graph-local ids), ``num_nodes``,
    ``x`` [N_g, 23] and ``edge_attr`` [E_g, 1] (``to_dgl``, data/utils.py:54-67).
    Nodes are concatenated graph-major; edge endpoints are shifted by the running node
    count; edge order is preserved; no sorting, no de-duplication.
    """
    node_counts = torch.tensor([int(g["num_nodes"]) for g in graphs], dtype=torch.int64)
    edge_counts = torch.tensor([int(g["src"].numel()) for g in graphs], dtype=torch.int64)
    node_off = torch.zeros(len(graphs) + 1, dtype=torch.int64)
    node_off[1:] = torch.cumsum(node_counts, 0)
    src = torch.cat([g["src"].to(torch.int64) + node_off[i] for i, g in enumerate(graphs)])
    dst = torch.cat([g["dst"].to(torch.int64) + node_off[i] for i, g in enumerate(graphs)])
    return {
        "src": src,
        "dst": dst,
        "x": torch.cat([g["x"] for g in graphs], 0),
        "edge_attr": torch.cat([g["edge_attr"] for g in graphs], 0),
        "batch_num_nodes": node_counts,
        "batch_num_edges": edge_counts,
        "num_nodes": int(node_off[-1]),
    }


def batch_vector(batch_num_nodes: Tensor) -> Tensor:
    """Graph id of every node; the flattened ``batch_tensor`` of hybrid_models.py:86-87,97."""
    return torch.repeat_interleave(torch.arange(batch_num_nodes.numel(), dtype=torch.int64),
                                   batch_num_nodes.to(torch.int64))


def csr_from_coo(src: Tensor, dst: Tensor, num_nodes: int) -> dict:
    """Canonical destination-sorted CSR + its CSC transpose.

    ``perm = argsort(dst, stable)``; ``indptr = [0, cumsum(bincount(dst))]``;
    ``csr_src = src[perm]``; ``csr_dst = dst[perm]``; ``csr_eid = perm`` (DGL's lazily built
    CSC uses a stable sort by destination, so in-edges of a node appear in ascending edge id).
    The transpose groups the *CSR positions* by source node (stable in edge id):
    ``outptr = [0, cumsum(bincount(src))]``, ``csc_pos = inv_perm[argsort(src, stable)]``.
    """
    src = src.to(torch.int64)
    dst = dst.to(torch.int64)
    perm = torch.argsort(dst, stable=True)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(torch.bincount(dst, minlength=num_nodes), 0)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), dtype=torch.int64)
    perm_s = torch.argsort(src, stable=True)
    outptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    outptr[1:] = torch.cumsum(torch.bincount(src, minlength=num_nodes), 0)
    return {
        "indptr": indptr.to(torch.int32),
        "csr_src": src[perm].to(torch.int32),
        "csr_dst": dst[perm].to(torch.int32),
        "csr_eid": perm.to(torch.int32),
        "outptr": outptr.to(torch.int32),
        "csc_pos": inv[perm_s].to(torch.int32),
    }


# --------------------------------------------------------------------------------------
# EGNNConv  (reference call sites: models/hybrid_models.py:29-31,89-90; upstream
# dgl/nn/pytorch/conv/egnnconv.py -- un-vendored, restated from SURVEY Appendix A.3)
# --------------------------------------------------------------------------------------
def silu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def egnn_conv(p: Params, prefix: str, src: Tensor, dst: Tensor, h: Tensor, x: Tensor,
              a: Tensor) -> Tuple[Tensor, Tensor]:
    """One ``dgl.nn.EGNNConv(in, 64, 64, edge_feat_size=1)`` layer.

    edge e = (s -> d):  diff = x_s - x_d ; radial = |diff|^2 ; diff /= sqrt(radial) + 1e-30
    msg_h = edge_mlp([h_s, h_d, radial, a_e]) ; msg_x = coord_mlp(msg_h) * diff
    node  d : h_neigh = sum msg_h ; x_neigh = mean msg_x (0 for in-degree 0)
    h' = node_mlp([h, h_neigh]) ; x' = x + x_neigh.
    """
    src = src.long()
    dst = dst.long()
    n = h.shape[0]
    diff = x[src] - x[dst]
    radial = (diff * diff).sum(-1, keepdim=True)
    diff = diff / (radial.sqrt() + 1e-30)
    f = torch.cat([h[src], h[dst], radial, a], dim=-1)
    t = silu(F.linear(f, p[prefix + "edge_mlp.0.weight"], p[prefix + "edge_mlp.0.bias"]))
    msg_h = silu(F.linear(t, p[prefix + "edge_mlp.2.weight"], p[prefix + "edge_mlp.2.bias"]))
    u = silu(F.linear(msg_h, p[prefix + "coord_mlp.0.weight"], p[prefix + "coord_mlp.0.bias"]))
    msg_x = F.linear(u, p[prefix + "coord_mlp.2.weight"]) * diff
    h_neigh = torch.zeros(n, msg_h.shape[1], dtype=h.dtype).index_add_(0, dst, msg_h)
    x_sum = torch.zeros(n, 3, dtype=h.dtype).index_add_(0, dst, msg_x)
    deg = torch.bincount(dst, minlength=n).clamp(min=1).to(h.dtype).unsqueeze(-1)
    x_neigh = x_sum / deg
    t5 = silu(F.linear(torch.cat([h, h_neigh], -1), p[prefix + "node_mlp.0.weight"],
                       p[prefix + "node_mlp.0.bias"]))
    h_out = F.linear(t5, p[prefix + "node_mlp.2.weight"], p[prefix + "node_mlp.2.bias"])
    return h_out, x + x_neigh


def egnn_stack(p: Params, src: Tensor, dst: Tensor, x23: Tensor, a: Tensor,
               n_layers: int, return_all: bool = False):
    """The ``for layer in self.GCN_layers`` loop of hybrid_models.py:82,89-90 / :316,323-324."""
    h, x = x23[:, :20], x23[:, 20:]
    hs = []
    for l in range(n_layers):
        h, x = egnn_conv(p, f"GCN_layers.{l}.", src, dst, h, x, a)
        hs.append((h, x))
    return (h, x, hs) if return_all else (h, x)


# --------------------------------------------------------------------------------------
# Pooling (torch_geometric.nn.global_mean_pool / global_max_pool; Appendix A.4)
# --------------------------------------------------------------------------------------
def global_mean_pool(x: Tensor, batch: Tensor) -> Tensor:
    nb = int(batch.max()) + 1
    s = torch.zeros(nb, x.shape[1], dtype=x.dtype).index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=nb).clamp(min=1).to(x.dtype).unsqueeze(-1)
    return s / cnt


def global_max_pool(x: Tensor, batch: Tensor) -> Tensor:
    nb = int(batch.max()) + 1
    out = torch.full((nb, x.shape[1]), -float("inf"), dtype=x.dtype)
    return out.scatter_reduce(0, batch.unsqueeze(-1).expand_as(x), x, reduce="amax")


# --------------------------------------------------------------------------------------
# Attention blocks (reference: immunostruct/models/layers.py)
# --------------------------------------------------------------------------------------
def self_attention(p: Params, prefix: str, x: Tensor) -> Tuple[Tensor, Tensor]:
    """``SelfAttention.forward`` layers.py:13-22 (single head, no output projection)."""
    q = F.linear(x, p[prefix + "query.weight"], p[prefix + "query.bias"])
    k = F.linear(x, p[prefix + "key.weight"], p[prefix + "key.bias"])
    v = F.linear(x, p[prefix + "value.weight"], p[prefix + "value.bias"])
    w = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(k.shape[-1]), dim=-1)
    return w @ v, w


def multi_head_attention(p: Params, prefix: str, x: Tensor, n_head: int) -> Tuple[Tensor, Tensor]:
    """``MultiHeadAttention.forward`` layers.py:67-78 with ``ScaleDotProductAttention`` :29-48."""
    b, l, _ = x.shape
    q = F.linear(x, p[prefix + "w_q.weight"], p[prefix + "w_q.bias"])
    k = F.linear(x, p[prefix + "w_k.weight"], p[prefix + "w_k.bias"])
    v = F.linear(x, p[prefix + "w_v.weight"], p[prefix + "w_v.bias"])
    d = q.shape[-1] // n_head

    def split(t):  # layers.py:80-93
        return t.reshape(b, l, n_head, d).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    w = torch.softmax(q @ k.transpose(2, 3) / math.sqrt(d), dim=-1)
    o = (w @ v).transpose(1, 2).reshape(b, l, n_head * d)  # layers.py:95-106
    return F.linear(o, p[prefix + "w_concat.weight"], p[prefix + "w_concat.bias"]), w


# --------------------------------------------------------------------------------------
# Dense branches shared by the hybrid models (hybrid_models.py:46-74 / 280-308)
# --------------------------------------------------------------------------------------
def property_embedding(p: Params, prop: Tensor) -> Tensor:
    """Linear(2,32)-ReLU-Dropout(eval: identity)-Linear(32,8)-ReLU (hybrid_models.py:280-286)."""
    t = F.relu(F.linear(prop, p["property_embedding.0.weight"], p["property_embedding.0.bias"]))
    return F.relu(F.linear(t, p["property_embedding.3.weight"], p["property_embedding.3.bias"]))


def vae_branch(p: Params, seq: Tensor, prop_emb: Tensor, eps: Tensor):
    """encode_vae / reparameterize / decode_vae (hybrid_models.py:297-308,337-340); ``eps`` is the
    ``randn_like`` draw, injected so that oracle and product consume the same noise."""
    flat = seq.reshape(seq.shape[0], -1)
    h1 = F.relu(F.linear(flat, p["vae_fc1.weight"], p["vae_fc1.bias"]))
    mu = F.linear(h1, p["vae_fc21.weight"], p["vae_fc21.bias"])
    logvar = F.linear(h1, p["vae_fc22.weight"], p["vae_fc22.bias"])
    z = torch.cat([mu + eps * torch.exp(0.5 * logvar), prop_emb], dim=1)
    h3 = F.relu(F.linear(z, p["vae_fc3.weight"], p["vae_fc3.bias"]))
    recon = F.linear(h3, p["vae_fc4.weight"], p["vae_fc4.bias"])
    return recon, mu, logvar, z


def classifier(p: Params, x: Tensor, prefix: str = "classifier.", with_out: bool = True) -> Tensor:
    """Flatten-Linear(.,32)-ReLU-Dropout(eval)-[Linear(32,1)] (hybrid_models.py:288-295)."""
    t = F.relu(F.linear(x.flatten(1), p[prefix + "1.weight"], p[prefix + "1.bias"]))
    if with_out:
        t = F.linear(t, p[prefix + "4.weight"], p[prefix + "4.bias"])
    return t


def fusion_attention(p: Params, combined: Tensor, n_head: int = 8) -> Tensor:
    """``combined_attention`` over the fused scalars + ``mean(dim=2)`` (hybrid_models.py:344-347)."""
    out, _ = multi_head_attention(p, "combined_attention.", combined.unsqueeze(2), n_head)
    return out.mean(dim=2)


# --------------------------------------------------------------------------------------
# Structure trunk: EGNN stack -> per-graph attention -> mean pool
# --------------------------------------------------------------------------------------
def structure_trunk(p: Params, g: dict, n_layers: int, attention: str, n_head: int = 1):
    """hybrid_models.py:316-331 (v2) / :82-97 (v1).  Requires equal node count per graph, exactly
    like the reference's ``view(B, -1, 64)``; padded nodes take part in attention and the mean."""
    h, _ = egnn_stack(p, g["src"], g["dst"], g["x"], g["edge_attr"], n_layers)
    b = int(g["batch_num_nodes"].numel())
    hb = h.reshape(b, -1, h.shape[-1])
    if attention == "mha":
        out, w = multi_head_attention(p, "self_attention.", hb, n_head)
    else:
        out, w = self_attention(p, "self_attention.", hb)
    flat = out.reshape(-1, out.shape[-1])
    bvec = batch_vector(g["batch_num_nodes"])
    return flat, bvec, w


def hybrid_forward(p: Params, g: dict, seq: Tensor, prop: Tensor, eps: Tensor, *,
                   version: str = "v2", n_layers: int = 6, self_heads: int = 1,
                   fusion_heads: int = 8, return_embedding=False, return_attention=False):
    """``HybridModelv2.forward`` (hybrid_models.py:315-359) or ``HybridModel.forward`` (:81-119),
    eval-mode dropout, ``eps`` injected for ``randn_like``."""
    flat, bvec, w = structure_trunk(p, g, n_layers, "mha" if version == "v2" else "sa", self_heads)
    pooled = global_mean_pool(flat, bvec)
    pe = property_embedding(p, prop)
    recon, mu, logvar, z = vae_branch(p, seq, pe, eps)
    combined = torch.cat([pooled, z], dim=1)
    if version == "v2":
        combined = fusion_attention(p, combined, fusion_heads)
    out = classifier(p, combined)
    if return_embedding:
        return pooled, mu, logvar, out
    if return_attention:
        return w, mu, logvar, out
    return recon, mu, logvar, out


def comparative_forward(p: Params, g_pair, seq_pair, prop_pair, eps_pair, *, n_layers: int = 6,
                        self_heads: int = 1, fusion_heads: int = 8, use_wt: bool = True):
    """``HybridModelv2_Comparative.forward_comparative`` (comparative_models.py:463-496):
    cancer item then wild-type item through the same trunk (``forward_item`` :433-461)."""
    embs, recons, mus, logvars = [], [], [], []
    for g, seq, prop, eps in zip(g_pair, seq_pair, prop_pair, eps_pair):
        flat, bvec, _ = structure_trunk(p, g, n_layers, "mha", self_heads)
        pooled = global_mean_pool(flat, bvec)
        pe = property_embedding(p, prop)
        recon, mu, logvar, z = vae_branch(p, seq, pe, eps)
        embs.append(torch.cat([pooled, z], 1))
        recons.append(recon)
        mus.append(mu)
        logvars.append(logvar)
    combined = torch.cat(embs, 1) if use_wt else embs[0]
    out = classifier(p, fusion_attention(p, combined, fusion_heads))
    return embs, recons, mus, logvars, out


def comparative_single_forward(p: Params, g, seq, prop, eps, *, n_layers: int = 6, self_heads: int = 1,
                               fusion_heads: int = 8, use_wt: bool = True):
    """``HybridModelv2_Comparative.forward`` ("repeat features hot fix", comparative_models.py:498-527)."""
    flat, bvec, _ = structure_trunk(p, g, n_layers, "mha", self_heads)
    pooled = global_mean_pool(flat, bvec)
    pe = property_embedding(p, prop)
    recon, mu, logvar, z = vae_branch(p, seq, pe, eps)
    emb = torch.cat([pooled, z], 1)
    combined = torch.cat([emb, emb], 1) if use_wt else emb
    out = classifier(p, fusion_attention(p, combined, fusion_heads))
    return recon, mu, logvar, out


# --------------------------------------------------------------------------------------
# Losses (reference: immunostruct/utils/loss.py:13-61)
# --------------------------------------------------------------------------------------
def kld(mu: Tensor, logvar: Tensor) -> Tensor:
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def bce_loss(recon, seq, mu, logvar, out, y, pos_weight, sequence=True, vae_input_dim=None):
    """``Losses.BCE_loss`` loss.py:23-31."""
    bce = F.binary_cross_entropy_with_logits(out.view(-1), y.view(-1),
                                             pos_weight=torch.as_tensor(pos_weight, dtype=out.dtype))
    if not sequence:
        return bce
    mse = F.mse_loss(recon, seq.reshape(recon.shape[0], -1))
    return 5.0 * bce + 0.1 * mse + 0.1 * kld(mu, logvar)


def regression_loss(recon, seq, mu, logvar, out, y, sequence=True):
    """``Losses.regression_loss`` loss.py:13-21 (note ``squeeze`` on both operands)."""
    reg = F.mse_loss(out.squeeze(), y.squeeze())
    if not sequence:
        return reg
    mse = F.mse_loss(recon, seq.reshape(recon.shape[0], -1))
    return 2.0 * reg + 0.5 * mse + 0.5 * kld(mu, logvar)


# --------------------------------------------------------------------------------------
# Paired contrastive loss (reference: immunostruct/utils/contrastive.py:37-83)
# --------------------------------------------------------------------------------------
def paired_contrastive(p: Params, emb_c: Tensor, emb_w: Tensor, target: Tensor,
                       z_dim: int = 128, lambda_off: float = 1e-2, bn_eps: float = 1e-5):
    """Projector (Linear-BatchNorm1d[batch statistics]-ReLU-Linear, no biases), mean-centre,
    std hinge, pair-similarity vs diag(is_immunogenic), cross-correlation vs I, off-diagonals
    down-weighted.  Returns python ``0`` unless the batch holds exactly two target values."""
    if torch.unique(target).numel() != 2:
        return 0
    imm = target > target.mean()

    def project(e):
        t = F.linear(e, p["projector.0.weight"])
        m, v = t.mean(0), t.var(0, unbiased=False)
        t = (t - m) / torch.sqrt(v + bn_eps) * p["projector.1.weight"] + p["projector.1.bias"]
        return F.linear(F.relu(t), p["projector.3.weight"])

    zc, zw = project(emb_c), project(emb_w)
    b = zc.shape[0]
    zc = zc - zc.mean(0)
    zw = zw - zw.mean(0)
    std_loss = (F.relu(1 - torch.sqrt(zc.var(0) + 1e-4)).mean() / 2
                + F.relu(1 - torch.sqrt(zw.var(0) + 1e-4)).mean() / 2)
    pair = zc @ zw.T / z_dim
    cross = zc.T @ zw / b
    eye_b = torch.eye(b, dtype=zc.dtype)
    ideal = eye_b * imm.to(zc.dtype).unsqueeze(1)
    pd = (pair - ideal).pow(2)
    pd = torch.where(eye_b.bool(), pd, pd * lambda_off)
    eye_z = torch.eye(z_dim, dtype=zc.dtype)
    cd = (cross - eye_z).pow(2)
    cd = torch.where(eye_z.bool(), cd, cd * lambda_off)
    return pd.sum() + cd.sum() + std_loss
