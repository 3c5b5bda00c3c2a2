"""Generate the committed golden vectors from the UNMODIFIED reference model code.

Run in the build container (needs ``/root/reference``):  ``python tests/golden/make_golden.py``.
The reference's own ``models/*.py``, ``utils/loss.py`` and ``utils/contrastive.py`` are imported
through ``oracle/shim.py`` (stand-ins only for the absent dgl / torch_geometric packages) and
run on seeded synthetic pMHC batches; inputs, weights, the injected ``randn_like`` noise, outputs,
losses and parameter gradients are written to ``tests/golden/*.npz``.  The GPU box has no
``/root/reference``; tests there read these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import shim  # noqa: E402
from oracle import reference_ops as R  # noqa: E402
from immunostruct_b200.synthetic import synthetic_graph_arrays, split_graphs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEQ = (11, 21)            # vae_input_dim = 231 (the reference's peptide-only setting)


class _Eps:
    """Replace ``torch.randn_like`` by a fixed sequence while the reference model runs."""

    def __init__(self, draws):
        self.draws, self.i = draws, 0

    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = self
        return self

    def __call__(self, t, **kw):
        e = self.draws[self.i].to(t.dtype)
        self.i += 1
        return e

    def __exit__(self, *a):
        torch.randn_like = self.orig


def make_inputs(seed, b, n, k, n_pad):
    arr = synthetic_graph_arrays(b, n, k, seed=seed, n_pad=n_pad, coord_scale=3.0)
    g = R.dgl_batch(split_graphs(arr))
    gen = torch.Generator().manual_seed(seed + 1)
    tok = torch.randint(0, SEQ[1], (b, SEQ[0]), generator=gen)
    seq = torch.nn.functional.one_hot(tok, SEQ[1]).float()
    prop = torch.rand(b, 2, generator=gen)
    y = (torch.arange(b) % 2).float()
    eps = torch.randn(b, 32, generator=gen)
    return arr, g, seq, prop, y, eps


def scaled_init(model, seed, scale):
    torch.manual_seed(seed)
    for m in model.modules():
        if hasattr(m, "reset_parameters"):
            m.reset_parameters()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith("GCN_layers") and name.endswith("weight"):
                p.mul_(scale)
    return model


def pack(prefix, d):
    # fp64 reference values are stored rounded to fp32 (6e-8 relative: far below the 1e-5 tolerance)
    cast = (lambda v: v.float()) if prefix.startswith("grads64") else (lambda v: v)
    return {f"{prefix}{k}": (cast(v.detach().cpu()).numpy() if torch.is_tensor(v) else np.asarray(v))
            for k, v in d.items()}


def save(name, **groups):
    flat = {}
    for gname, d in groups.items():
        flat.update(pack(gname + "/", d))
    np.savez_compressed(os.path.join(OUT, name), **flat)
    print(name, f"{os.path.getsize(os.path.join(OUT, name)) / 1e6:.2f} MB", len(flat), "arrays")


def graph_inputs(arr):
    return {k: arr[k] for k in ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")}


def hybrid_case(model_map, Losses, cls, fname, gcn_layers, seed):
    b, n, k, n_pad = 3, 12, 4, 2
    arr, g, seq, prop, y, eps = make_inputs(seed, b, n, k, n_pad)
    model = model_map[cls](vae_input_dim=SEQ[0] * SEQ[1], device="cpu", gcn_layers=gcn_layers,
                           vae_hidden_dim=32)
    scaled_init(model, seed, 1.6).eval()
    losses = Losses(SEQ[0] * SEQ[1], [2.0, 1.0], sequence=True)
    sg = shim.graph_from_dict(g)
    trunk = {}
    hook = model.GCN_layers[-1].register_forward_hook(
        lambda m, i, o: trunk.update(h_last=o[0].detach().clone(), x_last=o[1].detach().clone()))
    with _Eps([eps]):
        recon, mu, logvar, out = model(sg, seq, prop)
    hook.remove()
    loss = losses.BCE_loss(recon, seq, mu, logvar, out, y)
    loss_reg = losses.regression_loss(recon, seq, mu, logvar, out, y * 0.5 - 0.1)
    model.zero_grad()
    loss.backward()
    grads = {k_: (p.grad if p.grad is not None else torch.full((1,), float("nan")))
             for k_, p in model.named_parameters()}
    with _Eps([eps]):
        emb = model(sg, seq, prop, return_embedding=True)[0]
    with _Eps([eps]):
        attn = model(sg, seq, prop, return_attention=True)[0]
    # fp64 run of the same reference code: the tolerance yardstick
    m64 = model_map[cls](vae_input_dim=SEQ[0] * SEQ[1], device="cpu", gcn_layers=gcn_layers,
                         vae_hidden_dim=32).double().eval()
    m64.load_state_dict({k_: v.double() for k_, v in model.state_dict().items()})
    g64 = dict(g, x=g["x"].double(), edge_attr=g["edge_attr"].double())
    with _Eps([eps]):
        r64 = m64(shim.graph_from_dict(g64), seq.double(), prop.double())
    losses.BCE_loss(r64[0], seq.double(), r64[1], r64[2], r64[3], y.double()).backward()
    grads64 = {k_: (p.grad if p.grad is not None else torch.full((1,), float("nan")))
               for k_, p in m64.named_parameters()}
    save(fname,
         graph=graph_inputs(arr), dense={"seq": seq, "prop": prop, "target": y, "eps": eps},
         weights=dict(model.state_dict()),
         out={"recon": recon, "mu": mu, "logvar": logvar, "logits": out, "embedding": emb,
              "attention": attn, "h_last": trunk["h_last"], "x_last": trunk["x_last"],
              "loss_bce": loss, "loss_reg": loss_reg, "logits64": r64[3], "recon64": r64[0]},
         grads=grads, grads64=grads64,
         meta={"gcn_layers": gcn_layers + 1, "pos_weight": 2.0, "vae_hidden_dim": 32,
               "n_nodes": n, "k": k})


def comparative_case(model_map, Losses, PCL, fname, seed):
    b, n, k, n_pad = 4, 10, 3, 1
    arr_c, g_c, seq_c, prop_c, y, eps_c = make_inputs(seed, b, n, k, n_pad)
    arr_w, g_w, seq_w, prop_w, _, eps_w = make_inputs(seed + 100, b, n, k, n_pad)
    model = model_map["HybridModelv2_Comparative"](vae_input_dim=SEQ[0] * SEQ[1], device="cpu",
                                                   gcn_layers=1, vae_hidden_dim=32)
    scaled_init(model, seed, 1.6).eval()
    torch.manual_seed(seed + 5)
    pcl = PCL(embedding_dim=104)
    pcl_state = {k_: v.clone() for k_, v in pcl.state_dict().items()}     # before BatchNorm running stats move
    losses = Losses(SEQ[0] * SEQ[1], [3.0, 1.0], sequence=True)
    with _Eps([eps_c, eps_w]):
        embs, recons, mus, logvars, out = model.forward_comparative(
            (shim.graph_from_dict(g_c), shim.graph_from_dict(g_w)), (seq_c, seq_w), (prop_c, prop_w))
    l_c = losses.BCE_loss(recons[0], seq_c, mus[0], logvars[0], out, y)
    l_w = losses.BCE_loss(recons[1], seq_w, mus[1], logvars[1], out, y)
    l_con = pcl(embs[0], embs[1], y)
    loss = (l_c + l_w) / 2 + 0.01 * l_con       # procedures/train.py:107-118
    model.zero_grad()
    loss.backward()
    grads = {k_: (p.grad if p.grad is not None else torch.full((1,), float("nan")))
             for k_, p in model.named_parameters()}
    with _Eps([eps_c]):
        single = model(shim.graph_from_dict(g_c), seq_c, prop_c)
    # fp64 run of the same reference code (model, losses and contrastive module in double)
    m64 = model_map["HybridModelv2_Comparative"](vae_input_dim=SEQ[0] * SEQ[1], device="cpu",
                                                 gcn_layers=1, vae_hidden_dim=32).double().eval()
    m64.load_state_dict({k_: v.double() for k_, v in model.state_dict().items()})
    pcl64 = PCL(embedding_dim=104).double()
    pcl64.load_state_dict({k_: v.double() if v.is_floating_point() else v for k_, v in pcl_state.items()})
    dd = lambda gg: shim.graph_from_dict(dict(gg, x=gg["x"].double(), edge_attr=gg["edge_attr"].double()))
    with _Eps([eps_c, eps_w]):
        e64, r64, mu64, lv64, o64 = m64.forward_comparative((dd(g_c), dd(g_w)), (seq_c.double(), seq_w.double()),
                                                            (prop_c.double(), prop_w.double()))
    y64 = y.double()
    l64 = (losses.BCE_loss(r64[0], seq_c.double(), mu64[0], lv64[0], o64, y64)
           + losses.BCE_loss(r64[1], seq_w.double(), mu64[1], lv64[1], o64, y64)) / 2 + 0.01 * pcl64(e64[0], e64[1], y64)
    l64.backward()
    grads64 = {k_: (p.grad if p.grad is not None else torch.full((1,), float("nan")))
               for k_, p in m64.named_parameters()}
    save(fname,
         graph_c=graph_inputs(arr_c), graph_w=graph_inputs(arr_w),
         dense={"seq_c": seq_c, "seq_w": seq_w, "prop_c": prop_c, "prop_w": prop_w, "target": y,
                "eps_c": eps_c, "eps_w": eps_w},
         weights=dict(model.state_dict()), projector=pcl_state,
         out={"emb_c": embs[0], "emb_w": embs[1], "recon_c": recons[0], "recon_w": recons[1],
              "mu_c": mus[0], "mu_w": mus[1], "logvar_c": logvars[0], "logvar_w": logvars[1],
              "logits": out, "loss_contrastive": l_con, "loss": loss,
              "single_logits": single[3], "single_recon": single[0]},
         grads=grads, grads64=grads64, meta={"gcn_layers": 2, "pos_weight": 3.0, "coeff_contrastive": 0.01})


def structure_case(model_map, fname, seed):
    """StructureModelv2: 8-head per-graph attention, mean+max pooling, SSL heads (ablation_models.py:244-307)."""
    b, n, k, n_pad = 3, 9, 3, 1
    arr, g, seq, prop, y, eps = make_inputs(seed, b, n, k, n_pad)
    model = model_map["StructureModelv2"](vae_input_dim=SEQ[0] * SEQ[1], device="cpu", gcn_layers=1)
    scaled_init(model, seed, 1.6).eval()
    _, _, _, out, node_pred = model(shim.graph_from_dict(g), seq, prop)
    (out.sum() + node_pred.pow(2).sum()).backward()
    grads = {k_: (p.grad if p.grad is not None else torch.full((1,), float("nan")))
             for k_, p in model.named_parameters()}
    save(fname, graph=graph_inputs(arr), dense={"seq": seq, "prop": prop},
         weights=dict(model.state_dict()), out={"logits": out, "node_pred": node_pred},
         grads=grads, meta={"gcn_layers": 2})


def state_dict_shapes(model_map):
    import json
    out = {}
    for name, cls in model_map.items():
        m = cls(vae_input_dim=5943, device="cpu")
        out[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_shapes.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("state_dict_shapes.json", len(out), "classes")


def main():
    torch.set_num_threads(1)
    model_map, Losses, PCL = shim.load_reference()
    state_dict_shapes(model_map)
    hybrid_case(model_map, Losses, "HybridModelv2", "hybrid_v2.npz", gcn_layers=5, seed=1)
    hybrid_case(model_map, Losses, "HybridModel", "hybrid_v1.npz", gcn_layers=1, seed=2)
    comparative_case(model_map, Losses, PCL, "comparative_v2.npz", seed=3)
    structure_case(model_map, "structure_v2.npz", seed=4)


if __name__ == "__main__":
    main()
