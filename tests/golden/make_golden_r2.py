"""Round-2 golden vectors: the remaining model classes, one case at the production width (vae_input_dim = 5943,
vae_hidden_dim = 512) and a TRAIN-MODE trace (dropout live, RNG draw order, two optimizer steps).

Run in the build container (needs ``/root/reference``):  ``python tests/golden/make_golden_r2.py``.
Like make_golden.py, the reference's own ``models/*.py`` / ``utils/*.py`` are imported UNMODIFIED through
``oracle/shim.py`` and run on seeded synthetic pMHC batches.  Files written: ``tests/golden/r2_<class>.npz``,
``r2_big_hybrid_v2.npz`` (no weights inside: they are regenerated from a name-keyed seed on both sides, see
``seeded_state_dict``), ``r2_train_mode.npz``.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import shim  # noqa: E402
from oracle import reference_ops as R  # noqa: E402
from immunostruct_b200.synthetic import synthetic_graph_arrays, split_graphs  # noqa: E402
from golden_util import flatten_outputs, objective, sample_big, seeded_state_dict  # noqa: E402
from golden.make_golden import _Eps, graph_inputs, save  # noqa: E402

SEQ = (11, 21)


def inputs(seed, b, n, k, n_pad, seq=SEQ):
    arr = synthetic_graph_arrays(b, n, k, seed=seed, n_pad=n_pad, coord_scale=3.0)
    g = R.dgl_batch(split_graphs(arr))
    gen = torch.Generator().manual_seed(seed + 1)
    tok = torch.randint(0, seq[1], (b, seq[0]), generator=gen)
    return arr, g, torch.nn.functional.one_hot(tok, seq[1]).float(), torch.rand(b, 2, generator=gen), \
        torch.randn(b, 32, generator=gen)


def grads_of(model):
    return {k: (p.grad if p.grad is not None else torch.full((1,), float("nan"))) for k, p in model.named_parameters()}


def class_case(model_map, cls, seed, pair=False, **kw):
    """Forward (eval mode, injected noise) + gradient of a fixed weighted sum of every tensor output."""
    b, n, k, n_pad = 3, 10, 3, 1
    ctor = dict(vae_input_dim=SEQ[0] * SEQ[1], device="cpu", gcn_layers=1, vae_hidden_dim=32)
    ctor.update(kw)

    def run(dtype):
        model = model_map[cls](**ctor).to(dtype).eval()
        model.load_state_dict({k_: v.to(dtype) for k_, v in seeded_state_dict(model, seed).items()})
        cast = lambda g_: shim.graph_from_dict(dict(g_, x=g_["x"].to(dtype), edge_attr=g_["edge_attr"].to(dtype)))
        if pair:
            (arr_c, g_c, s_c, p_c, e_c), (arr_w, g_w, s_w, p_w, e_w) = inputs(seed, b, n, k, n_pad), inputs(seed + 50, b, n, k, n_pad)
            with _Eps([e_c, e_w]):
                outs = model.forward_comparative((cast(g_c), cast(g_w)), (s_c.to(dtype), s_w.to(dtype)), (p_c.to(dtype), p_w.to(dtype)))
            extra = dict(graph_c=graph_inputs(arr_c), graph_w=graph_inputs(arr_w),
                         dense={"seq_c": s_c, "seq_w": s_w, "prop_c": p_c, "prop_w": p_w, "eps_c": e_c, "eps_w": e_w})
        else:
            arr, g, s, p, e = inputs(seed, b, n, k, n_pad)
            with _Eps([e]):
                outs = model(cast(g), s.to(dtype), p.to(dtype))
            extra = dict(graph=graph_inputs(arr), dense={"seq": s, "prop": p, "eps": e})
        flat = flatten_outputs(outs)
        objective(flat).backward()
        return model, flat, extra

    model, flat, extra = run(torch.float32)
    m64, flat64, _ = run(torch.float64)
    save(f"r2_{cls}.npz", **extra, out={f"o{i}": t for i, t in enumerate(flat)},
         out64={f"o{i}": t.float() for i, t in enumerate(flat64)}, grads=grads_of(model), grads64=grads_of(m64),
         meta={"seed": seed, "n_out": len(flat), "pair": int(pair)})


def big_case(model_map, Losses, seed=21):
    """HybridModelv2 at vae_input_dim = 5943, vae_hidden_dim = 512 (the production width of vae_fc1 / vae_fc4); gradients
    of the two 3 M-parameter matrices are stored as strided samples + norms."""
    b, n, k, n_pad = 4, 24, 4, 2
    seq_shape = (283, 21)

    def run(dtype):
        model = model_map["HybridModelv2"](vae_input_dim=5943, device="cpu", gcn_layers=2).to(dtype).eval()
        model.load_state_dict({k_: v.to(dtype) for k_, v in seeded_state_dict(model, seed).items()})
        arr, g, s, p, e = inputs(seed, b, n, k, n_pad, seq_shape)
        y = (torch.arange(b) % 2).to(dtype)
        with _Eps([e]):
            recon, mu, logvar, out = model(shim.graph_from_dict(dict(g, x=g["x"].to(dtype), edge_attr=g["edge_attr"].to(dtype))),
                                           s.to(dtype), p.to(dtype))
        loss = Losses(5943, [2.0, 1.0], sequence=True).BCE_loss(recon, s.to(dtype), mu, logvar, out, y)
        loss.backward()
        return model, arr, s, p, e, y, recon, mu, logvar, out, loss

    model, arr, s, p, e, y, recon, mu, logvar, out, loss = run(torch.float32)
    m64, *_rest, loss64 = run(torch.float64)
    g32 = {k_: sample_big(v) for k_, v in grads_of(model).items()}
    g64 = {k_: sample_big(v.float()) for k_, v in grads_of(m64).items()}
    save("r2_big_hybrid_v2.npz", graph=graph_inputs(arr), dense={"seq_tokens": s.argmax(-1).to(torch.uint8), "prop": p, "eps": e, "target": y},
         out={"recon_sample": sample_big(recon), "mu": mu, "logvar": logvar, "logits": out, "loss": loss, "loss64": loss64.float(),
              "logits64": _rest[-1].float()},
         grads=g32, grads64=g64, meta={"seed": seed, "gcn_layers": 3, "pos_weight": 2.0})


def train_mode_case(model_map, Losses, seed=31):
    """Two optimizer steps of HybridModelv2 in TRAIN mode on the CPU generator: property-embedding dropout ->
    randn_like -> classifier dropout (hybrid_models.py:334,338,351).  Stored: per-step losses, logits and the final
    parameters.  The product, run on CPU through the kernel contracts with the same torch.manual_seed, must draw the
    same random numbers in the same order to reproduce them."""
    b, n, k, n_pad = 4, 10, 3, 1
    arr, g, s, p, _ = inputs(seed, b, n, k, n_pad)
    y = (torch.arange(b) % 2).float()
    model = model_map["HybridModelv2"](vae_input_dim=SEQ[0] * SEQ[1], device="cpu", gcn_layers=1, vae_hidden_dim=32).train()
    model.load_state_dict(seeded_state_dict(model, seed))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6)
    losses = Losses(SEQ[0] * SEQ[1], [2.0, 1.0], sequence=True)
    torch.manual_seed(seed)
    trace = {}
    for step in range(2):
        opt.zero_grad()
        recon, mu, logvar, out = model(shim.graph_from_dict(g), s, p)
        loss = losses.BCE_loss(recon, s, mu, logvar, out, y)
        loss.backward()
        opt.step()
        trace[f"loss{step}"], trace[f"logits{step}"], trace[f"mu{step}"] = loss.detach(), out.detach().clone(), mu.detach().clone()
    trace["rng_after"] = torch.rand(4)                      # the generator ends in the same state
    save("r2_train_mode.npz", graph=graph_inputs(arr), dense={"seq": s, "prop": p, "target": y}, out=trace,
         final={k_: v.detach().clone() for k_, v in model.state_dict().items()}, meta={"seed": seed})


def main():
    torch.set_num_threads(1)
    model_map, Losses, PCL = shim.load_reference()
    if "--big-only" in sys.argv:
        return big_case(model_map, Losses)
    for i, cls in enumerate(["HybridModel_SSL", "HybridModelv2_SSL", "DualModel", "SequenceModel", "SequenceFpModel",
                             "StructureModel", "StructureModel_SSL"]):
        class_case(model_map, cls, seed=40 + i)
    for i, cls in enumerate(["HybridModel_Comparative", "HybridModel_Comparative_SSL", "HybridModelv2_Comparative_SSL"]):
        class_case(model_map, cls, seed=60 + i, pair=True)
    big_case(model_map, Losses)
    train_mode_case(model_map, Losses)


if __name__ == "__main__":
    main()
