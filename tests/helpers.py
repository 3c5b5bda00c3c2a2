"""Shared test helpers: build product inputs from golden groups, inject eps, collect grads."""
import types

import torch

import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch


def graph_batch(arrays, device="cpu"):
    """golden/synthetic 'graph' group (graph-local endpoints) -> GraphBatch on ``device``."""
    gb = GraphBatch.from_arrays(arrays["x"].float(), arrays["src"], arrays["dst"], arrays["edge_attr"].float(),
                                arrays["node_counts"], arrays["edge_counts"],
                                max_nodes=int(arrays["node_counts"].max()))
    if torch.device(device).type == "cpu":
        gb._collate()          # CPU: only reachable with the kernel contracts patched in (tests)
        return gb
    return gb.to(device)


def inject_eps(model, *eps_list):
    """Make the model's reparameterisation consume the given noise tensors in order (instead of randn_like);
    ``sample_eps`` feeds both ``reparameterize`` and the fused eval-mode kernel."""
    it = iter(eps_list)

    def sample_eps(self, like):
        return next(it).to(like.device, like.dtype)

    model.sample_eps = types.MethodType(sample_eps, model)
    return model


def named_grads(model):
    return {k: p.grad for k, p in model.named_parameters()}


def build_model(name, golden, device="cpu", **kw):
    meta = golden["meta"]
    model = I.model_map[name](vae_input_dim=231, device=device, gcn_layers=int(meta["gcn_layers"]) - 1,
                              vae_hidden_dim=int(meta.get("vae_hidden_dim", torch.tensor(32))), **kw)
    missing = model.load_state_dict(golden["weights"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.to(device).eval()
