"""torch.library registration (immunostruct_b200/ops.py): schema / fake-tensor / autograd-registration checks of every
registered operator (torch.library.opcheck), and the whole model routed through ``torch.ops.immunostruct_b200.*`` against
the default autograd.Function path -- on CPU over the kernel contracts, and (marked gpu) on the device."""
import pytest
import torch

import immunostruct_b200 as I
from immunostruct_b200 import _C, ops, trunk
from oracle import kernel_contracts as KC

from conftest import load_golden, rel_err
from helpers import build_model, graph_batch, inject_eps, named_grads


@pytest.fixture
def cpu_backend(monkeypatch):
    for name in KC.ALL:
        monkeypatch.setattr(_C, name, getattr(KC, name))
    monkeypatch.setattr(trunk, "_require_device_batch", lambda g: None)
    yield
    I.use_custom_ops(False)


def _model_run(device, custom, backward=True):
    gd = load_golden("hybrid_v2")
    model = build_model("HybridModelv2", gd, device=device)
    g = graph_batch(gd["graph"], device)
    d = {k: v.to(device) for k, v in gd["dense"].items()}
    inject_eps(model, d["eps"])
    I.use_custom_ops(custom)
    try:
        recon, mu, logvar, out = model(g, d["seq"], d["prop"])
        loss = I.Losses(231, [float(gd["meta"]["pos_weight"]), 1.0], sequence=True).BCE_loss(recon, d["seq"], mu, logvar, out, d["target"])
        assert loss.requires_grad and loss.grad_fn is not None
        if backward:
            loss.backward()
    finally:
        I.use_custom_ops(False)
    return out.detach(), loss.detach(), named_grads(model), gd


def _check_model(device, backward=True):
    out_a, loss_a, grads_a, gd = _model_run(device, False, backward)
    out_b, loss_b, grads_b, _ = _model_run(device, True, backward)
    assert rel_err(out_b, gd["out"]["logits"]) < 2e-5 and rel_err(loss_b, gd["out"]["loss_bce"]) < 2e-5
    assert torch.equal(out_a, out_b) and torch.equal(loss_a, loss_b)          # same launchers, same arithmetic
    if not backward:
        return
    for k, ga in grads_a.items():
        gb = grads_b[k]
        assert (ga is None) == (gb is None), k
        if ga is not None:
            assert torch.equal(ga, gb), k
    assert grads_b["GCN_layers.5.coord_mlp.0.weight"] is None                 # still no gradient, not zeros


def test_model_through_registered_ops_matches_default_path(cpu_backend):
    # forward only on CPU: the CPU contracts derive gradients with autograd, which is switched off inside a registered
    # operator's implementation (the device kernels have no such dependence: the gpu test below runs the backward too)
    _check_model("cpu", backward=False)


@pytest.mark.gpu
def test_model_through_registered_ops_matches_default_path_gpu():
    _check_model("cuda")


def _opcheck_all(device, tests=("test_schema", "test_faketensor", "test_autograd_registration")):
    gen = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=gen).to(device)
    off = torch.tensor([0, 5, 5, 12], device=device)
    torch.library.opcheck(ops.segment_pool, (r(12, 64).requires_grad_(True), off, "max"), test_utils=tests)
    torch.library.opcheck(ops.linear, (r(9, 40).requires_grad_(True), r(16, 40).requires_grad_(True), r(16).requires_grad_(True), True),
                          test_utils=tests)
    torch.library.opcheck(ops.fusion_attention, (r(4, 104).requires_grad_(True), r(33).requires_grad_(True), 8), test_utils=tests)
    torch.library.opcheck(ops.fused_loss, (r(4, 231).requires_grad_(True), r(4, 231), r(4, 32).requires_grad_(True),
                                           r(4, 32).requires_grad_(True), r(4, 1).requires_grad_(True), torch.tensor([0., 1, 0, 1], device=device),
                                           0, 2.0, 5.0, 0.1, 0.1), test_utils=tests)
    gd = load_golden("hybrid_v2")
    g = graph_batch(gd["graph"], device)
    qkv = r(g.n_nodes, 192).requires_grad_(True)
    torch.library.opcheck(ops.attention_pool, (qkv, g.node_off, 1, int(g.max_nodes)), test_utils=tests)
    model = build_model("HybridModelv2", gd, device=device)
    flat = [t for l in model.GCN_layers for t in l.kernel_params()]
    torch.library.opcheck(ops.egnn_stack, (g.ndata["x"], g.edata["edge_attr"], flat, ops.graph_tensors(g), len(model.GCN_layers),
                                           g.n_edges, g.n_graphs, int(g.max_nodes), []), test_utils=tests)
    # with the attention projections fused into the last node kernel (QKV is the operator's second output)
    torch.library.opcheck(ops.egnn_stack, (g.ndata["x"], g.edata["edge_attr"], flat, ops.graph_tensors(g), len(model.GCN_layers),
                                           g.n_edges, g.n_graphs, int(g.max_nodes), list(model.self_attention.qkv_params())),
                          test_utils=tests)


def test_registered_ops_pass_opcheck(cpu_backend):
    _opcheck_all("cpu", tests=("test_schema", "test_faketensor"))


@pytest.mark.gpu
def test_registered_ops_pass_opcheck_gpu():
    _opcheck_all("cuda")
