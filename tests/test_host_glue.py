"""Host logic of the product (autograd wiring, buffer plumbing, model classes, state_dict layout)
exercised on CPU: ``immunostruct_b200._C`` is monkeypatched with the per-kernel CPU contracts
(oracle/kernel_contracts.py), then the product models are compared with the golden vectors that the
unmodified reference produced.  The CUDA kernels themselves are checked against the same contracts in
tests/test_kernels_gpu.py."""
import json
import os

import pytest
import torch

import immunostruct_b200 as I
from immunostruct_b200 import _C, trunk
from oracle import kernel_contracts as KC
from oracle import reference_ops as R

from conftest import GOLDEN_DIR, assert_grads_close, load_golden, rel_err
from helpers import build_model, graph_batch, inject_eps, named_grads

TOL = 2e-5


@pytest.fixture
def cpu_backend(monkeypatch):
    for name in KC.ALL:
        monkeypatch.setattr(_C, name, getattr(KC, name))
    monkeypatch.setattr(trunk, "_require_device_batch", lambda g: None)
    yield


@pytest.mark.parametrize("name,cls", [("hybrid_v2", "HybridModelv2"), ("hybrid_v1", "HybridModel")])
def test_hybrid_model_matches_reference_golden(cpu_backend, name, cls):
    gd = load_golden(name)
    model = build_model(cls, gd)
    g = graph_batch(gd["graph"])
    d, o = gd["dense"], gd["out"]
    inject_eps(model, d["eps"], d["eps"], d["eps"])
    recon, mu, logvar, out = model(g, d["seq"], d["prop"])
    for got, key in ((recon, "recon"), (mu, "mu"), (logvar, "logvar"), (out, "logits")):
        assert rel_err(got, o[key]) < TOL, key
    losses = I.Losses(231, [float(gd["meta"]["pos_weight"]), 1.0], sequence=True)
    loss = losses.BCE_loss(recon, d["seq"], mu, logvar, out, d["target"])
    assert rel_err(loss, o["loss_bce"]) < TOL
    assert rel_err(losses.regression_loss(recon, d["seq"], mu, logvar, out, d["target"] * 0.5 - 0.1),
                   o["loss_reg"]) < TOL
    loss.backward()
    assert_grads_close(named_grads(model), gd["grads"], 1e-4)
    inject_eps(model, d["eps"], d["eps"], d["eps"])
    with torch.no_grad():                         # inference path: fused stack + pooled-only attention
        r_ng = model(g, d["seq"], d["prop"])
    assert rel_err(r_ng[3], o["logits"]) < TOL and rel_err(r_ng[0], o["recon"]) < TOL
    emb = model(g, d["seq"], d["prop"], return_embedding=True)[0]
    att = model(g, d["seq"], d["prop"], return_attention=True)[0]
    assert rel_err(emb, o["embedding"]) < TOL
    assert att.shape == o["attention"].shape and rel_err(att, o["attention"]) < TOL


def test_comparative_model_and_contrastive_match_reference_golden(cpu_backend):
    gd = load_golden("comparative_v2")
    model = build_model("HybridModelv2_Comparative", gd)
    gc, gw = graph_batch(gd["graph_c"]), graph_batch(gd["graph_w"])
    d, o = gd["dense"], gd["out"]
    inject_eps(model, d["eps_c"], d["eps_w"], d["eps_c"])
    embs, recons, mus, logvars, out = model.forward_comparative(
        (gc, gw), (d["seq_c"], d["seq_w"]), (d["prop_c"], d["prop_w"]))
    assert rel_err(out, o["logits"]) < TOL
    assert rel_err(embs[0], o["emb_c"]) < TOL and rel_err(embs[1], o["emb_w"]) < TOL
    losses = I.Losses(231, [float(gd["meta"]["pos_weight"]), 1.0], sequence=True)
    pcl = I.PairedContrastiveLoss(embedding_dim=104)
    pcl.load_state_dict(gd["projector"])
    l_c = losses.BCE_loss(recons[0], d["seq_c"], mus[0], logvars[0], out, d["target"])
    l_w = losses.BCE_loss(recons[1], d["seq_w"], mus[1], logvars[1], out, d["target"])
    l_con = pcl(embs[0], embs[1], d["target"])
    assert rel_err(l_con, o["loss_contrastive"]) < TOL
    loss = (l_c + l_w) / 2 + float(gd["meta"]["coeff_contrastive"]) * l_con      # procedures/train.py:107-118
    assert rel_err(loss, o["loss"]) < TOL
    loss.backward()
    assert_grads_close(named_grads(model), gd["grads"], 1e-4)
    single = model(gc, d["seq_c"], d["prop_c"])
    assert rel_err(single[3], o["single_logits"]) < TOL


def test_structure_model_v2_matches_reference_golden(cpu_backend):
    gd = load_golden("structure_v2")
    model = I.model_map["StructureModelv2"](vae_input_dim=231, device="cpu", gcn_layers=1)
    model.load_state_dict(gd["weights"])
    model.eval()
    g = graph_batch(gd["graph"])
    z0, z1, z2, out, node_pred = model(g, gd["dense"]["seq"], gd["dense"]["prop"])
    assert (z0, z1, z2) == (0, 0, 0)
    assert rel_err(out, gd["out"]["logits"]) < TOL and rel_err(node_pred, gd["out"]["node_pred"]) < TOL
    (out.sum() + node_pred.pow(2).sum()).backward()
    assert_grads_close(named_grads(model), gd["grads"], 1e-4)


def test_contrastive_gate_is_zero_without_two_classes(cpu_backend):
    pcl = I.PairedContrastiveLoss(embedding_dim=104)
    e1, e2 = torch.randn(6, 104, requires_grad=True), torch.randn(6, 104)
    for target in (torch.ones(6), torch.linspace(0, 1, 6)):
        loss = pcl(e1, e2, target)
        assert float(loss) == 0.0
        loss.backward()
        assert float(e1.grad.abs().max()) == 0.0


def test_fusion_closed_form_equals_dense_attention(cpu_backend):
    """The collapsed fusion attention == mean(MultiHeadAttention(x.unsqueeze(2))[0], dim=2)."""
    torch.manual_seed(3)
    for dim, length in ((16, 104), (32, 208)):
        mha = I.MultiHeadAttention(dim, 8, input_dim=1)
        c = torch.randn(5, length, requires_grad=True)
        dense = mha(c.unsqueeze(2))[0].mean(dim=2)
        fused = mha.fused_mean(c)
        assert rel_err(fused, dense) < 1e-5
        gd = torch.autograd.grad(dense.pow(2).sum(), [c] + list(mha.parameters()), retain_graph=True)
        gf = torch.autograd.grad(fused.pow(2).sum(), [c] + list(mha.parameters()))
        names = ["c"] + [n for n, _ in mha.named_parameters()]
        assert_grads_close(dict(zip(names, gf)), dict(zip(names, gd)), 1e-4)


def test_state_dict_layout_matches_reference_for_every_model():
    """Keys, order-insensitive, and shapes of all 14 model_map classes vs the reference's own classes
    (tests/golden/state_dict_shapes.json, written by make_golden.py)."""
    ref = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_shapes.json")))
    assert sorted(ref) == sorted(I.model_map)
    for name, shapes in ref.items():
        kw = {} if name == "DualModel" else {"use_wt_for_downstream": True} if "Comparative" in name else {}
        m = I.model_map[name](vae_input_dim=5943, device="cpu", **kw)
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == shapes, name


def test_models_refuse_cpu_batches():
    gd = load_golden("hybrid_v2")
    model = build_model("HybridModelv2", gd)
    from immunostruct_b200.graph import GraphBatch
    a = gd["graph"]
    g = GraphBatch.from_arrays(a["x"], a["src"], a["dst"], a["edge_attr"], a["node_counts"], a["edge_counts"])
    with pytest.raises(RuntimeError, match="CUDA"):
        model(g, gd["dense"]["seq"], gd["dense"]["prop"])


def test_graph_batch_host_vocabulary():
    gd = load_golden("hybrid_v2")
    a = gd["graph"]
    from immunostruct_b200.synthetic import split_graphs
    graphs = []
    for s in split_graphs(a):
        g = I.graph((s["src"], s["dst"]), num_nodes=s["num_nodes"])
        g.ndata["x"], g.edata["edge_attr"] = s["x"], s["edge_attr"]
        graphs.append(g)
    gb = I.batch(graphs)
    ref = R.dgl_batch(split_graphs(a))
    assert torch.equal(gb.edge_index, torch.stack([ref["src"], ref["dst"]]))
    assert torch.equal(gb.batch, R.batch_vector(ref["batch_num_nodes"]))
    assert torch.equal(gb.batch_num_nodes(), ref["batch_num_nodes"])
    assert torch.equal(gb.ndata["x"], ref["x"]) and torch.equal(gb.edata["edge_attr"], ref["edge_attr"])
    assert gb.num_nodes() == ref["num_nodes"] and gb.num_edges() == ref["src"].numel()
    pair = I.collate([((graphs[0], graphs[1]), (torch.zeros(2), torch.ones(2)), torch.tensor(1.0),
                       (torch.zeros(2), torch.ones(2)))] * 2)
    assert pair[0][0].n_graphs == 2 and pair[1][1].shape == (2, 2)


def test_pack_graph_batch_and_sequence_round_trip():
    """Compact format (immunostruct_b200/packed.py): packing is lossless for one-hot / zero-padded inputs (checked
    through the CPU contracts of the unpack kernels), refuses anything else, and is ~4x smaller than the dense form."""
    from immunostruct_b200.graph import GraphBatch
    from immunostruct_b200.packed import PAD_RESIDUE, pack_graph_batch, pack_sequence
    from immunostruct_b200.synthetic import synthetic_dense, synthetic_graph_arrays
    arr = synthetic_graph_arrays(3, 40, 6, seed=9, n_pad=5)
    keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")
    gb = GraphBatch.from_arrays(*(arr[k] for k in keys), max_nodes=40)
    pk = pack_graph_batch(gb)
    assert pk.edge_attr is None and int((pk.aa == PAD_RESIDUE).sum()) == 15 and pk.aa.dtype == torch.uint8
    x = torch.empty(gb.n_nodes, 23)
    KC.unpack_nodes(pk.aa, pk.xyz, x)
    assert torch.equal(x, arr["x"])
    s64, d64, ea = torch.empty(gb.n_edges, dtype=torch.int64), torch.empty(gb.n_edges, dtype=torch.int64), torch.empty(gb.n_edges)
    KC.unpack_edges(pk.src, pk.dst, pk.edge_attr, s64, d64, ea)
    assert torch.equal(s64, arr["src"].long()) and torch.equal(d64, arr["dst"].long()) and torch.equal(ea, arr["edge_attr"].reshape(-1))
    dense_bytes = sum(arr[k].numel() * arr[k].element_size() for k in keys)
    assert pk.nbytes * 3 < dense_bytes
    seq = synthetic_dense(3, seed=9)["seq"]
    ps = pack_sequence(seq)
    out = torch.empty_like(seq)
    KC.onehot_tokens(ps.tokens, out, ps.vocab)
    assert torch.equal(out, seq) and ps.nbytes * 80 < seq.numel() * 4
    bad = arr["x"].clone(); bad[0, :20] = 1.0                      # the SSL "mask to one" row is not packable
    with pytest.raises(ValueError):
        pack_graph_batch(GraphBatch.from_arrays(bad, *(arr[k] for k in keys[1:]), max_nodes=40))
    with pytest.raises(RuntimeError):
        pk.expand()                                                 # expansion is a device operation: no CPU path


def test_packed_dataset_batches_equal_reference_collate():
    """PackedGraphDataset.batch (vectorised gather over flat arrays) == packing the reference-style collate of the same
    samples, for contiguous, shuffled and repeated index sets; the loader keeps the last partial batch."""
    import immunostruct_b200 as I
    from immunostruct_b200.packed import PackedGraphDataset, pack_graph_batch
    from immunostruct_b200.synthetic import split_graphs, synthetic_dense, synthetic_graph_arrays
    arr = synthetic_graph_arrays(7, 30, 5, seed=21, n_pad=3)
    dense = synthetic_dense(7, seed=21)
    samples = []
    for g in split_graphs(arr):
        gr = I.graph((g["src"], g["dst"]), num_nodes=g["x"].shape[0])
        gr.ndata["x"], gr.edata["edge_attr"] = g["x"], g["edge_attr"]
        samples.append(gr)
    ds = PackedGraphDataset.from_samples(samples, dense["seq"], dense["target"], dense["prop"])
    assert len(ds) == 7
    for idx in ([0, 1, 2], [5, 2, 6, 0], [3, 3]):
        pk, ps, tgt, prop = ds.batch(idx)
        ref = pack_graph_batch(I.batch([samples[i] for i in idx]))
        for a, b in ((pk.aa, ref.aa), (pk.xyz, ref.xyz), (pk.src, ref.src), (pk.dst, ref.dst),
                     (pk.node_counts, ref.node_counts), (pk.edge_counts, ref.edge_counts)):
            assert torch.equal(a, b)
        assert pk.edge_attr is None and pk.max_nodes == 30
        assert torch.equal(tgt, dense["target"][idx]) and torch.equal(prop, dense["prop"][idx])
        assert torch.equal(ps.tokens, dense["seq"][idx].argmax(2).to(torch.uint8))
    sizes = [b[0].node_counts.numel() for b in ds.loader(3)]
    assert sizes == [3, 3, 1] and [b[0].node_counts.numel() for b in ds.loader(3, drop_last=True)] == [3, 3]
    g = torch.Generator().manual_seed(0)
    assert sorted(int(t) for b in ds.loader(2, shuffle=True, generator=g) for t in b[2].tolist()) == sorted(int(t) for t in dense["target"].tolist())
