"""Round-2 kernels through the C ABI against their CPU contracts (oracle/kernel_contracts.py) on the same seeded
inputs: segment pooling, block-level collation (large / empty graphs, statistics), the paired contrastive loss
forward + backward, the fused Adam step, the augmentation kernels, default batch validation."""
import copy

import numpy as np
import pytest
import torch

import immunostruct_b200 as I
import importlib

from immunostruct_b200 import _C, augment
from immunostruct_b200 import functional as IF
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.optim import FusedAdam, FusedAdamW
from immunostruct_b200.synthetic import split_graphs, synthetic_graph_arrays
from oracle import kernel_contracts as KC
from oracle import reference_ops as R

from test_kernels_gpu import close, random_multigraph_arrays, to_dev

G = importlib.import_module("immunostruct_b200.graph")      # the package attribute `graph` is the dgl.graph() stand-in

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ---- segment pooling -----------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["mean", "max", "sum"])
@pytest.mark.parametrize("cols", [64, 192, 300])
def test_segment_pool_matches_contract(mode, cols):
    gen = torch.Generator().manual_seed(cols)
    counts = [17, 0, 200, 1, 64, 0, 33]
    off = torch.zeros(len(counts) + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor(counts), 0)
    x = torch.randn(int(off[-1]), cols, generator=gen)
    x[20:60] = x[20]                                  # identical rows inside graph 2: ties for the max
    x[17 + 190:17 + 200] = 0.0                        # padded rows
    g_out = torch.randn(len(counts), cols, generator=gen)
    out = torch.empty(len(counts), cols, device=DEV)
    _C.segment_pool_fwd(x.to(DEV), off.to(DEV), mode, out)
    ref = torch.empty(len(counts), cols)
    KC.segment_pool_fwd(x, off, mode, ref)
    close(out, ref, 1e-6, f"segment_pool {mode} fwd")
    gx = torch.full((x.shape[0], cols), float("nan"), device=DEV)
    _C.segment_pool_bwd(x.to(DEV), off.to(DEV), mode, out, g_out.to(DEV), gx)
    gref = torch.empty_like(x)
    KC.segment_pool_bwd(x, off, mode, ref, g_out, gref)
    close(gx, gref, 1e-6, f"segment_pool {mode} bwd")


def test_segment_pool_autograd_and_strided_input():
    gen = torch.Generator().manual_seed(1)
    arr = synthetic_graph_arrays(4, 50, 5, seed=2, n_pad=3)
    gb = to_dev(arr)
    wide = torch.randn(gb.n_nodes, 128, generator=gen).to(DEV).requires_grad_(True)
    pooled = IF.segment_pool(gb, wide[:, 32:96], "max")          # row stride 128, 64 columns
    pooled.pow(2).sum().backward()
    w2 = wide.detach().cpu().requires_grad_(True)
    ref = w2[:, 32:96].view(4, 50, 64).amax(1)
    ref.pow(2).sum().backward()
    close(pooled, ref, 1e-6, "segment max strided")
    close(wide.grad, w2.grad, 1e-6, "segment max strided grad")


# ---- collation: block-level stable counting sort ----------------------------------------------------
@pytest.mark.parametrize("name", ["large_and_empty", "batch512"])
def test_collate_block_kernel_bit_exact(name):
    if name == "large_and_empty":
        arrays = random_multigraph_arrays(9, [600, 0, 513, 3, 0, 512, 40, 1100], 6.5)     # > 512 nodes: single-warp path
    else:
        arrays = synthetic_graph_arrays(512, 200, 10, seed=1, n_pad=10)
    G.set_validation("off")
    try:
        gb = to_dev(arrays)
    finally:
        G.set_validation("deferred")
    ref = R.dgl_batch(split_graphs(arrays))
    csr = R.csr_from_coo(ref["src"], ref["dst"], ref["num_nodes"])
    assert torch.equal(gb.edge_index.cpu(), torch.stack([ref["src"], ref["dst"]]))
    assert torch.equal(gb.batch.cpu(), R.batch_vector(ref["batch_num_nodes"]))
    for k in ("indptr", "csr_src", "csr_dst", "csr_eid", "outptr", "csc_pos"):
        assert torch.equal(getattr(gb, k).cpu(), csr[k].to(torch.int32)), k
    deg = torch.bincount(ref["dst"], minlength=ref["num_nodes"])
    nc = arrays["node_counts"]
    assert gb.stats.cpu().tolist() == [int(deg.max()), 0, int(nc.max()), int((nc == 0).sum())]


def test_bad_batches_raise_by_default():
    arr = synthetic_graph_arrays(2, 140, 3, seed=5)
    G._validation["first_done"] = True                 # behave like "not the first batch of the process"
    G._validation["pending"].clear()
    # a node with 129 in-edges: unsupported by the edge tiles
    hub = {k: v.clone() for k, v in arr.items()}
    hub["dst"][:129] = 7
    hub["src"][:129] = (torch.arange(129) + 8) % 140
    to_dev(hub)                                        # deferred: nothing raised yet
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="in-degree"):
        to_dev(arr)                                    # the next collation reports the previous batch
    G._validation["pending"].clear()
    bad = {k: v.clone() for k, v in arr.items()}
    bad["src"][5] = 10_000
    with pytest.raises(ValueError, match="outside"):
        to_dev(bad).validate()
    G._validation["pending"].clear()
    # an understated max_nodes is caught (the attention kernels size their tiles by it)
    gb = GraphBatch.from_arrays(*(arr[k].to(DEV) for k in ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")),
                                max_nodes=100)
    with pytest.raises(ValueError, match="max_nodes"):
        gb.validate()
    G._validation["pending"].clear()
    G.set_validation("sync")
    try:
        with pytest.raises(ValueError, match="in-degree"):
            to_dev(hub)
    finally:
        G.set_validation("deferred")
        G._validation["pending"].clear()


# ---- paired contrastive loss -----------------------------------------------------------------------
def _ctr_inputs(b, d=104, z=128, seed=0, two_classes=True):
    gen = torch.Generator().manual_seed(seed)
    ec, ew = torch.randn(b, d, generator=gen) * 1.5, torch.randn(b, d, generator=gen) * 1.5
    ew = 0.6 * ec + 0.4 * ew                                           # correlated pairs
    t = (torch.rand(b, generator=gen) < 0.3).float() if two_classes else torch.rand(b, generator=gen)
    if two_classes:
        t[0], t[1] = 0.0, 1.0
    w1 = torch.randn(z, d, generator=gen) / d ** 0.5
    w2 = torch.randn(z, z, generator=gen) / z ** 0.5
    gamma, beta = 1 + 0.2 * torch.randn(z, generator=gen), 0.1 * torch.randn(z, generator=gen)
    return ec, ew, t, w1, gamma, beta, w2


@pytest.mark.parametrize("b,d,z", [(12, 104, 128), (256, 104, 128), (37, 50, 96), (512, 104, 128)])
def test_contrastive_kernels_match_contract(b, d, z):
    ec, ew, t, w1, gamma, beta, w2 = _ctr_inputs(b, d, z, seed=b)
    rm, rv, nt = torch.zeros(z), torch.ones(z), torch.zeros((), dtype=torch.int64)
    rm_d, rv_d, nt_d = rm.to(DEV), rv.to(DEV), nt.to(DEV)
    dv = [x.to(DEV) for x in (ec, ew, t, w1, gamma, beta, w2)]
    scratch = torch.empty(_C.contrastive_scratch_floats(b, z), device=DEV)
    out = torch.empty(4, device=DEV)
    _C.contrastive_fwd(*dv, 1e-5, 0.1, rm_d, rv_d, nt_d, 1e-2, scratch, out)
    s_ref, o_ref = torch.empty(KC.contrastive_scratch_floats(b, z)), torch.empty(4)
    KC.contrastive_fwd(ec, ew, t, w1, gamma, beta, w2, 1e-5, 0.1, rm, rv, nt, 1e-2, s_ref, o_ref)
    close(out[:1], o_ref[:1], 2e-5, f"contrastive loss b={b}")
    assert abs(float(out[1] + out[2] + out[3]) - float(out[0])) <= 1e-5 * abs(float(out[0]))
    close(rm_d, rm, 1e-5, "running_mean"); close(rv_d, rv, 1e-5, "running_var")
    assert int(nt_d) == int(nt) == 2
    gout = torch.tensor([0.37])
    outs = [torch.full(s, float("nan"), device=DEV) for s in ((b, d), (b, d), (z, d), (z,), (z,), (z, z))]
    work = torch.empty(4 * b * z, device=DEV)
    _C.contrastive_bwd(dv[0], dv[1], dv[3], dv[4], dv[5], dv[6], scratch, gout.to(DEV), work, *outs)
    refs = [torch.empty(s) for s in ((b, d), (b, d), (z, d), (z,), (z,), (z, z))]
    KC.contrastive_bwd(ec, ew, w1, gamma, beta, w2, s_ref, gout, None, *refs)
    for name, g, r in zip(("gEc", "gEw", "gW1", "g_gamma", "g_beta", "gW2"), outs, refs):
        close(g, r, 5e-5, f"contrastive {name} b={b}")


def test_contrastive_module_gate_buffers_and_determinism():
    ec, ew, t, *_ = _ctr_inputs(64, seed=3)
    torch.manual_seed(0)
    pcl = I.PairedContrastiveLoss(embedding_dim=104, device=DEV)
    e1, e2 = ec.to(DEV).requires_grad_(True), ew.to(DEV).requires_grad_(True)
    before = copy.deepcopy(pcl.state_dict())
    for target in (torch.ones(64), torch.linspace(0, 1, 64)):           # one class / continuous: gate closed
        loss = pcl(e1, e2, target.to(DEV))
        assert float(loss) == 0.0
        loss.backward()
        assert float(e1.grad.abs().max()) == 0.0 and float(pcl.projector[0].weight.grad.abs().max()) == 0.0
    for k, v in pcl.state_dict().items():
        assert torch.equal(v, before[k]), k                             # BatchNorm buffers untouched
    l1 = pcl(e1, e2, t.to(DEV)); l1.backward()
    g1 = e1.grad.clone()
    assert int(pcl.projector[1].num_batches_tracked) == 2
    pcl.load_state_dict(before); e1.grad = None
    l2 = pcl(e1, e2, t.to(DEV)); l2.backward()
    assert torch.equal(l1, l2) and torch.equal(g1, e1.grad)             # bit-reproducible
    assert float(pcl(e1[:1], e2[:1], t[:1].to(DEV))) == 0.0              # a single pair: nothing to contrast
    # against the reference-derived CPU restatement
    p = {k: v.cpu() for k, v in before.items()}
    ref = R.paired_contrastive(p, ec, ew, t)
    assert abs(float(l1) - float(ref)) <= 2e-5 * abs(float(ref))


# ---- fused Adam ------------------------------------------------------------------------------------
@pytest.mark.parametrize("capturable", [False, True])
@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (False, 1e-6), (True, 1e-6), (True, 1e-2)])
def test_fused_adam_matches_torch_on_the_model(decoupled, wd, capturable):
    torch.manual_seed(1)
    a = I.model_map["HybridModelv2"](vae_input_dim=231, device=DEV, vae_hidden_dim=32).to(DEV)
    b = copy.deepcopy(a)
    opt_t = (torch.optim.AdamW if decoupled else torch.optim.Adam)(a.parameters(), lr=1e-3, weight_decay=wd)
    opt_f = (FusedAdamW if decoupled else FusedAdam)(b.parameters(), lr=1e-3, weight_decay=wd, capturable=capturable)
    arr = synthetic_graph_arrays(3, 40, 5, seed=2, n_pad=2)
    gb = to_dev(arr)
    gen = torch.Generator().manual_seed(0)
    seq = torch.nn.functional.one_hot(torch.randint(0, 21, (3, 11), generator=gen), 21).float().to(DEV)
    prop, y = torch.rand(3, 2, generator=gen).to(DEV), torch.tensor([0., 1., 0.]).to(DEV)
    eps = torch.randn(3, 32, generator=gen).to(DEV)
    losses = I.Losses(231, [2.0, 1.0], sequence=True)
    b.eval()
    b.reparameterize = lambda mu, lv: mu + eps * torch.exp(0.5 * lv)
    for step in range(4):
        # the SAME gradients go to both optimisers (Adam normalises: two separately evolving models would diverge by
        # ~lr on parameters whose gradient is pure rounding noise, e.g. the key bias under softmax shift invariance)
        opt_f.zero_grad()
        r, mu, lv, out = b(gb, seq, prop)
        losses.BCE_loss(r, seq, mu, lv, out, y).backward()
        for pa, pb in zip(a.parameters(), b.parameters()):
            pa.grad = None if pb.grad is None else pb.grad.detach().clone()
        opt_t.step()
        opt_f.step()
    for (k, p), (_, q) in zip(b.named_parameters(), a.named_parameters()):
        close(p, q, 2e-6, f"fused adam {k}")
    assert b.GCN_layers[5].coord_mlp[0].weight.grad is None
    assert torch.equal(b.GCN_layers[5].coord_mlp[0].weight, a.GCN_layers[5].coord_mlp[0].weight)
    assert len(opt_f._flat[0]["runs"]) == 1                             # one launch covers every live parameter


# ---- augmentations ---------------------------------------------------------------------------------
def test_rotation_matches_numpy_qr():
    arr = synthetic_graph_arrays(6, 33, 4, seed=8, n_pad=3)
    gb = to_dev(arr)
    x0 = gb.ndata["x"].clone()
    gen = torch.Generator().manual_seed(4)
    m = torch.randn(6, 9, generator=gen)
    m[5] = torch.eye(3).reshape(-1)                          # zero sub-columns: the tau = 0 path of dlarfg
    q = torch.empty(6, 9, device=DEV)
    _C.rotate_coords(gb.ndata["x"], 20, gb.node_off, m.to(DEV), q)
    xr, qr = x0.cpu().clone(), torch.empty(6, 9)
    KC.rotate_coords(xr, 20, gb.node_off.cpu(), m, qr)
    close(q, qr, 2e-6, "rotation Q")
    close(gb.ndata["x"], xr, 2e-6, "rotated x")
    assert torch.equal(gb.ndata["x"][:, :20], x0[:, :20])
    # the library path draws M on the device and leaves distances unchanged
    augment.random_rotation_(gb, torch.Generator(device=DEV).manual_seed(1))
    mode = "donot_use_mm_for_euclid_dist"
    d0 = torch.cdist(x0[:33, 20:], x0[:33, 20:], compute_mode=mode)
    d1 = torch.cdist(gb.ndata["x"][:33, 20:], gb.ndata["x"][:33, 20:], compute_mode=mode)
    close(d1, d0, 1e-5, "pairwise distances after two rotations")


def test_masking_kernels_match_contracts_bit_exactly():
    arr = random_multigraph_arrays(3, [40, 7, 1, 200, 64], 3.0)
    arr["x"][45:47, :20] = 0.0                               # padded rows inside graph 1
    arr["x"][47, :20] = 0.0                                  # graph 2 (one node) has no valid residue
    gb = to_dev(arr)
    off = gb.node_off
    gen = torch.Generator().manual_seed(9)
    u = torch.rand(5, generator=gen); u[0] = 0.999999
    xd, xc = gb.ndata["x"].clone(), arr["x"].clone()
    aa_d, node_d = torch.empty(5, dtype=torch.int64, device=DEV), torch.empty(5, dtype=torch.int64, device=DEV)
    _C.mask_single_residue(xd, 20, off, u.to(DEV), None, aa_d, node_d)
    aa_c, node_c = torch.empty(5, dtype=torch.int64), torch.empty(5, dtype=torch.int64)
    KC.mask_single_residue(xc, 20, off.cpu(), u, None, aa_c, node_c)
    assert torch.equal(aa_d.cpu(), aa_c) and torch.equal(node_d.cpu(), node_c) and torch.equal(xd.cpu(), xc)
    assert int(node_c[2]) == -1 and int(aa_c[2]) == 0
    # partner graph: residue of the requested type
    xd2, xc2 = gb.ndata["x"].clone(), arr["x"].clone()
    want = aa_c.clone(); want[2] = -1
    u2 = torch.rand(5, generator=gen)
    _C.mask_single_residue(xd2, 20, off, u2.to(DEV), want.to(DEV), aa_d, node_d)
    KC.mask_single_residue(xc2, 20, off.cpu(), u2, want, aa_c, node_c)
    assert torch.equal(aa_d.cpu(), aa_c) and torch.equal(node_d.cpu(), node_c) and torch.equal(xd2.cpu(), xc2)
    # mask_structure on top of the SSL-masked rows (all-ones rows are skipped), then mask_sequence with limits
    keys = torch.rand(gb.n_nodes, generator=gen)
    _C.mask_rows(xd, 20, off, None, keys.to(DEV), 5, -1, 200)
    KC.mask_rows(xc, 20, off.cpu(), None, keys, 5, -1, 200)
    assert torch.equal(xd.cpu(), xc)
    seq = torch.nn.functional.one_hot(torch.randint(0, 20, (7, 283), generator=gen), 21).float()
    sd = seq.to(DEV).view(-1, 21).contiguous()
    sc = seq.view(-1, 21).clone()
    seg = torch.arange(8, dtype=torch.int64) * 283
    limit = torch.tensor([272, 272, 274, 0, 1, 283, 272], dtype=torch.int64)
    k2 = torch.rand(7 * 283, generator=gen)
    _C.mask_rows(sd, 21, seg.to(DEV), limit.to(DEV), k2.to(DEV), 4, 20, 283)
    KC.mask_rows(sc, 21, seg, limit, k2, 4, 20, 283)
    assert torch.equal(sd.cpu(), sc)
    assert int((sc.view(7, 283, 21).argmax(-1) == 20).sum()) == 4 * 5 + 0 + 1


def test_structure_model_v2_uses_segment_max_kernel():
    from conftest import assert_grads_close, load_golden, rel_err
    from helpers import graph_batch, named_grads
    gd = load_golden("structure_v2")
    model = I.model_map["StructureModelv2"](vae_input_dim=231, device=DEV, gcn_layers=1)
    model.load_state_dict(gd["weights"])
    model = model.to(DEV).eval()
    g = graph_batch(gd["graph"], DEV)
    before = _C.LAUNCHES
    z0, z1, z2, out, node_pred = model(g, gd["dense"]["seq"].to(DEV), gd["dense"]["prop"].to(DEV))
    assert rel_err(out, gd["out"]["logits"]) < 2e-5 and rel_err(node_pred, gd["out"]["node_pred"]) < 2e-5
    (out.sum() + node_pred.pow(2).sum()).backward()
    assert_grads_close(named_grads(model), gd["grads"], 1e-4)
    assert _C.LAUNCHES > before


# ---- TMA-fed tcgen05 GEMM (csrc/gemm_tma.cu) ----------------------------------------------------------
def test_split_planes_bit_exact_against_contract():
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(70, 45, generator=gen) * torch.logspace(-3, 3, 45)
    mask_src = torch.randn(70, 45, generator=gen)
    for n in (3, 1):
        for relu_src in (None, mask_src):
            got = _C.split_planes(x.to(DEV), n, rows=True, transposed=True, relu_src=None if relu_src is None else relu_src.to(DEV),
                                  colsum=True, flag=True)
            ref = KC.split_planes(x, n, rows=True, transposed=True, relu_src=relu_src, colsum=True, flag=True)
            assert got[0].shape == (n, 70, 48) and got[1].shape == (n, 45, 72)
            assert torch.equal(got[0].cpu().view(torch.int16), ref[0].view(torch.int16))
            assert torch.equal(got[1].cpu().view(torch.int16), ref[1].view(torch.int16))
            close(got[2].sum(0), ref[2].sum(0), 1e-6, "colsum")
            assert int(ref[3]) == 1 and int(got[3]) == (1 if n == 3 else 0)     # the flag is about planes 2 / 3
    onehot = torch.nn.functional.one_hot(torch.randint(0, 21, (64,), generator=gen), 21).float()
    assert int(_C.split_planes(onehot.to(DEV), 3, flag=True)[3]) == 0          # exact in bf16
    xs = x.to(DEV)
    p3 = _C.split_planes(xs, 3)[0].float().sum(0)[:, :45]
    close(p3, x, 1e-7, "three planes reconstruct fp32")


@pytest.mark.parametrize("m,n,k", [(512, 512, 5943), (512, 5943, 512), (5943, 512, 512), (130, 70, 100), (3, 32, 231),
                                   (3, 231, 32), (256, 300, 64), (1, 1, 8)])
def test_gemm_planes_tma_is_fp32_accurate(m, n, k):
    gen = torch.Generator().manual_seed(m + n + k)
    a, b = torch.randn(m, k, generator=gen), torch.randn(n, k, generator=gen) / k ** 0.5
    bias = torch.randn(n, generator=gen)
    ap = _C.split_planes(a.to(DEV), 3)[0]
    bp = _C.split_planes(b.to(DEV), 3)[0]
    ref = a.double() @ b.double().t() + bias.double()
    out = _C.gemm_planes(ap, bp, bias.to(DEV))
    tol = 1e-5 if k > 4096 or max(m, n) > 4096 else 2e-6          # fp32 accumulation over K: sqrt(K) 2^-24
    close(out, ref, tol, f"gemm_planes {m}x{n}x{k}")
    out_relu = _C.gemm_planes(ap, bp, bias.to(DEV), relu=True)
    close(out_relu, ref.clamp_min(0), tol, f"gemm_planes relu {m}x{n}x{k}")
    # strided output rows (a column slice of a wider buffer)
    wide = torch.full((m, n + 5), 7.0, device=DEV)
    _C.gemm_planes(ap, bp, None, out=wide[:, 2:2 + n])
    close(wide[:, 2:2 + n], ref - bias.double(), tol, "strided out")
    assert float(wide[:, :2].min()) == 7.0 and float(wide[:, 2 + n:].max()) == 7.0
    # single bf16 plane: the 1e-2 mode
    a1, b1 = _C.split_planes(a.to(DEV), 1)[0], _C.split_planes(b.to(DEV), 1)[0]
    close(_C.gemm_planes(a1, b1, bias.to(DEV)), ref, 1e-2, f"gemm_planes bf16 {m}x{n}x{k}")


def test_gemm_planes_exact_operand_flags():
    gen = torch.Generator().manual_seed(3)
    seq = torch.nn.functional.one_hot(torch.randint(0, 21, (96, 283), generator=gen), 21).float().reshape(96, -1)
    w = torch.randn(200, 5943, generator=gen) * 0.02
    sp, spt, _, flag = _C.split_planes(seq.to(DEV), 3, rows=True, transposed=True, flag=True)
    wp = _C.split_planes(w.to(DEV), 3)[0]
    assert int(flag) == 0
    ref = seq.double() @ w.double().t()
    close(_C.gemm_planes(sp, wp, a_flag=flag), ref, 2e-6, "A exact: three products")
    g = torch.randn(96, 200, generator=gen)
    gpt = _C.split_planes(g.to(DEV), 3, rows=False, transposed=True)[1]
    close(_C.gemm_planes(gpt, spt, b_flag=flag), g.double().t() @ seq.double(), 2e-6, "B exact: three products")


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("b,k,n", [(512, 5943, 512), (512, 512, 5943), (37, 231, 32)])
def test_linear_tc_autograd_matches_fp64(b, k, n, relu):
    gen = torch.Generator().manual_seed(b + k)
    x = torch.randn(b, k, generator=gen)
    w, bias = torch.randn(n, k, generator=gen) / k ** 0.5, torch.randn(n, generator=gen) * 0.1
    gy = torch.randn(b, n, generator=gen)
    xd = x.to(DEV).requires_grad_(True)
    wd, bd = torch.nn.Parameter(w.to(DEV)), torch.nn.Parameter(bias.to(DEV))      # Parameters: their planes are cached
    y = IF.linear_tc(xd, wd, bd, relu)
    y.backward(gy.to(DEV))
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, bias))
    y64 = torch.nn.functional.linear(x64, w64, b64)
    if relu:
        # the product's own activation pattern: among 3 M outputs a few dozen lie within rounding of zero, and a flipped
        # ReLU there changes whole gradient rows (a discontinuity of the function, not an error of the kernel)
        mask = (y.detach().cpu() > 0).double()
        assert float(((y64 > 0).double() - mask).abs().mean()) < 1e-4
        y64 = y64 * mask
    y64.backward(gy.double())
    close(y, y64, 1e-5, "linear_tc y")
    close(xd.grad, x64.grad, 1e-5, "linear_tc gx")
    close(wd.grad, w64.grad, 1e-5, "linear_tc gw")
    close(bd.grad, b64.grad, 1e-5, "linear_tc gb")
    # weights are re-split when (and only when) their version changes
    from immunostruct_b200.functional import _weight_planes
    key0 = _weight_planes[id(wd)][0]
    with torch.no_grad():
        IF.linear_tc(xd, wd, bd, relu)
        assert _weight_planes[id(wd)][0] == key0
        wd.mul_(2.0)
        y2 = IF.linear_tc(xd, wd, bd, False)
    assert _weight_planes[id(wd)][0] != key0
    close(y2, torch.nn.functional.linear(x64, 2 * w64, b64), 1e-5, "linear_tc after in-place update")


@pytest.mark.gpu
def test_captured_step_reproduces_eager_steps():
    """immunostruct_b200.CapturedStep (whole training step -- collation, forward, loss, backward, capturable FusedAdam -- in one
    CUDA graph) against the same steps run eagerly: same losses and same parameters after warm-up + 3 steps on new batches
    (dropout off and a fixed reparameterisation draw, so both runs are deterministic functions of the data)."""
    from immunostruct_b200.synthetic import synthetic_dense
    dev = "cuda"
    keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")
    batches = []
    for s in range(4):
        arr = synthetic_graph_arrays(16, 60, 6, seed=40 + s, device=dev)
        den = synthetic_dense(16, seed=40 + s, device=dev)
        batches.append({**{k: arr[k] for k in keys}, "seq": den["seq"], "prop": den["prop"], "target": den["target"]})
    eps = torch.randn(16, 32, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    losses = I.Losses(5943, [0.81, 0.19], sequence=True)

    def make():
        torch.manual_seed(11)
        model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev).eval()      # eval(): dropout off; autograd on
        model.sample_eps = lambda like: eps.to(like.dtype)
        opt = FusedAdam(model.parameters(), lr=1e-3, capturable=True)

        def body(t):
            gb = GraphBatch.from_arrays(*(t[k] for k in keys), max_nodes=60)
            opt.zero_grad(set_to_none=True)
            recon, mu, logvar, out = model(gb, t["seq"], t["prop"])
            loss = losses.BCE_loss(recon, t["seq"], mu, logvar, out, t["target"])
            loss.backward()
            opt.step()
            return loss
        return model, body

    model_e, body_e = make()
    for _ in range(3):
        body_e(batches[0])
    eager = [float(body_e(b)) for b in batches[1:]]
    model_c, body_c = make()
    step = I.CapturedStep(body_c, batches[0], warmup=3)
    assert step.warmup_steps == 3
    captured = [float(step(b)) for b in batches[1:]]
    for a, c in zip(eager, captured):
        assert abs(a - c) <= 1e-6 * max(1.0, abs(a)), (eager, captured)
    for (k, pe), (_, pc) in zip(model_e.named_parameters(), model_c.named_parameters()):
        assert torch.allclose(pe, pc, rtol=1e-6, atol=1e-8), k
