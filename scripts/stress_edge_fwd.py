"""Stress run of the warp-specialised edge forward kernel: every variant (precision x coordinate branch x SiLU form x A-operand
variant) over several batch shapes and in-degrees, 30 launches each back to back; every launch must reproduce the first one bit
for bit and agree with the lock-step first-generation kernel.  Exits non-zero on any CUDA error or mismatch."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_graph_arrays

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
W1, b1, W2, b2, W3, b3, w4 = r(64, 130), r(64), r(64, 64), r(64), r(64, 64), r(64), r(1, 64)
n_fail = 0
for n_graphs, n_nodes, k in ((1, 200, 10), (7, 37, 3), (64, 200, 10), (300, 190, 12), (512, 200, 10), (40, 120, 30)):
    arr = synthetic_graph_arrays(n_graphs, n_nodes, k, seed=n_graphs, device=dev)
    gb = GraphBatch.from_arrays(*(arr[kk] for kk in ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")), max_nodes=n_nodes)
    n = gb.n_nodes
    PQ, x = r(n, 128), arr["x"][:, 20:]
    for prec, tol in ((_C.PREC_BF16X3, 2e-5), (_C.PREC_FP16X2, 2e-5), (_C.PREC_BF16, 3e-2)):
        for upd in (True, False):
            ref_hn, ref_x = torch.zeros(n, 64, device=dev), torch.zeros(n, 3, device=dev)
            _C.egnn_edge_fwd_tc(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, upd,
                                (_C.PREC_BF16X3 if prec == _C.PREC_FP16X2 else prec) | 16, ref_hn, ref_x, fast_act=False)
            for fast in (True, False):
                for var in (0, 1):
                    _C.set_ws_variant(var)
                    first = None
                    for it in range(30):
                        hn, xo = torch.full((n, 64), float("nan"), device=dev), torch.full((n, 3), float("nan"), device=dev)
                        _C.egnn_edge_fwd_tc(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, upd, prec, hn, xo, fast_act=fast)
                        if first is None:
                            first = (hn, xo)
                        elif not (torch.equal(hn, first[0]) and (not upd or torch.equal(xo, first[1]))):
                            print("NOT REPRODUCIBLE", n_graphs, prec, upd, fast, var, it); n_fail += 1; break
                    torch.cuda.synchronize()
                    e_h = float((first[0] - ref_hn).abs().max() / ref_hn.abs().max())
                    e_x = float(((first[1] - x) - (ref_x - x)).abs().max() / (ref_x - x).abs().max().clamp_min(1e-30)) if upd else 0.0
                    ok = e_h < tol and e_x < tol and bool(torch.isfinite(first[0]).all())
                    n_fail += 0 if ok else 1
                    print(f"graphs {n_graphs:4d} n {n_nodes} k {k:2d} prec {prec} coords {int(upd)} fast {int(fast)} var {var}: hn {e_h:.1e} x {e_x:.1e} {'ok' if ok else 'MISMATCH'}", flush=True)
_C.set_ws_variant(1)
print("failures:", n_fail)
sys.exit(1 if n_fail else 0)
