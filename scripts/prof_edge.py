"""Run the EGNN edge kernels in isolation on a 512-graph batch (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_graph_arrays

prec = {"bf16": _C.PREC_BF16, "tf32x3": _C.PREC_TF32X3, "bf16x3": _C.PREC_BF16X3, "fp16x2": _C.PREC_FP16X2, "fp32": None}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
dev = "cuda"
arr = synthetic_graph_arrays(512, 200, 10, seed=1, device=dev)
gb = GraphBatch.from_arrays(*(arr[k] for k in ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")), max_nodes=200)
n = gb.n_nodes
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
W1, b1, W2, b2, W3, b3, w4 = r(64, 130), r(64), r(64, 64), r(64), r(64, 64), r(64), r(1, 64)
PQ, hn, xo = r(n, 128), torch.empty(n, 64, device=dev), torch.empty(n, 3, device=dev)
x = arr["x"][:, 20:]
for _ in range(4):
    if prec is None:
        _C.egnn_edge_fwd(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, True, hn, xo)
    else:
        _C.egnn_edge_fwd_tc(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, True, prec, hn, xo, fast_act=True)
torch.cuda.synchronize()
print("done", float(hn.abs().mean()))
