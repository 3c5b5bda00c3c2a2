"""A/B timing of the warp-specialised edge forward kernel: A operand of MMA 2 from shared memory (variant 0) or from
tensor memory (variant 1, default); batch 512, L2 flushed; the results must be bit-identical."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_graph_arrays

dev = "cuda"
arr = synthetic_graph_arrays(512, 200, 10, seed=1, device=dev)
gb = GraphBatch.from_arrays(*(arr[k] for k in ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")), max_nodes=200)
n = gb.n_nodes
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
W1, b1, W2, b2, W3, b3, w4 = r(64, 130), r(64), r(64, 64), r(64), r(64, 64), r(64), r(1, 64)
PQ = r(n, 128)
x = arr["x"][:, 20:]
flush = torch.empty(64 * 1024 * 1024, device=dev)
ref = {}
for prec, name in ((_C.PREC_BF16X3, "bf16x3"), (_C.PREC_FP16X2, "fp16x2"), (_C.PREC_BF16, "bf16")):
    for fast in (True, False):
        for upd in (True, False):
            for nb in (0, 1):
                _C.set_ws_variant(nb)
                hn, xo = torch.zeros(n, 64, device=dev), torch.zeros(n, 3, device=dev)
                ts = []
                for it in range(12):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _C.egnn_edge_fwd_tc(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, upd, prec, hn, xo, fast_act=fast)
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 4:
                        ts.append(e0.elapsed_time(e1) * 1e3)
                key = (name, fast, upd)
                if nb == 0:
                    ref[key] = (hn.clone(), xo.clone())
                    same = ""
                else:
                    same = f" identical to variant 0: hn {torch.equal(hn, ref[key][0])} x {torch.equal(xo, ref[key][1])}"
                print(f"{name} fast={fast} coords={upd} variant={nb}: {sum(ts) / len(ts):.1f} us (min {min(ts):.1f}){same}", flush=True)
_C.set_ws_variant(1)
