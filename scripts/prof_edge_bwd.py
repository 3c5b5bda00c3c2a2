"""Time the tensor-core edge backward kernels at the benchmark shape (batch 512): lock-step (egnn_bwd_tc.cu) against
the two-stream kernel (egnn_bwd_ws.cu), with and without the coordinate branch; checks that both agree."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_graph_arrays

dev = "cuda"
arr = synthetic_graph_arrays(512, 200, 10, seed=1, device=dev)
gb = GraphBatch.from_arrays(*(arr[k] for k in ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")), max_nodes=200)
n, e = gb.n_nodes, gb.n_edges
print("nodes", n, "edges", e, "stats", gb.stats.tolist())
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
PQ, ghn, gxo = r(n, 128), r(n, 64), r(n, 3)
x = arr["x"][:, 20:]
ea = arr["edge_attr"].float()
W1, W2, b2, W3, b3, w4 = r(64, 130), r(64, 64), r(64), r(64, 64), r(64), r(1, 64)
grid = _C.egnn_edge_bwd_grid(n)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def outs():
    return [torch.zeros(e, 64, device=dev), torch.zeros(n, 64, device=dev), torch.zeros(e, 3, device=dev),
            torch.zeros(n, 3, device=dev), torch.zeros(grid, 8512, device=dev)]


res = {}
for coord in (True, False):
    for name, fn in (("tc", _C.egnn_edge_bwd_tc), ("ws8", _C.egnn_edge_bwd_ws), ("ws12", _C.egnn_edge_bwd_ws), ("ws16", _C.egnn_edge_bwd_ws)):
        if name.startswith("ws"):
            _C.set_bwd_ws_warps(int(name[2:]))
        o = outs()
        call = lambda: fn(gb, PQ, x, ea, 64, W1, W2, b2, W3, b3, w4, ghn, gxo if coord else None, *o)
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        red = torch.empty(8512, device=dev)
        _C.reduce_partials(o[4], red)
        res[(coord, name)] = o[:4] + [red]
        print(f"edge_bwd {name} coord={coord}: {sum(ts) / len(ts):.1f} us  (min {min(ts):.1f})", flush=True)
    for w in ("ws8", "ws12", "ws16"):
        a, b = res[(coord, "tc")], res[(coord, w)]
        print(f"   coord={coord} {w}: max |tc - ws| / scale: " + "  ".join(
            f"{nm} {float((u - v).abs().max()) / (float(u.abs().max()) + 1e-30):.2e}" for nm, u, v in zip(("gz1", "gQ", "gD", "gxd", "partials"), a, b)))
print("status", int(gb.status.item()))
