"""Time the node-side backward kernels (SIMT vs tensor-core) at the benchmark shape."""
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
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
h, hn, gh_out = r(n, 64), r(n, 64), r(n, 64)
W5, b5, W6, W1 = r(64, 128), r(64), r(64, 64), r(64, 130)
grid = _C.egnn_node_grid(n)
ghd, ghn, pp = torch.empty(n, 64, device=dev), torch.empty(n, 64, device=dev), torch.empty(grid, 64 * 128 + 64 + 4096 + 64, device=dev)
gz1, gQ, gD, gxd, gxo = r(e, 64), r(n, 64), r(e, 3), r(n, 3), r(n, 3)
gh, gx, pq = torch.empty(n, 64, device=dev), torch.empty(n, 3, device=dev), torch.empty(grid, 2 * 64 * 64 + 64, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
fns = {"node_post_bwd simt": lambda: _C.egnn_node_post_bwd(gh_out, h, hn, W5, b5, W6, ghd, ghn, pp),
       "node_post_bwd tc": lambda: _C.egnn_node_post_bwd_tc(gh_out, h, hn, W5, b5, W6, ghd, ghn, pp),
       "node_pre_bwd simt": lambda: _C.egnn_node_pre_bwd(gz1, gQ, gD, gxd, gxo, ghd, gb, h, W1, gh, gx, pq)}
if hasattr(_C, "egnn_node_pre_bwd_tc"):
    fns["node_pre_bwd tc"] = lambda: _C.egnn_node_pre_bwd_tc(gz1, gQ, gD, gxd, gxo, ghd, gb, h, W1, gh, gx, pq)
for name, fn in fns.items():
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name}: {sum(ts) / len(ts):.1f} us")
