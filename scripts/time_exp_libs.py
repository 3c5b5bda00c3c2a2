"""Times the warp-specialised edge forward kernel of every experiment build (scripts/build_exp_libs.sh) in its own process:
python scripts/time_exp_libs.py egnn_tc2 0 1 2 ...   (batch 512, L2 flushed, bf16x3 inference + training SiLU + bf16)"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] != "--child":
    f = sys.argv[1]
    for n in sys.argv[2:]:
        env = dict(os.environ, IS_B200_DEBUG_LIB=f"immunostruct_b200/build/exp/lib_{f}_{n}.so")
        r = subprocess.run([sys.executable, __file__, "--child", n], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-2000:], flush=True)
    sys.exit(0)

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
out = []
for prec, name, fast, upd in ((_C.PREC_BF16X3, "bf16x3", True, True), (_C.PREC_BF16X3, "bf16x3-train", False, True),
                              (_C.PREC_BF16X3, "bf16x3-nocoord", True, False), (_C.PREC_BF16, "bf16", True, True)):
    hn, xo = torch.zeros(n, 64, device=dev), torch.zeros(n, 3, device=dev)
    ts = []
    for it in range(14):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.egnn_edge_fwd_tc(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, upd, prec, hn, xo, fast_act=fast)
        e1.record()
        torch.cuda.synchronize()
        if it >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3)
    out.append(f"{name} {sum(ts) / len(ts):.1f} (min {min(ts):.1f}) chk {float(hn.double().abs().sum()):.6e} {float(xo.double().abs().sum()):.6e}")
print(f"exp {sys.argv[2]}: " + " | ".join(out))
