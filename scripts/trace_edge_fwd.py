"""Timeline of the warp-specialised edge forward kernel (debug build, -DIS_TRACE): clock64() stamps of CTA 0 per role and
tile.  Run:  bash scripts/build_debug_lib.sh && IS_B200_DEBUG_LIB=immunostruct_b200/build/libimmunostruct_b200_debug.so python scripts/trace_edge_fwd.py"""
import ctypes
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
hn, xo = torch.zeros(n, 64, device=dev), torch.zeros(n, 3, device=dev)
NEV = 18
trace = torch.zeros(NEV * 128, dtype=torch.int64, device=dev)
fn = _C.lib().is_debug_set_trace
fn.restype = ctypes.c_int
assert fn(ctypes.c_void_p(trace.data_ptr())) == 0
prec = _C.PREC_BF16 if "bf16" in sys.argv[1:] else _C.PREC_BF16X3
for it in range(3):
    trace.zero_()
    _C.egnn_edge_fwd_tc(gb, PQ, x, arr["edge_attr"], 64, W1, W2, b2, W3, b3, w4, True, prec, hn, xo, fast_act=True)
    torch.cuda.synchronize()
t = trace.view(NEV, 128).cpu()
ntiles = int((t[9] > 0).sum())
print("tiles of CTA 0:", ntiles, " kernel span (cycles):", int(t.max()) - int(t[t > 0].min()))
print("tile period (gather done -> gather done), tiles 10..30:", [int(t[9, i + 1] - t[9, i]) for i in range(10, min(30, ntiles - 1))])
lo, hi = 8, max(9, ntiles - 4)
def mean(a, b, sa=0, sb=0):
    d = [int(t[b, i + sb] - t[a, i + sa]) for i in range(lo, hi) if t[a, i + sa] > 0 and t[b, i + sb] > 0]
    return sum(d) / max(len(d), 1)
rows = [("scalars: wait for a free slot", 0, 1), ("scalars: CSR walk + geometry of a tile", 1, 2),
        ("gather: first loads issued, then wait for the operand buffer", 7, 8), ("gather: selector + round 0 (loads land, silu, split, store)", 8, 17),
        ("gather: round 1 (load, silu, split, store)", 17, 9), ("a_full arrive -> MMA warp 1 sees it", 9, 3), ("MMA 1 issue (24 MMAs)", 3, 4),
        ("MMA 1 issued -> epilogue sees acc1 full", 4, 11), ("epilogue: wait for acc1 (idle)", 10, 11), ("epilogue 1 (silu, split, smem + TMEM stores)", 11, 12),
        ("m_full arrive -> MMA warp 2 sees it", 12, 5), ("MMA 2 + MMA 3 issue (48 MMAs)", 5, 6), ("MMA 2 issued -> acc2 full seen", 6, 13),
        ("epilogue 2 (c = w4 . silu)", 13, 14), ("epilogue: wait for hn (idle)", 14, 15), ("hn rows + coordinates", 15, 16)]
for name, a, b in rows:
    print(f"{mean(a, b):8.0f} cycles   {name}")
print(f"{mean(7, 7, 0, 1):8.0f} cycles   gather period;  {mean(10, 10, 0, 1):8.0f} epilogue period")
i = 20
t0 = int(t[7, i])
print("tile 20, stamps relative to its gather start:", {k: int(t[k, i]) - t0 for k in range(NEV) if t[k, i] > 0})
