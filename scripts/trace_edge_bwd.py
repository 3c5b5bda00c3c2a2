"""Timeline of the two-stream edge backward kernel (debug build, -DIS_TRACE): clock64() stamps of both teams of CTA 0 per
phase and tile.  Run:  bash scripts/build_debug_lib.sh &&
IS_B200_DEBUG_LIB=immunostruct_b200/build/libimmunostruct_b200_debug.so python scripts/trace_edge_bwd.py [warps]"""
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
n, e = gb.n_nodes, gb.n_edges
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
PQ, ghn, gxo = r(n, 128), r(n, 64), r(n, 3)
x = arr["x"][:, 20:]
ea = arr["edge_attr"].float()
W1, W2, b2, W3, b3, w4 = r(64, 130), r(64, 64), r(64), r(64, 64), r(64), r(1, 64)
grid = _C.egnn_edge_bwd_grid(n)
o = [torch.zeros(e, 64, device=dev), torch.zeros(n, 64, device=dev), torch.zeros(e, 3, device=dev),
     torch.zeros(n, 3, device=dev), torch.zeros(grid, 8512, device=dev)]
NEV = 20
trace = torch.zeros(2 * NEV * 64, dtype=torch.int64, device=dev)
fn = _C.lib().is_debug_set_trace_bwd
fn.restype = ctypes.c_int
assert fn(ctypes.c_void_p(trace.data_ptr())) == 0
if len(sys.argv) > 1:
    _C.set_bwd_ws_warps(int(sys.argv[1]))
for it in range(3):
    trace.zero_()
    _C.egnn_edge_bwd_ws(gb, PQ, x, ea, 64, W1, W2, b2, W3, b3, w4, ghn, gxo, *o)
    torch.cuda.synchronize()
t = trace.view(2, NEV, 64).cpu()
names = ["start", "meta", "gather", "mma1 issued", "mma1 done", "epi1", "mma2 done", "epi2", "mma3 done", "epi3 compute",
         "wg3 done", "epi3 stores", "mma4 done", "epi4 compute", "wg2 done", "f32 copy", "geometry", "dst sums"]
for team in range(2):
    nt = int((t[team, 17] > 0).sum())
    print(f"team {team}: {nt} tiles; span {int(t[team].max()) - int(t[team][t[team] > 0].min())} cycles; "
          f"period {[(int(t[team, 0, i + 1] - t[team, 0, i])) for i in range(4, min(14, nt - 1))]}")
    lo, hi = 3, nt - 2
    prev = 0
    tot = 0.0
    for ev in range(1, 18):
        d = [int(t[team, ev, i] - t[team, ev - 1, i]) for i in range(lo, hi)]
        m = sum(d) / max(len(d), 1)
        tot += m
        print(f"   {names[ev - 1]:>14} -> {names[ev]:<14} {m:8.0f} cycles")
    print(f"   sum {tot:.0f}")
    # the LAST publish of a tile (MMA 4): thread 0's stores done -> its fences done -> team barrier passed -> MMAs issued + data ready
    d1 = [int(t[team, 18, i] - t[team, 11, i]) for i in range(lo, hi)]
    d2 = [int(t[team, 19, i] - t[team, 18, i]) for i in range(lo, hi)]
    d3 = [int(t[team, 12, i] - t[team, 19, i]) for i in range(lo, hi)]
    print(f"   publish of MMA 4: fences {sum(d1) / len(d1):.0f} | team barrier (slowest warp) {sum(d2) / len(d2):.0f} | issue + MMA 4 + wait {sum(d3) / len(d3):.0f}")
# interleaving of the two teams: start stamps relative to team 0's tile 4
b0 = int(t[0, 0, 0])
print("team 0 starts:", [int(t[0, 0, i]) - b0 for i in range(0, 12)])
print("team 1 starts:", [int(t[1, 0, i]) - b0 for i in range(0, 12)])
for i in range(0, 3):
    print(f"tile {i}: team 0", [int(t[0, ev, i]) - b0 for ev in range(18)])
    print(f"tile {i}: team 1", [int(t[1, ev, i]) - b0 for ev in range(18)])
