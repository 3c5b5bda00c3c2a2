"""Time the fused node kernel alone for several node counts (fixed cost vs per-tile cost)."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
W5, b5, W6, b6, W1n, b1n = r(64, 128), r(64), r(64, 64), r(64), r(64, 130), r(64)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for prec in (_C.PREC_FP16X2, _C.PREC_BF16X3, _C.PREC_BF16):
    for tiles_per_sm in (1, 2, 4, 5.4, 8, 16):
        n = int(128 * 148 * tiles_per_sm) if tiles_per_sm != 5.4 else 102400
        h, hn = r(n, 64), r(n, 64)
        ho, PQ = torch.empty(n, 64, device=dev), torch.empty(n, 128, device=dev)
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _C.egnn_node_post_pre_tc(h, hn, W5, b5, W6, b6, ho, W1n, b1n, PQ, prec, fast_act=True)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"prec {prec} tiles/SM {tiles_per_sm}: n={n} {sum(ts) / len(ts):.1f} us")
