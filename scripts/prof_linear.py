"""Time the tcgen05 Linear kernel on the VAE shapes against torch / cuBLAS fp32."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g) * 0.2
flush = torch.empty(64 * 1024 * 1024, device=dev)
for m, n, k in ((512, 512, 5943), (512, 5943, 512), (512, 32, 512), (512, 512, 40), (4096, 512, 5943), (4096, 5943, 512)):
    x, w, b = r(m, k), r(n, k), r(n)
    for name, fn in (("tc bf16x3", lambda: _C.linear_tc(x, w, b, relu=True)), ("tc bf16", lambda: _C.linear_tc(x, w, b, relu=True, precision=_C.PREC_BF16)),
                     ("cublas fp32", lambda: torch.relu(torch.nn.functional.linear(x, w, b)))):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{m}x{n}x{k} {name}: {sum(ts) / len(ts):.1f} us")
