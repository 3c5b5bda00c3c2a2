"""Time the pooled-only attention backward (SIMT vs tensor-core) at the benchmark shape (512 graphs x 200 nodes)."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C

dev = "cuda"
b, m = 512, 200
g = torch.Generator(device=dev).manual_seed(0)
QKV = torch.randn(b * m, 192, device=dev, generator=g) * 0.7
gp = torch.randn(b, 64, device=dev, generator=g)
node_off = torch.arange(b + 1, device=dev, dtype=torch.int64) * m
out_s, out_t = torch.empty_like(QKV), torch.empty_like(QKV)
lse = torch.empty(b * m, 1, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    z.record()
    torch.cuda.synchronize()
    return a.elapsed_time(z) / reps * 1e3


print(f"attn_bwd simt: {timed(lambda: _C.attn_pool_bwd(QKV, None, lse, node_off, 1, m, gp, None, out_s)):.1f} us")
print(f"attn_bwd tc: {timed(lambda: _C.attn_pool_bwd_tc(QKV, node_off, m, gp, out_t)):.1f} us")
d = (out_s - out_t).abs().max().item()
print("max |simt - tc| =", d, "scale", out_s.abs().max().item())
# distance of both kernels from the fp64 gradient (first 8 graphs)
worst = {"simt": 0.0, "tc": 0.0}
for gi in range(8):
    x = QKV[gi * m:(gi + 1) * m].double().cpu()
    Q, K, V = x[:, :64], x[:, 64:128], x[:, 128:]
    g0 = gp[gi].double().cpu() / m
    P = torch.softmax(Q @ K.T / 8.0, dim=1)
    c = V @ g0
    G = P * (c[None, :] - (P @ c)[:, None])
    ref = torch.cat([G @ K / 8.0, G.T @ Q / 8.0, P.sum(0)[:, None] * g0[None, :]], dim=1)
    for name, out in (("simt", out_s), ("tc", out_t)):
        for sec in range(3):
            r = ref[:, 64 * sec:64 * sec + 64]
            e = (out[gi * m:(gi + 1) * m, 64 * sec:64 * sec + 64].double().cpu() - r).abs().max() / r.abs().max()
            worst[name] = max(worst[name], float(e))
print("worst relative distance from fp64 (per Q/K/V block):", worst)
