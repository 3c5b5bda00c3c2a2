"""Per-parameter distance of the product's gradients from the fp64 run of the reference (golden hybrid_v2), worst first."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import immunostruct_b200 as I
from conftest import load_golden
from helpers import build_model, graph_batch, inject_eps, named_grads

DEV = "cuda"
name, cls = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("hybrid_v2", "HybridModelv2")
if len(sys.argv) > 3:
    I.set_precision(sys.argv[3])
gd = load_golden(name)
model = build_model(cls, gd, device=DEV)
g = graph_batch(gd["graph"], DEV).validate()
d = gd["dense"]
inject_eps(model, d["eps"], d["eps"], d["eps"])
recon, mu, logvar, out = model(g, d["seq"].to(DEV), d["prop"].to(DEV))
losses = I.Losses(231, [float(gd["meta"]["pos_weight"]), 1.0], sequence=True)
losses.BCE_loss(recon, d["seq"].to(DEV), mu, logvar, out, d["target"].to(DEV)).backward()
got, ref, truth = named_grads(model), gd["grads"], gd["grads64"]
gmax = max(float(v.abs().max()) for v in ref.values() if v is not None)
rows = []
for k, r in ref.items():
    if r is None or got[k] is None:
        continue
    t = truth[k].double()
    lim = max(1e-5 * max(float(r.abs().max()), 1e-2 * gmax), 3.0 * float((r.double() - t).abs().max()))
    err = float((got[k].detach().double().cpu() - t).abs().max())
    rows.append((err / lim, k, err, lim, float((r.double() - t).abs().max())))
for ratio, k, err, lim, referr in sorted(rows, reverse=True)[:8]:
    print(f"{ratio:6.3f}  {k:40s} err {err:.3e} lim {lim:.3e} ref-vs-truth {referr:.3e}")
