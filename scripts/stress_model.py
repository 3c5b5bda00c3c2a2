"""Model-level stress run: HybridModelv2 inference and training steps (forward, loss, backward) over many batch shapes
(graphs x nodes x in-degree, incl. a single graph and ragged tails) in both tensor-core precisions, several repeats each;
every output / gradient must be finite and a repeat must reproduce the first run bit for bit.  Exits non-zero otherwise."""
import sys
import torch
sys.path.insert(0, ".")
import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_graph_arrays, synthetic_dense

dev = "cuda"
keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")
torch.manual_seed(1)
model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev)
losses = I.Losses(5943, [0.81, 0.19], sequence=True)
n_fail = 0
for prec in ("bf16x3", "fp16x2", "bf16"):
    I.set_precision(prec)
    for n_graphs, n_nodes, k in ((1, 200, 10), (2, 50, 4), (5, 190, 10), (17, 200, 10), (64, 128, 20), (129, 200, 10), (376, 200, 10), (512, 200, 10)):
        arr = synthetic_graph_arrays(n_graphs, n_nodes, k, seed=n_graphs, device=dev)
        dense = synthetic_dense(n_graphs, seed=n_graphs, device=dev)
        outs = []
        for rep in range(3):
            model.eval()
            with torch.no_grad():
                torch.manual_seed(7)
                p = torch.sigmoid(model(GraphBatch.from_arrays(*(arr[kk] for kk in keys), max_nodes=n_nodes), dense["seq"], dense["prop"])[3])
            model.train()
            model.zero_grad(set_to_none=True)
            torch.manual_seed(7)
            recon, mu, logvar, out = model(GraphBatch.from_arrays(*(arr[kk] for kk in keys), max_nodes=n_nodes), dense["seq"], dense["prop"])
            loss = losses.BCE_loss(recon, dense["seq"], mu, logvar, out, dense["target"])
            loss.backward()
            gsum = torch.stack([q.grad.double().abs().sum() for q in model.parameters() if q.grad is not None]).sum()
            torch.cuda.synchronize()
            outs.append((p.clone(), loss.detach().clone(), gsum))
        fin = all(bool(torch.isfinite(t).all()) for o in outs for t in o)
        same = all(torch.equal(outs[0][i], o[i]) for o in outs[1:] for i in range(3))
        ok = fin and same
        n_fail += 0 if ok else 1
        print(f"{prec} graphs {n_graphs:4d} n {n_nodes} k {k}: loss {float(outs[0][1]):.6f} |grad| {float(outs[0][2]):.4e} finite {fin} reproducible {same}", flush=True)
print("failures:", n_fail)
sys.exit(1 if n_fail else 0)
