"""Large-batch sanity: batch-composition independence at 2 048 graphs (the first 64 graphs give the same logits alone)."""
import sys
import torch
sys.path.insert(0, ".")
import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_dense, synthetic_graph_arrays

dev = "cuda"
arr = synthetic_graph_arrays(2048, 200, 10, seed=5, device=dev)
dense = synthetic_dense(2048, seed=5, device=dev)
torch.manual_seed(1)
model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev).eval()
keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")
eps = torch.randn(2048, 32, device=dev)
model.sample_eps = lambda like: eps[: like.shape[0]].to(like)
with torch.no_grad():
    big = model(GraphBatch.from_arrays(*(arr[k] for k in keys), max_nodes=200), dense["seq"], dense["prop"])[3]
    sub = {"x": arr["x"][: 64 * 200], "src": arr["src"][: 64 * 2000], "dst": arr["dst"][: 64 * 2000],
           "edge_attr": arr["edge_attr"][: 64 * 2000], "node_counts": arr["node_counts"][:64], "edge_counts": arr["edge_counts"][:64]}
    small = model(GraphBatch.from_arrays(*(sub[k] for k in keys), max_nodes=200), dense["seq"][:64], dense["prop"][:64])[3]
diff = float((big[:64] - small).abs().max())
print("max |diff| of the first 64 logits, batch 2048 vs batch 64:", diff, "finite:", bool(torch.isfinite(big).all()))
assert diff <= 1e-5 * float(small.abs().max()) + 1e-7 and bool(torch.isfinite(big).all())
