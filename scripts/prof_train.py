"""One training step (fwd + bwd + Adam) of HybridModelv2 at batch 512 inside a cudaProfiler range (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_dense, synthetic_graph_arrays

dev = "cuda"
I.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
arr = synthetic_graph_arrays(512, 200, 10, seed=1, device=dev)
dense = synthetic_dense(512, seed=1, device=dev)
torch.manual_seed(1)
model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
losses = I.Losses(5943, [0.81, 0.19], sequence=True)
keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")


def step():
    gb = GraphBatch.from_arrays(*(arr[k] for k in keys), max_nodes=200)
    opt.zero_grad(set_to_none=True)
    recon, mu, logvar, out = model(gb, dense["seq"], dense["prop"])
    loss = losses.BCE_loss(recon, dense["seq"], mu, logvar, out, dense["target"])
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
