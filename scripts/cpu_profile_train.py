"""Host-side cost of one training step: cProfile over 20 eager steps (the step is ~200 launches in ~11 ms, so the
python / ctypes / autograd overhead per launch decides whether the GPU or the host is the bottleneck)."""
import cProfile
import pstats
import sys
import time
import torch
sys.path.insert(0, ".")
import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_dense, synthetic_graph_arrays

dev = "cuda"
keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")
arr = synthetic_graph_arrays(512, 200, 10, seed=1, device=dev)
dense = synthetic_dense(512, seed=1, device=dev)
torch.manual_seed(1)
model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev).train()
opt = I.FusedAdam(model.parameters(), lr=1e-3)
losses = I.Losses(5943, [0.81, 0.19], sequence=True)


def step():
    gb = GraphBatch.from_arrays(*(arr[k] for k in keys), max_nodes=200)
    opt.zero_grad()
    recon, mu, logvar, out = model(gb, dense["seq"], dense["prop"])
    loss = losses.BCE_loss(recon, dense["seq"], mu, logvar, out, dense["target"])
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
# host time alone: enqueue 20 steps without waiting for the device
t0 = time.perf_counter()
for _ in range(20):
    step()
t_host = (time.perf_counter() - t0) / 20
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 20
print(f"host enqueue time per step {t_host * 1e3:.2f} ms; with device sync {t_all * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
