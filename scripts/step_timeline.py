"""Kernel timeline of the inference (default) or training step (argument `train`) from torch.profiler (CUPTI): per kernel its
duration and the idle gap since the previous kernel ended; sums per step.  python scripts/step_timeline.py [train] [bf16x3 | fp16x2 | bf16 | ...]"""
import sys
import torch
sys.path.insert(0, ".")
import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_graph_arrays, synthetic_dense

train = "train" in sys.argv[1:]
dev = "cuda"
B = 512
I.set_precision(next((a for a in sys.argv[1:] if a in ("bf16x3", "fp16x2", "bf16", "tf32x3", "fp32")), "bf16x3"))
torch.manual_seed(1)
model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev)
arr = synthetic_graph_arrays(B, 200, 10, seed=1, device=dev)
dense = synthetic_dense(B, seed=1, device=dev)
losses = I.Losses(5943, [0.81, 0.19], sequence=True)
keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")


def step():
    gb = GraphBatch.from_arrays(*(arr[k] for k in keys), max_nodes=200)
    if train:
        recon, mu, logvar, out = model(gb, dense["seq"], dense["prop"])
        loss = losses.BCE_loss(recon, dense["seq"], mu, logvar, out, dense["target"])
        loss.backward()
        return loss
    with torch.no_grad():
        return torch.sigmoid(model(gb, dense["seq"], dense["prop"])[3])


model.train(train)
for _ in range(5):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
n = len(ev) // 3
ev = ev[2 * n:]                       # last step
t_end = None
tot_k = tot_gap = 0.0
for e in ev:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = 0.0 if t_end is None else s - t_end
    t_end = max(t_end or 0, e.time_range.end)
    tot_k += d
    tot_gap += max(gap, 0.0)
    print(f"{d:9.1f} us  gap {gap:7.1f}  {e.name[:90]}")
print(f"kernels {len(ev)}: busy {tot_k:.1f} us, gaps {tot_gap:.1f} us, span {ev[-1].time_range.end - ev[0].time_range.start:.1f} us")
