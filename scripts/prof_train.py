"""One training step inside a cudaProfiler range (for ncu): HybridModelv2 fwd + bwd + optimizer at batch 512, or with
`comparative` the cancer fine-tune step (HybridModelv2_Comparative, 256 pairs, contrastive loss, AdamW).

  python scripts/prof_train.py [precision] [train|comparative] [torch|fused]"""
import sys
import torch
sys.path.insert(0, ".")
import immunostruct_b200 as I
from immunostruct_b200.graph import GraphBatch
from immunostruct_b200.synthetic import synthetic_dense, synthetic_graph_arrays

dev = "cuda"
I.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
optim = sys.argv[3] if len(sys.argv) > 3 else "fused"
keys = ("x", "src", "dst", "edge_attr", "node_counts", "edge_counts")
losses = I.Losses(5943, [0.81, 0.19], sequence=True)
torch.manual_seed(1)

if mode == "train":
    arr = synthetic_graph_arrays(512, 200, 10, seed=1, device=dev)
    dense = synthetic_dense(512, seed=1, device=dev)
    model = I.model_map["HybridModelv2"](vae_input_dim=5943, device=dev).to(dev).train()
    opt = I.FusedAdam(model.parameters(), lr=1e-3) if optim == "fused" else torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        gb = GraphBatch.from_arrays(*(arr[k] for k in keys), max_nodes=200)
        opt.zero_grad()
        recon, mu, logvar, out = model(gb, dense["seq"], dense["prop"])
        loss = losses.BCE_loss(recon, dense["seq"], mu, logvar, out, dense["target"])
        loss.backward()
        opt.step()
        return loss
else:
    pools = [(synthetic_graph_arrays(256, 200, 10, seed=s, device=dev), synthetic_dense(256, seed=s, device=dev)) for s in (1, 2)]
    model = I.model_map["HybridModelv2_Comparative"](vae_input_dim=5943, device=dev, use_wt_for_downstream=True).to(dev).train()
    opt = (I.FusedAdamW if optim == "fused" else torch.optim.AdamW)(model.parameters(), lr=1e-4, weight_decay=1e-6)
    pcl = I.PairedContrastiveLoss(embedding_dim=104, device=dev)

    def step():
        (ac, dc), (aw, dw) = pools
        gc = GraphBatch.from_arrays(*(ac[k] for k in keys), max_nodes=200)
        gw = GraphBatch.from_arrays(*(aw[k] for k in keys), max_nodes=200)
        embs, recons, mus, lvs, out = model.forward_comparative((gc, gw), (dc["seq"], dw["seq"]), (dc["prop"], dw["prop"]))
        y = dc["target"]
        lc = losses.BCE_loss(recons[0], dc["seq"], mus[0], lvs[0], out, y)
        lw = losses.BCE_loss(recons[1], dw["seq"], mus[1], lvs[1], out, y)
        loss = (lc + lw) / 2 + 0.01 * pcl(embs[0], embs[1], y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"{mode} step wall: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
