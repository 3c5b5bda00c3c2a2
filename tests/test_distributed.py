"""N > 1 host logic on CPU: world_size-2 gloo process group, shard ranges, flat gradient all-reduce
(including parameters whose gradient stays None)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from immunostruct_b200.distributed import GradientAllReducer, broadcast_parameters, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 512, 27000, 1000000):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init per rank on purpose
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3), torch.nn.Linear(3, 3))
    broadcast_parameters(model)
    w0 = model[0].weight.detach().clone()
    red = GradientAllReducer(model.parameters())
    data = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10
    lo, hi = shard_range(8, rank, world)
    out = model[1](model[0](data[lo:hi]))               # model[2] unused -> grad None
    (out.pow(2).sum() / 8).backward()                   # sum over shards == full-batch mean * ... below
    red.step()
    q.put((rank, w0.tolist(), model[0].weight.grad.tolist(), model[2].weight.grad is None, red.nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, w_a, g_a, none_a, nb), (_, w_b, g_b, none_b, _) = res
    w_a, g_a, w_b, g_b = (torch.tensor(t) for t in (w_a, g_a, w_b, g_b))
    assert torch.equal(w_a, w_b)                        # broadcast made the replicas identical
    assert torch.equal(g_a, g_b) and none_a and none_b  # averaged grads agree; unused layer untouched
    assert nb == (5 * 4 + 4 + 4 * 3 + 3) * 4
    # reference value: single process, full batch, same weights
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3), torch.nn.Linear(3, 3))
    data = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10
    (model[1](model[0](data)).pow(2).sum() / 8).backward()
    assert torch.allclose(g_a * 2, model[0].weight.grad, rtol=1e-5, atol=1e-6)   # mean of 2 shard-sums = total / 2
