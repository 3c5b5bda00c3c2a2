"""Helpers shared by tests/golden/make_golden_r2.py (reference side) and the parity tests (product side): both must
build identical weights, objectives and samples without storing multi-megabyte tensors in the fixtures."""
import zlib

import torch


def seeded_state_dict(model, seed, scale=1.0):
    """Weights keyed on the PARAMETER NAME, not on module construction order (the product registers its modules in a
    different order than the reference): every floating tensor of ``state_dict`` is drawn from a generator seeded by
    ``crc32(name) + seed``; matrices ~ N(0, (0.6 / sqrt(fan_in))^2) -- the spread of torch's default Linear init --
    with the EGNN weights 1.5 x larger so that the outputs depend visibly on the graph (the regime of the round-1
    golden files: well-conditioned enough that the fp32 reference itself sits within ~1e-5 of its fp64 run); vectors
    ~ N(0, 0.1); BatchNorm statistics stay at init."""
    out = {}
    for name, t in model.state_dict().items():
        if not t.is_floating_point() or "running_" in name:
            out[name] = t.clone()
            continue
        gen = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) % (2 ** 31))
        if t.dim() >= 2:
            std = 0.6 / t.shape[-1] ** 0.5 * scale * (1.5 if name.startswith("GCN_layers") else 1.0)
            out[name] = (torch.randn(t.shape, generator=gen) * std).to(t.dtype)
        else:
            out[name] = (torch.randn(t.shape, generator=gen) * 0.1).to(t.dtype)
    return out


def flatten_outputs(outs):
    """Tensor leaves of a model's return value (tuples / lists nested at most twice), python scalars dropped."""
    flat = []
    for o in outs:
        if torch.is_tensor(o):
            flat.append(o)
        elif isinstance(o, (tuple, list)):
            flat += [t for t in o if torch.is_tensor(t)]
    return flat


def objective(flat):
    """Fixed weighted sum of every output element (weights 0.5 .. 1.5 along the flattened tensor, scaled per output)."""
    total = 0
    for i, t in enumerate(flat):
        w = torch.linspace(0.5, 1.5, t.numel(), dtype=t.dtype, device=t.device).view_as(t)
        total = total + (1.0 + 0.25 * i) * (t * w).sum() / max(t.numel(), 1) ** 0.5
    return total


def sample_big(t, limit=4096, stride=997):
    """Tensors above ``limit`` elements are stored as [norm, strided sample...]; smaller ones unchanged."""
    if not torch.is_tensor(t) or t.numel() <= limit:
        return t
    f = t.detach().reshape(-1)
    return torch.cat([f.double().norm().reshape(1).to(f.dtype), f[::stride]])
