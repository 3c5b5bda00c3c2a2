import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """tests/golden/<name>.npz -> {group: {key: torch tensor}} (None for the NaN 'no grad' marker)."""
    raw = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in raw.files:
        group, key = k.split("/", 1)
        arr = raw[k]
        t = torch.from_numpy(np.array(arr))
        if group in ("grads", "grads64") and t.numel() == 1 and torch.isnan(t).all():
            t = None
        out.setdefault(group, {})[key] = t
    return out


def oracle_graph(arrays):
    """golden 'graph' group -> the dict the oracle consumes (dgl.batch restatement)."""
    from oracle import reference_ops as R
    from immunostruct_b200.synthetic import split_graphs
    return R.dgl_batch(split_graphs(arrays))


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_grads_close(got: dict, ref: dict, tol: float, floor_frac: float = 1e-2, truth: dict = None,
                       truth_factor: float = 3.0):
    """Gradient parity.  Per parameter, with scale = max(|ref|max of that parameter, floor_frac *
    largest gradient magnitude of the whole model):
      * without ``truth``: max|got - ref| <= tol * scale;
      * with ``truth`` (the same reference code run in fp64): max|got - truth| <= max(tol * scale,
        3 * max|ref - truth|), i.e. the product may be no further from the exact gradient than the
        tolerance or ``truth_factor`` (three) times the fp32 reference's own rounding error, whichever is larger (the
        reference's fp32 gradients sit up to 1.2e-5 from their fp64 values on these inputs, and the
        maximum over thousands of elements of two independent fp32 roundings is 2-3x a single one).
    The floor keeps parameters whose true gradient is exactly zero (e.g. the key bias under softmax
    shift-invariance) from being compared on noise.  ``ref[k] is None`` demands that the product
    also reports NO gradient (``None``)."""
    gmax = max(float(v.abs().max()) for v in ref.values() if v is not None)
    bad, worst = [], ("", 0.0)
    for k, r in ref.items():
        g = got[k]
        if r is None:
            if g is not None:
                bad.append((k, "expected grad None"))
            continue
        if g is None:
            bad.append((k, "missing grad"))
            continue
        g = g.detach().double().cpu()
        lim = tol * max(float(r.abs().max()), floor_frac * gmax)
        if truth is not None:
            t = truth[k].double()
            lim = max(lim, truth_factor * float((r.double() - t).abs().max()))
            err = float((g - t).abs().max())
        else:
            err = float((g - r.double()).abs().max())
        if lim > 0 and err / lim > worst[1]:
            worst = (k, err / lim)
        if not err <= lim:
            bad.append((k, f"err {err:.3e} > {lim:.3e}"))
    try:                                    # margin log (gpurun_out/kernel_errors.txt): worst error / bound of this call
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "kernel_errors.txt"), "a") as f:
            test = os.environ.get("PYTEST_CURRENT_TEST", "").split("::")[-1]
            f.write(f"{test} | gradient parity: worst err / bound = {worst[1]:.3f} ({worst[0]})\n")
    except OSError:
        pass
    assert not bad, bad
