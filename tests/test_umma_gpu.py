"""tcgen05 building blocks (descriptors, canonical smem layout, TMEM load) against torch matmul."""
import pytest
import torch

from immunostruct_b200 import _C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [(0, 1e-2), (1, 2e-3), (2, 2e-6), (3, 2e-6), (4, 1e-2), (5, 1e-2), (7, 1e-2), (8, 1e-2)])
def test_umma_selftest(mode, tol):
    gen = torch.Generator().manual_seed(5 + mode)
    A = torch.randn(128, 64, generator=gen).cuda()
    B = torch.randn(64, 64, generator=gen).cuda()
    D = torch.full((128, 64), float("nan"), device="cuda")
    _C.umma_selftest(A, B, D, mode)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T)
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    print(f"umma mode {mode}: rel err {err:.3e}")
    assert err < tol, err
