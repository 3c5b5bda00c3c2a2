"""Probe MN-major operands and the M=64 accumulator layout of tcgen05 (groundwork for a tensor-core backward)."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C

g = torch.Generator().manual_seed(3)
A = torch.randn(128, 64, generator=g).cuda()
B = torch.randn(64, 64, generator=g).cuda()
ref = A.double() @ B.double().T
for mode in (4, 5):
    D = torch.full((128, 64), float("nan"), device="cuda")
    _C.umma_selftest(A, B, D, mode)
    torch.cuda.synchronize()
    err = float((D.double() - ref).abs().max() / ref.abs().max())
    print(f"mode {mode} ({'B' if mode == 4 else 'A'} MN-major): rel err {err:.3e}")
D = torch.full((128, 64), float("nan"), device="cuda")
torch.cuda.synchronize()
# pre-fill TMEM-visible output with a first full-size MMA of zeros? not needed: lanes untouched by M=64 stay garbage
_C.umma_selftest(A, B, D, 6)
torch.cuda.synchronize()
ref64 = ref[:64]
# for every TMEM lane find the reference row it matches (if any)
mapping = []
for lane in range(128):
    d = (ref64.float().cuda() - D[lane]).abs().max(dim=1).values / ref64.abs().max().float().cuda()
    r = int(d.argmin())
    mapping.append((lane, r if float(d[r]) < 2e-2 else None))
print("M=64: TMEM lane -> D row:", [(l, r) for l, r in mapping if r is not None][:80])
print("lanes holding rows:", sum(1 for _, r in mapping if r is not None))
