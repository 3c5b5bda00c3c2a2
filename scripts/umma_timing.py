"""Cycles per tcgen05.mma (96 MMAs, N = 64, bf16, K = 16; issue to completion, unrolled with compile-time descriptors) for
the operand layouts of the edge kernels.  SWIZZLE_NONE tiles: 8-row x 16-byte core matrices LBO bytes apart along K, SBO
bytes apart along M / N; "k acc" = the sequence rotates over k TMEM accumulators (1 = a dependent chain)."""
import sys
import torch
sys.path.insert(0, ".")
from immunostruct_b200 import _C
names = ["M128 K-major A 128/1024, B 128/1024, 1 acc", "same, 2 acc", "same, 4 acc", "A 144/1152 (padded), B 128/1024, 1 acc",
         "A 144/1152, B 144/1152, 1 acc", "A 144/1152, B 144/1152, 4 acc", "A 128/1040, B 128/1040 (SBO padded), 1 acc",
         "A from TMEM, B 128/1024, 1 acc", "A from TMEM, B 128/1024, 4 acc", "dgrad: B = MN-major view of 128/1024, 1 acc",
         "dgrad: A 144/1152, B = MN-major view of 144/1152, 1 acc", "wgrad M64: A, B MN-major views of 128/1024, 1 acc",
         "wgrad M64: same, 2 acc", "wgrad M64: views of 144/1152, 1 acc", "segment sum M64: A 128/2048, B MN-major view of 144/1152",
         "M64 K-major 128/1024 both, 1 acc"]
out = torch.zeros(16, device="cuda")
for _ in range(3):
    _C.umma_timing(out)
torch.cuda.synchronize()
for n, v in zip(names, out.tolist()):
    print(f"{v:7.1f} cycles / MMA   {n}")
