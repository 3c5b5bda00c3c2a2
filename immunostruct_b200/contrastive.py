"""Drop-in for ``immunostruct/utils/contrastive.py``: cancer-vs-wild-type paired contrastive loss.

Same module (``projector`` Sequential: Linear(104,128,no bias) - BatchNorm1d - ReLU - Linear(128,128,
no bias), randomly initialised, BatchNorm in batch-statistics mode) and the same arithmetic, but:
  * the "exactly two target values" gate is evaluated ON THE DEVICE and applied as a 0/1 factor, so
    there is no ``unique()`` host synchronisation per step (reference contrastive.py:38-43 returns the
    python int 0; here a zero tensor with zero gradient -- identical once added to the loss);
  * masks are built without boolean-index assignment (which is a scatter under
    ``torch.use_deterministic_algorithms``).
The four GEMMs ([B,104]x[104,128], [B,128]x[128,128], [B,128]x[128,B], [128,B]x[B,128]) are plain
library GEMMs (55.6 MFLOP per 256-pair batch).
"""
from __future__ import annotations

import torch

__all__ = ["PairedContrastiveLoss"]


class PairedContrastiveLoss(torch.nn.Module):
    def __init__(self, embedding_dim: int = 104, z_dim: int = 128, lambda_off_diag: float = 1e-2,
                 device: torch.device = torch.device("cpu")):
        super().__init__()
        self.z_dim = z_dim
        self.lambda_off_diag = lambda_off_diag
        self.projector = torch.nn.Sequential(
            torch.nn.Linear(embedding_dim, z_dim, bias=False),
            torch.nn.BatchNorm1d(z_dim),
            torch.nn.ReLU(inplace=True),
            torch.nn.Linear(z_dim, z_dim, bias=False))
        self.device = device
        self.projector.to(self.device)

    def forward(self, embedding_cancer, embedding_wt, is_immunogenic):
        t = is_immunogenic.reshape(-1).float()
        lo, hi = t.min(), t.max()
        two_classes = ((lo != hi) & ((t == lo) | (t == hi)).all()).float()      # device-side gate
        imm = (t > t.mean()).float()

        assert embedding_cancer.shape == embedding_wt.shape
        z_c = self.projector(embedding_cancer)
        z_w = self.projector(embedding_wt)
        b = z_c.shape[0]
        assert self.z_dim == z_c.shape[1]
        z_c = z_c - z_c.mean(0)
        z_w = z_w - z_w.mean(0)
        std_loss = (torch.relu(1 - torch.sqrt(z_c.var(dim=0) + 0.0001)).mean() / 2
                    + torch.relu(1 - torch.sqrt(z_w.var(dim=0) + 0.0001)).mean() / 2)
        pair_sim = torch.mm(z_c, z_w.T) / self.z_dim
        cross_corr = torch.mm(z_c.T, z_w) / b
        eye_b = torch.eye(b, device=z_c.device, dtype=z_c.dtype)
        w_b = eye_b + (1 - eye_b) * self.lambda_off_diag
        pair_diff = ((pair_sim - eye_b * imm.unsqueeze(1)).pow(2) * w_b).sum()
        eye_z = torch.eye(self.z_dim, device=z_c.device, dtype=z_c.dtype)
        w_z = eye_z + (1 - eye_z) * self.lambda_off_diag
        corr_diff = ((cross_corr - eye_z).pow(2) * w_z).sum()
        return (pair_diff + corr_diff + std_loss) * two_classes
