"""Drop-in for ``immunostruct/utils/contrastive.py``: cancer-vs-wild-type paired contrastive loss.

Same module (``projector`` Sequential: Linear(104,128,no bias) - BatchNorm1d - ReLU - Linear(128,128,
no bias), randomly initialised, BatchNorm in batch-statistics mode, identical ``state_dict``) and the same
arithmetic, computed by the kernels of ``csrc/contrastive.cu`` (forward 9 launches, backward 9; no cuBLAS / ATen
kernels):
  * the "exactly two target values" gate (reference contrastive.py:38-43) is evaluated ON THE DEVICE and applied
    as a 0/1 factor, so there is no ``unique()`` host synchronisation per step (the reference returns the python
    int 0; here a zero tensor with zero gradient -- identical once added to the loss);
  * BatchNorm running statistics / ``num_batches_tracked`` are updated for the cancer call and then the wild-type
    call, and only when the gate is open -- exactly the batches on which the reference reaches the projector;
  * a batch of fewer than two pairs returns 0 before anything is launched (the reference's gate closes for it as
    well: one sample cannot hold two target values).
"""
from __future__ import annotations

import torch

from . import functional as IF

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
        assert embedding_cancer.shape == embedding_wt.shape
        if embedding_cancer.shape[0] < 2:
            return embedding_cancer.sum() * 0.0                   # nothing to contrast (reference: python 0)
        lin1, bn, _, lin2 = self.projector
        assert self.z_dim == lin2.weight.shape[0]
        if not bn.training:
            raise NotImplementedError("PairedContrastiveLoss runs its BatchNorm on batch statistics (the reference never "
                                      "puts the module in eval mode, procedures/train.py:74-78)")
        if bn.momentum is None or not bn.affine:
            raise NotImplementedError("BatchNorm1d with momentum=None or affine=False")
        track = bn.track_running_stats and bn.running_mean is not None
        return IF.paired_contrastive(embedding_cancer.float(), embedding_wt.float(), is_immunogenic, lin1.weight, bn.weight,
                                     bn.bias, lin2.weight, bn.running_mean if track else None,
                                     bn.running_var if track else None, bn.num_batches_tracked if track else None,
                                     bn.eps, bn.momentum, self.lambda_off_diag)
