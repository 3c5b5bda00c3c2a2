"""Drop-in for ``immunostruct/utils/loss.py``: same ``Losses`` callables, one fused CUDA reduction
(csrc/loss.cu) per call instead of a chain of elementwise + mean kernels."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import functional as IF

__all__ = ["Losses"]


class Losses:
    def __init__(self, vae_input_dim, class_weights, sequence=True):
        self.vae_input_dim = vae_input_dim
        self.sequence = sequence
        # reference loss.py:11 keeps pos_weight = n_neg / n_pos as a 0-d CPU tensor
        self.pos_weight = torch.tensor(float(class_weights[0]) / float(class_weights[1])).float()

    def _fused(self, mode, w, recon_x, x, mu, logvar, final_output, y):
        w_pred, w_rest = (w, None) if not self.sequence else w
        if not self.sequence:
            return IF.fused_loss(None, None, None, None, final_output, y, mode, float(self.pos_weight), 1.0, 0.0, 0.0)
        return IF.fused_loss(recon_x, x, mu, logvar, final_output, y, mode, float(self.pos_weight),
                             w_pred, w_rest, w_rest)

    def regression_loss(self, recon_x, x, mu, logvar, final_output, y):
        """2.0 * MSE(out, y) + 0.5 * MSE(recon, seq) + 0.5 * KLD   (reference loss.py:13-21)."""
        return self._fused(1, (2.0, 0.5), recon_x, x, mu, logvar, final_output, y)

    def BCE_loss(self, recon_x, x, mu, logvar, final_output, y):
        """5.0 * BCEWithLogits(out, y, pos_weight) + 0.1 * MSE + 0.1 * KLD   (reference loss.py:23-31)."""
        return self._fused(0, (5.0, 0.1), recon_x, x, mu, logvar, final_output, y)

    @staticmethod
    def _amino(pred_amino_acid, amino_acid):
        return F.cross_entropy(pred_amino_acid, amino_acid) if pred_amino_acid.numel() else 0

    def regression_loss_SSL(self, recon_x, x, mu, logvar, final_output, y, pred_amino_acid, amino_acid):
        """reference loss.py:33-46."""
        return self.regression_loss(recon_x, x, mu, logvar, final_output, y) + self._amino(pred_amino_acid, amino_acid)

    def BCE_loss_SSL(self, recon_x, x, mu, logvar, final_output, y, pred_amino_acid, amino_acid):
        """reference loss.py:48-61."""
        return self.BCE_loss(recon_x, x, mu, logvar, final_output, y) + self._amino(pred_amino_acid, amino_acid)
