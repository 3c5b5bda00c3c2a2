"""Drop-in replacements for ``immunostruct/models/hybrid_models.py`` (same class names, constructor
arguments, forward signature, return tuples and state_dict keys); the structure trunk, pooling and
fusion attention run in the hand-written sm_100a kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .layers import MultiHeadAttention
from .trunk import LoadTrained, SequenceVAE, StructureTrunk, classifier_mlp, property_mlp

__all__ = ["HybridModel", "HybridModelv2", "HybridModel_SSL", "HybridModelv2_SSL"]


class _Hybrid(LoadTrained, nn.Module, StructureTrunk, SequenceVAE):
    """Common body of the four hybrid models.  ``fusion_dim``: None = plain concat (v1), else the
    feature_dim of ``combined_attention`` (16 for v2, 32 for v2-SSL); ``ssl``: two heads."""

    _attention, _fusion_dim, _ssl = "sa", None, False

    def __init__(self, vae_input_dim, device, gcn_layers=5, vae_hidden_dim=512, vae_latent_dim=32,
                 gat_hidden_channels=64, property_embedding_dim=8, self_attention_heads=1,
                 combined_attention_heads=8, mlp_features=32, *args, **kwargs):
        super().__init__()
        self.device = device
        self.property_embedding_dim = property_embedding_dim
        self.mlp_features = mlp_features
        self._build_trunk(gcn_layers, gat_hidden_channels, self._attention, self_attention_heads)
        self._build_vae(vae_input_dim, vae_hidden_dim, vae_latent_dim, property_embedding_dim)
        if self._fusion_dim:
            self.combined_attention = MultiHeadAttention(self._fusion_dim, combined_attention_heads, input_dim=1)
        self.classifier = self.get_classifier()
        if self._ssl:
            self.classifier_head = nn.Linear(mlp_features, 1)
            self.node_predictor_head = nn.Linear(mlp_features, 20)
        self.property_embedding = property_mlp(property_embedding_dim)

    def get_classifier(self):
        return classifier_mlp(self.vae_latent_dim + self.property_embedding_dim + self.gat_hidden_channels,
                              with_out=not self._ssl)

    def _fused_eval_ok(self, graph_data, peptide_property, return_attention):
        """The eval-mode small-layer fusions (csrc/head.cu) apply: no autograd, dropout inactive, CUDA inputs, the
        reference's layer shapes, tensor-core arithmetic mode, attention weights not requested."""
        from . import functional as IF
        return (not torch.is_grad_enabled() and not self.training and not return_attention
                and "reparameterize" not in self.__dict__          # a patched reparameterize must be honoured
                and IF._PRECISIONS[IF.get_precision()] is not None and peptide_property.is_cuda
                and peptide_property.dtype == torch.float32 and self.property_embedding[0].weight.shape == (32, 2)
                and self.gat_hidden_channels == 64 and self.mlp_features == 32
                and self.vae_latent_dim + self.property_embedding_dim <= 192)

    def _forward_eval_fused(self, graph_data, sequence_data, peptide_property, return_embedding):
        from . import functional as IF
        pooled, _, _ = self.structure_embedding(graph_data, project=False)
        recon_x, mu, logvar, z_vae = self.vae_branch_eval(sequence_data, self.property_embedding, peptide_property)
        wc = self.self_attention.out_projection
        fus = self.combined_attention if self._fusion_dim else None
        cls1, cls2 = self.classifier[1], (None if self._ssl else self.classifier[4])
        x_gat_node, out = IF._C.head_infer(
            pooled, None if wc is None else wc.weight.detach(), None if wc is None else wc.bias.detach(), z_vae,
            None if fus is None else fus.fusion_coefficients(), 0 if fus is None else fus.n_head,
            cls1.weight.detach(), cls1.bias.detach(), None if cls2 is None else cls2.weight.detach(),
            None if cls2 is None else cls2.bias.detach())
        tail = (self.classifier_head(out), self.node_predictor_head(out)) if self._ssl else (out,)
        if return_embedding:
            return (x_gat_node, mu, logvar) + tail
        return (recon_x, mu, logvar) + tail

    def forward(self, graph_data, sequence_data, peptide_property, return_embedding=False, return_attention=False):
        if self._fused_eval_ok(graph_data, peptide_property, return_attention):
            return self._forward_eval_fused(graph_data, sequence_data, peptide_property, return_embedding)
        x_gat_node, attention_weights, _ = self.structure_embedding(graph_data, want_attn=return_attention)
        peptide_property = self.property_embedding(peptide_property)          # consumes dropout RNG first
        recon_x, mu, logvar, z_vae = self.vae_branch(sequence_data, peptide_property)
        combined = torch.cat([x_gat_node, z_vae], dim=1)
        if self._fusion_dim:
            combined = self.combined_attention.fused_mean(combined)
        out = self.classifier(combined)
        tail = (self.classifier_head(out), self.node_predictor_head(out)) if self._ssl else (out,)
        if return_embedding:
            return (x_gat_node, mu, logvar) + tail
        if return_attention:
            return (attention_weights, mu, logvar) + tail
        return (recon_x, mu, logvar) + tail


class HybridModel(_Hybrid):            # reference hybrid_models.py:10-119
    pass


class HybridModel_SSL(_Hybrid):        # reference hybrid_models.py:121-238
    _ssl, _head_attr = True, "classifier_head"


class HybridModelv2(_Hybrid):          # reference hybrid_models.py:240-359
    _attention, _fusion_dim = "mha", 16


class HybridModelv2_SSL(_Hybrid):      # reference hybrid_models.py:361-488
    _attention, _fusion_dim, _ssl, _head_attr = "mha", 32, True, "classifier_head"
