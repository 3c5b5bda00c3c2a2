"""Drop-in replacements for ``immunostruct/models/comparative_models.py``: the cancer / wild-type
pair models (same trunk run on both items, embeddings concatenated for the classifier)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .layers import MultiHeadAttention
from .trunk import LoadTrained, SequenceVAE, StructureTrunk, classifier_mlp, property_mlp

__all__ = ["HybridModel_Comparative", "HybridModel_Comparative_SSL",
           "HybridModelv2_Comparative", "HybridModelv2_Comparative_SSL"]


class _Comparative(LoadTrained, nn.Module, StructureTrunk, SequenceVAE):
    _attention, _fusion_dim, _ssl = "sa", None, False

    def __init__(self, vae_input_dim, device, gcn_layers=5, vae_hidden_dim=512, vae_latent_dim=32,
                 gat_hidden_channels=64, property_embedding_dim=8, self_attention_heads=1,
                 combined_attention_heads=8, use_wt_for_downstream=True, mlp_features=32, *args, **kwargs):
        super().__init__()
        self.device = device
        self.property_embedding_dim = property_embedding_dim
        self.use_wt_for_downstream = use_wt_for_downstream
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
        d = self.vae_latent_dim + self.property_embedding_dim + self.gat_hidden_channels
        return classifier_mlp(d * 2 if self.use_wt_for_downstream else d, with_out=not self._ssl)

    def forward_item(self, graph_data, sequence_data, peptide_property, want_attn=False):
        """reference comparative_models.py:433-461: (mu, logvar, x_gat_node, z_vae, attention, recon_x)."""
        x_gat_node, attn, _ = self.structure_embedding(graph_data, want_attn=want_attn)
        peptide_property = self.property_embedding(peptide_property)
        recon_x, mu, logvar, z_vae = self.vae_branch(sequence_data, peptide_property)
        return mu, logvar, x_gat_node, z_vae, attn, recon_x

    def _head(self, combined):
        if self._fusion_dim:
            combined = self.combined_attention.fused_mean(combined)
        out = self.classifier(combined)
        return (self.classifier_head(out), self.node_predictor_head(out)) if self._ssl else (out,)

    def forward_comparative(self, graph_data_pair, sequence_data_pair, peptide_property_pair,
                            return_embedding=False, return_attention=False):
        """reference comparative_models.py:463-496: cancer item first, then wild type (RNG order)."""
        mu_c, lv_c, g_c, z_c, attn_c, rec_c = self.forward_item(
            graph_data_pair[0], sequence_data_pair[0], peptide_property_pair[0], want_attn=return_attention)
        mu_w, lv_w, g_w, z_w, _, rec_w = self.forward_item(
            graph_data_pair[1], sequence_data_pair[1], peptide_property_pair[1])
        emb_c, emb_w = torch.cat([g_c, z_c], dim=1), torch.cat([g_w, z_w], dim=1)
        combined = torch.cat([emb_c, emb_w], dim=1) if self.use_wt_for_downstream else emb_c
        tail = self._head(combined)
        if return_embedding:
            return (g_c, mu_c, lv_c) + tail
        if return_attention:
            return (attn_c, mu_c, lv_c) + tail
        return ([emb_c, emb_w], [rec_c, rec_w], [mu_c, mu_w], [lv_c, lv_w]) + tail

    def forward(self, graph_data, sequence_data, peptide_property, return_embedding=False, return_attention=False):
        """Single-graph (pre-training) path; repeats the features to fill the pair-wide classifier
        (reference "hot fix", comparative_models.py:498-527)."""
        mu, logvar, g, z, attn, recon_x = self.forward_item(graph_data, sequence_data, peptide_property,
                                                            want_attn=return_attention)
        combined = torch.cat([g, z, g, z], dim=1) if self.use_wt_for_downstream else torch.cat([g, z], dim=1)
        tail = self._head(combined)
        if return_embedding:
            return (g, mu, logvar) + tail
        if return_attention:
            return (attn, mu, logvar) + tail
        return (recon_x, mu, logvar) + tail


class HybridModel_Comparative(_Comparative):            # reference comparative_models.py:11-173
    pass


class HybridModel_Comparative_SSL(_Comparative):        # reference comparative_models.py:175-350
    _ssl, _head_attr = True, "classifier_head"


class HybridModelv2_Comparative(_Comparative):          # reference comparative_models.py:353-527
    _attention, _fusion_dim = "mha", 32


class HybridModelv2_Comparative_SSL(_Comparative):      # reference comparative_models.py:529-713
    _attention, _fusion_dim, _ssl, _head_attr = "mha", 32, True, "classifier_head"
