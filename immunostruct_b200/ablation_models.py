"""Drop-in replacements for ``immunostruct/models/ablation_models.py`` (sequence-only, structure-only,
mean+max pooling, no-property variants) built from the same fused blocks."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as IF
from .trunk import LoadTrained, SequenceVAE, StructureTrunk, classifier_mlp

__all__ = ["SequenceModel", "SequenceFpModel", "StructureModel", "StructureModel_SSL", "StructureModelv2", "DualModel"]


class SequenceModel(LoadTrained, nn.Module, SequenceVAE):          # reference ablation_models.py:10-66
    _cond = 0

    def __init__(self, vae_input_dim, device, gcn_layers=5, vae_hidden_dim=512, vae_latent_dim=32,
                 gat_hidden_channels=64, *args, **kwargs):
        super().__init__()
        self.device = device
        self._build_vae(vae_input_dim, vae_hidden_dim, vae_latent_dim, self._cond)
        self.classifier = self.get_classifier()

    def get_classifier(self):
        return classifier_mlp(self.vae_latent_dim + self._cond)

    def forward(self, graph_data, sequence_data, peptide_property, return_embedding=False, return_attention=False):
        recon_x, mu, logvar, z = self.vae_branch(sequence_data, peptide_property if self._cond else None)
        return recon_x, mu, logvar, self.classifier(z)


class SequenceFpModel(SequenceModel):                              # reference ablation_models.py:68-125
    _cond = 2


class _Structure(LoadTrained, nn.Module, StructureTrunk):
    _ssl, _maxpool = False, False

    def __init__(self, vae_input_dim, device, gcn_layers=5, vae_hidden_dim=512, vae_latent_dim=32,
                 gat_hidden_channels=64, mlp_features=32, *args, **kwargs):
        super().__init__()
        self.device = device
        self.vae_hidden_dim, self.vae_latent_dim, self.mlp_features = vae_hidden_dim, vae_latent_dim, mlp_features
        self._build_trunk(gcn_layers, gat_hidden_channels, "mha", 8)
        self.classifier = self.get_classifier()
        if self._ssl:
            self.classifier_head = nn.Linear(mlp_features, 1)
            self.node_predictor_head = nn.Linear(mlp_features, 20)

    def get_classifier(self):
        return classifier_mlp(self.gat_hidden_channels * (2 if self._maxpool else 1), with_out=not self._ssl)

    def forward(self, graph_data, sequence_data, peptide_property, return_embedding=False, return_attention=False):
        pooled, _, nodes = self.structure_embedding(graph_data, want_nodes=self._maxpool)
        if self._maxpool:
            # global_max_pool over each graph's rows (ablation_models.py:296-299): segment max kernel
            # (csrc/segment_pool.cu), ragged node counts allowed
            pooled = torch.cat([pooled, IF.segment_pool(graph_data, nodes, "max")], dim=-1)
        out = self.classifier(pooled)
        if self._ssl:
            return 0, 0, 0, self.classifier_head(out), self.node_predictor_head(out)
        return 0, 0, 0, out


class StructureModel(_Structure):                                  # reference ablation_models.py:127-180
    pass


class StructureModel_SSL(_Structure):                              # reference ablation_models.py:182-242
    _ssl, _head_attr = True, "classifier_head"


class StructureModelv2(_Structure):                                # reference ablation_models.py:244-307
    _ssl, _maxpool, _head_attr = True, True, "classifier_head"


class DualModel(LoadTrained, nn.Module, StructureTrunk, SequenceVAE):   # reference ablation_models.py:309-398
    def __init__(self, vae_input_dim, device, gcn_layers=5, vae_hidden_dim=512, vae_latent_dim=32,
                 gat_hidden_channels=64):
        super().__init__()
        self.device = device
        self._build_trunk(gcn_layers, gat_hidden_channels, "sa")
        self._build_vae(vae_input_dim, vae_hidden_dim, vae_latent_dim, 0)
        self.classifier = self.get_classifier()

    def get_classifier(self):
        return classifier_mlp(self.vae_latent_dim + self.gat_hidden_channels)

    def forward(self, graph_data, sequence_data, peptide_property, return_embedding=False, return_attention=False):
        x_gat_node, attn, _ = self.structure_embedding(graph_data, want_attn=return_attention)
        recon_x, mu, logvar, z = self.vae_branch(sequence_data)
        out = self.classifier(torch.cat([x_gat_node, z], dim=1))
        if return_embedding:
            return x_gat_node, mu, logvar, out
        if return_attention:
            return attn, mu, logvar, out
        return recon_x, mu, logvar, out
