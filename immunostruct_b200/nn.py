"""``from dgl.nn import EGNNConv`` compatibility (reference models/hybrid_models.py:5)."""
from .layers import EGNNConv  # noqa: F401
