"""``dgl.dataloading.GraphDataLoader`` stand-in: with ``use_ddp=False`` it is a plain DataLoader with the
user's ``collate_fn`` (reference train_IEDB_wFT.py:86)."""
import torch

from .graph import collate


class GraphDataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, collate_fn=None, **kwargs):
        kwargs.pop("use_ddp", None)
        super().__init__(dataset, collate_fn=collate_fn or collate, **kwargs)
