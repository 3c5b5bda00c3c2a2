"""On-disk graph ingest without torch_geometric.

The reference's graphs are pickled ``torch_geometric.data.Data`` objects (``torch.save`` of the output of
Graphein's converter, preprocessing/cancer_graph_construction_new_KBG.py:137-143) read back with ``torch.load`` in
``preprocess_graphs`` (data/preprocess.py:15-43).  Unpickling them normally imports ``torch_geometric``; here a
restricted ``Unpickler`` maps every ``torch_geometric.*`` global to an inert stand-in class, and the attribute
mapping is pulled out of the storage object (PyG >= 2.0: ``Data.__dict__['_store']._mapping``; PyG 1.x: the
object's own ``__dict__``).  Nothing of torch_geometric is executed.

Equivalents of the helpers that turn those objects into model inputs:
  ``preprocess_graphs(directory)``            data/preprocess.py:15-43  (name filters, de-duplication, h-bond columns cut)
  ``pad_graph(graph, max_nodes, 23, 3)``      data/utils.py:13-33
  ``to_dgl(graph)``                           data/utils.py:54-67       -> immunostruct_b200.Graph, edge_attr = ones [E,1]
  ``append_coords(graph)``                    data/preprocess.py:335-338 (x = [x | coords], float32)
  ``preprocess_graph(mapper, 23, 3)``         data/preprocess.py:343-349 (pad to the set's max node count, to_dgl)
"""
from __future__ import annotations

import os
import pickle
from types import SimpleNamespace
from typing import Dict, Iterable, List

import torch

from .graph import Graph

__all__ = ["PygData", "load_pyg_graph", "preprocess_graphs", "append_coords", "pad_graph", "to_dgl", "preprocess_graph"]


class PygData(SimpleNamespace):
    """Attribute bag standing in for ``torch_geometric.data.Data``: ``x``, ``coords``, ``edge_index``, ``name``,
    ``num_nodes`` (explicit, else inferred from ``x`` like PyG) plus whatever else the file held."""

    def __init__(self, **kw):
        explicit = kw.pop("num_nodes", None)
        super().__init__(**kw)
        self._num_nodes = explicit

    @property
    def num_nodes(self):
        if self._num_nodes is not None:
            return int(self._num_nodes)
        x = getattr(self, "x", None)
        if torch.is_tensor(x):
            return int(x.shape[0])
        ei = getattr(self, "edge_index", None)
        return int(ei.max()) + 1 if torch.is_tensor(ei) and ei.numel() else 0

    @num_nodes.setter
    def num_nodes(self, n):
        self._num_nodes = n

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]


class _Inert:
    """Stand-in for any torch_geometric class met while unpickling: keeps the pickled state, runs no code."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):      # (dict, slots)
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


_stub_cache: Dict[str, type] = {}


def _stub(module: str, name: str) -> type:
    key = f"{module}.{name}"
    if key not in _stub_cache:
        _stub_cache[key] = type(name, (_Inert,), {"__module__": module})
    return _stub_cache[key]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch_geometric" or module.startswith("torch_geometric."):
            return _stub(module, name)
        return super().find_class(module, name)


class _PickleModule:
    """What ``torch.load(pickle_module=...)`` needs: ``Unpickler``, ``load`` and the protocol constants."""
    __name__ = "immunostruct_b200.ingest._PickleModule"
    Unpickler = _Unpickler
    Pickler = pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    DEFAULT_PROTOCOL = pickle.DEFAULT_PROTOCOL
    UnpicklingError = pickle.UnpicklingError
    PickleError = pickle.PickleError

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()

    @staticmethod
    def loads(b, **kw):
        import io
        return _Unpickler(io.BytesIO(b), **kw).load()

    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)


def _mapping_of(obj) -> dict:
    d = getattr(obj, "__dict__", None)
    if isinstance(obj, dict):
        return dict(obj)
    if d is None:
        raise TypeError(f"cannot read a graph out of {type(obj)!r}")
    store = d.get("_store")
    if store is not None:                                   # PyG >= 2.0: Data -> GlobalStorage -> _mapping
        sd = store if isinstance(store, dict) else store.__dict__
        return dict(sd.get("_mapping", sd))
    return {k: v for k, v in d.items() if not k.startswith("_")}     # PyG 1.x


def load_pyg_graph(path: str, map_location="cpu") -> PygData:
    """Read one pickled PyG ``Data`` file -> ``PygData`` (no torch_geometric import)."""
    obj = torch.load(path, map_location=map_location, pickle_module=_PickleModule, weights_only=False)
    m = {k: v for k, v in _mapping_of(obj).items() if v is not None}
    if "edge_index" not in m or "x" not in m:
        raise ValueError(f"{path}: not a graph file (keys: {sorted(m)})")
    return PygData(**m)


def preprocess_graphs(directory: str, files: Iterable[str] = None) -> List[PygData]:
    """data/preprocess.py:15-43: load every ``.pt``, drop names containing 'NXVPMVATV' or 'X', keep the first graph
    per ``name.split("Immuno")[1]``, cut the two h-bond feature columns (x [n,22] -> [n,20])."""
    if files is None:
        files = [f for f in os.listdir(directory) if f.endswith(".pt")]
    graphs = [load_pyg_graph(os.path.join(directory, f)) for f in files]
    graphs = [g for g in graphs if ("NXVPMVATV" not in g.name) and ("X" not in g.name)]
    out, names = [], set()
    for g in graphs:
        key = g.name.split("Immuno")[1]
        if key not in names:
            names.add(key)
            out.append(g)
    for g in out:
        g.x = g.x[:, :-2]
    return out


def pad_graph(graph: PygData, max_nodes: int, feature_size: int, coord_size: int) -> PygData:
    """data/utils.py:13-33: zero-pad node features and coordinates to ``max_nodes`` rows (isolated zero nodes)."""
    add = max_nodes - graph.num_nodes
    if graph.x.shape[1] != feature_size:
        raise ValueError("`pad_graph`: graph.x shape mismatch.")
    if add > 0:
        graph.x = torch.cat([graph.x, torch.zeros(add, feature_size, dtype=graph.x.dtype)], dim=0)
        graph.coords = torch.cat([graph.coords, torch.zeros(add, coord_size, dtype=graph.coords.dtype)], dim=0)
        graph.num_nodes = max_nodes
    return graph


def to_dgl(graph: PygData) -> Graph:
    """data/utils.py:54-67: directed edges exactly as stored, ``edge_attr`` = ones [E,1], ``ndata['x']`` = x."""
    n_edges = graph.edge_index.size(1)
    graph.edge_attr = torch.ones((n_edges, 1))
    src, dst = graph.edge_index
    g = Graph((src, dst), num_nodes=graph.num_nodes)
    g.ndata["x"] = graph.x
    g.edata["edge_attr"] = graph.edge_attr
    return g


def append_coords(graph: PygData) -> PygData:
    """data/preprocess.py:335-338 (and :181,284,291,435): x <- [x | coords] as float32 -- the [n,23] node layout."""
    graph.x = torch.cat([graph.x, graph.coords], dim=-1).to(dtype=torch.float32)
    return graph


def preprocess_graph(graph_mapper: Dict[str, PygData], feature_size: int = 23, coord_size: int = 3) -> Dict[str, Graph]:
    """data/preprocess.py:343-349: pad every graph to the largest node count of the set, then convert."""
    max_nodes = max(g.num_nodes for g in graph_mapper.values())
    padded = {k: pad_graph(g, max_nodes, feature_size, coord_size) for k, g in graph_mapper.items()}
    return {k: to_dgl(g) for k, g in padded.items()}
