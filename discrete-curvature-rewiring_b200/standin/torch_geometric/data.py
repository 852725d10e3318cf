"""``torch_geometric.data.Data`` / ``InMemoryDataset`` stand-ins (attribute bag + ``num_nodes`` inference)."""
from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self.pos = pos
        self._num_nodes = None
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if self._num_nodes is not None:
            return self._num_nodes
        if self.x is not None:
            return int(self.x.size(0))
        if self.pos is not None:
            return int(self.pos.size(0))
        if self.edge_index is not None and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return 0

    @num_nodes.setter
    def num_nodes(self, n):
        self._num_nodes = n

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else int(self.edge_index.size(1))

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if not k.startswith("_") and v is not None]

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    def to(self, device):
        for k in self.keys:
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def __repr__(self):
        parts = []
        for k in self.keys:
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v}")
        if self._num_nodes is not None:
            parts.append(f"num_nodes={self._num_nodes}")
        return f"Data({', '.join(parts)})"


class InMemoryDataset:
    """Just enough for ``models/gcn.py:13-16``: ``dataset.data`` and ``dataset.num_classes``."""

    def __init__(self, data: Data | None = None, num_classes: int | None = None):
        self.data = data
        self._num_classes = num_classes

    @property
    def num_classes(self):
        if self._num_classes is not None:
            return self._num_classes
        return int(self.data.y.max()) + 1

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        return self.data
