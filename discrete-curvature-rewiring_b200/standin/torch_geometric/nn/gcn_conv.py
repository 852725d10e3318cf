"""``GCNConv`` with torch-geometric 2.0.3 semantics: ``out = D^-1/2 (A + I) D^-1/2 X W + b``.

Self-loops are added (weight 1, replacing existing ones is not needed for the simple graphs in scope), the
symmetric normalisation uses the in-degree of the target node computed over ``edge_weight``, the linear layer has
no bias of its own (Glorot initialised), the bias is added after propagation (zeros initialised).  Aggregation is a
plain ``index_add_`` — GCN training is a consumer of the rewired graph, not part of the hot path.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor
from torch.nn import Parameter


class GCNConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.improved = improved
        self.add_self_loops = add_self_loops
        self.normalize = normalize
        self.weight = Parameter(torch.empty(out_channels, in_channels))   # PyG 2.0: self.lin = Linear(.., bias=False)
        self.bias = Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))       # glorot
        torch.nn.init.uniform_(self.weight, -a, a)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Tensor | None = None) -> Tensor:
        n = x.size(0)
        row, col = edge_index[0], edge_index[1]
        if edge_weight is None:
            edge_weight = torch.ones(row.numel(), dtype=x.dtype, device=x.device)
        if self.normalize:
            if self.add_self_loops:
                keep = row != col
                loop = torch.arange(n, device=x.device, dtype=row.dtype)
                fill = 2.0 if self.improved else 1.0
                row = torch.cat([row[keep], loop])
                col = torch.cat([col[keep], loop])
                edge_weight = torch.cat([edge_weight[keep], torch.full((n,), fill, dtype=x.dtype, device=x.device)])
            deg = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(0, col, edge_weight)
            dinv = deg.pow(-0.5)
            dinv[torch.isinf(dinv)] = 0
            edge_weight = dinv[row] * edge_weight * dinv[col]
        h = x @ self.weight.t()
        out = torch.zeros_like(h).index_add_(0, col, h[row] * edge_weight.unsqueeze(-1))
        if self.bias is not None:
            out = out + self.bias
        return out

    def __repr__(self):
        return f"GCNConv({self.in_channels}, {self.out_channels})"
