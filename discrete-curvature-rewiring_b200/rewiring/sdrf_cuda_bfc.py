"""Drop-in for the reference's ``rewiring/sdrf_cuda_bfc.py`` (rewiring/sdrf_cuda_bfc.py:14-93).

Same signature and return type.  The whole loop — curvature, argmin/argmax, candidate scoring, softmax draw, edge
insertion/removal, incremental curvature refresh — runs inside one persistent sm_100a kernel of ``libdcr.so`` on a
device-resident adjacency; the host only sets the graph up (networkx insertion order, PyG canonicalisation) and
rebuilds the ``Data`` object.  Random numbers: one double from ``np.random``'s global generator per iteration that
has candidates, exactly as the reference consumes them (pass ``uniforms=`` to supply them explicitly).
``is_undirected=False`` runs the same loop on the ``DiGraph`` of the input (successors of ``x`` x predecessors of ``y``,
single directed entries added / removed, :47-49, :72-73, :87-88) with the definitional curvature of an asymmetric
adjacency (csrc/dcr_directed.cuh).
"""
import torch

from dcr import compat as _compat
from dcr import sdrf as _sdrf

_compat.ensure_torch_geometric()
import torch_geometric  # noqa: E402
from torch_geometric.data import Data  # noqa: E402


def sdrf_cuda_bfc(data: "torch_geometric.data.Data", loops: int, remove_edges: bool,
                  removal_bound: float, tau: int, is_undirected: bool, *, uniforms=None,
                  return_log: bool = False) -> "torch_geometric.data.Data":
    """
    Perform SDRF graph rewiring with Balanced Forman curvature on the GPU.
    :param data: data to be rewired (``edge_index`` int64 ``[2, E]``, ``num_nodes``).
    :param loops: number of edge addition/deletion iterations.
    :param remove_edges: whether to delete highly curved edges each iteration to compensate for the addition.
    :param removal_bound: curvature lower bound of deleting edges (delete edges only with higher curvature).
    :param tau: softmax temperature for choosing the edge to add; ``float('inf')`` picks the maximum.
    :param is_undirected: flag specifying whether the data is undirected.
    :return: rewired data (``edge_index`` in the order ``from_networkx`` yields, ``num_nodes``).
    """
    edge_index = data.edge_index
    num_nodes = int(data.num_nodes)
    res = _sdrf.sdrf(edge_index, num_nodes, int(loops), bool(remove_edges), float(removal_bound), tau,
                     uniforms=uniforms, return_log=return_log, is_undirected=bool(is_undirected))
    ei, log = res if return_log else (res, None)
    out = Data(edge_index=torch.from_numpy(ei).to(torch.long))
    out.num_nodes = max(num_nodes, int(ei.max()) + 1 if ei.size else 0)
    if return_log:
        out.sdrf_log = log
    return out
