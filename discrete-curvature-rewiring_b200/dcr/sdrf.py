"""Host driver of the device-resident SDRF loop (``dcr_sdrf_*`` in include/dcr.h).

Mirrors the control flow of ``rewiring/sdrf_cuda_bfc.py:14-93`` of the reference: set the graph up on the host
(once), run all iterations inside one persistent kernel, come back to the host only when the kernel asks for a
host-side decision (a uniform within ``guard`` of a softmax CDF boundary) or reports one of the
reference's exceptions.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import graph as G
from . import lib as L

# numpy's own messages for the ValueErrors np.random.choice raises (sdrf_cuda_bfc.py:64-68)
_MSG_NAN = "probabilities contain NaN"
_MSG_SUM = "probabilities do not sum to 1"


def host_choice(improvements: np.ndarray, tau, u: float) -> int:
    """The reference's host-side selection for one iteration: ``utils.softmax.softmax`` + the algorithm of
    ``np.random.choice(range(n), p=p)`` for the uniform ``u`` it would draw (App. E.3)."""
    a = np.asarray(improvements, dtype=np.float64)
    if tau == float("inf"):
        p = np.zeros(len(a))
        p[np.argmax(a)] = 1
    else:
        e = np.exp(a * tau)
        p = e / e.sum()
    if np.isnan(p).any():
        raise ValueError(_MSG_NAN)
    if abs(p.sum() - 1.0) > np.sqrt(np.finfo(np.float64).eps):
        raise ValueError(_MSG_SUM)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


class SdrfState:
    """Owns one ``dcr_sdrf`` handle (device-resident dynamic adjacency + incremental curvature)."""

    def __init__(self, rowptr_order: np.ndarray, order: np.ndarray, max_additions: int, mode: int = L.SDRF_MODE_BFC,
                 in_rowptr: np.ndarray | None = None, in_order: np.ndarray | None = None):
        """``mode``: one of ``L.SDRF_MODE_*``.  The directed mode takes the successor lists as ``rowptr_order/order`` and
        the predecessor lists as ``in_rowptr/in_order`` (all in networkx insertion order)."""
        L.require_cuda()
        self.lib = L.load()
        self.mode = int(mode)
        self.n = int(len(rowptr_order) - 1)
        rp = np.ascontiguousarray(rowptr_order, dtype=np.int32)
        od = np.ascontiguousarray(order, dtype=np.int32)
        handle = C.c_void_p()
        if self.mode == L.SDRF_MODE_BFC_DIRECTED:
            irp = np.ascontiguousarray(in_rowptr, dtype=np.int32)
            iod = np.ascontiguousarray(in_order, dtype=np.int32)
        else:
            irp = iod = None
        self._keep = (rp, od, irp, iod)
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(rp.astype(np.int64)))
        self.n_self = int(np.count_nonzero(od[: rows.size] == rows))      # nodes that list themselves (self-loops of G)
        L.check(self.lib.dcr_sdrf_create_mode(self.n, self.mode, rp.ctypes.data, od.ctypes.data if od.size else 0,
                                              irp.ctypes.data if irp is not None else 0,
                                              iod.ctypes.data if iod is not None and iod.size else 0,
                                              int(max_additions), C.byref(handle)), "dcr_sdrf_create_mode")
        self.handle = handle
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._result = torch.zeros(8, dtype=torch.int32, device=self.device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dcr_sdrf_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, loops: int, remove_edges: bool, removal_bound: float, tau, uniforms: torch.Tensor,
            forced_choice: int = -1, guard: float = 1e-9, log: torch.Tensor | None = None):
        """One kernel launch, up to ``loops`` iterations.  Returns ``(result dict, log int32[iters, 8])`` (device log
        sliced after one D2H of the 32-byte result)."""
        if log is None:
            log = torch.empty((max(loops, 1), L.SDRF_LOG_INTS), dtype=torch.int32, device=self.device)
        L.check(self.lib.dcr_sdrf_run(self.handle, int(loops), int(bool(remove_edges)), float(removal_bound),
                                      float(tau), uniforms.data_ptr(), int(uniforms.numel()), int(forced_choice),
                                      float(guard), log.data_ptr(), self._result.data_ptr(), L.current_stream()),
                "dcr_sdrf_run")
        r = self._result.cpu().tolist()
        res = {"status": r[0], "iterations_done": r[1], "draws_used": r[2], "stopped": r[3], "pending_n": r[4],
               "pending_x": r[5], "pending_y": r[6]}
        return res, log[: res["iterations_done"]]

    def pending_improvements(self, n: int) -> np.ndarray:
        out = torch.empty(max(n, 1), dtype=torch.float64, device=self.device)
        L.check(self.lib.dcr_sdrf_pending_improvements(self.handle, out.data_ptr(), int(n), L.current_stream()),
                "dcr_sdrf_pending_improvements")
        return out[:n].cpu().numpy()

    def nnz(self) -> int:
        return int(self.lib.dcr_sdrf_nnz(self.handle))

    def export(self, with_curvature: bool = False):
        """Current graph: ``(rowptr, order)`` numpy (networkx order); with curvature also sorted col / c32 / tri."""
        nnz = self.nnz()
        dev = self.device
        rowptr = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
        if self.n_self and not with_curvature:
            # nodes that list themselves (self-loops of G): the insertion-order rows are longer than the adjacency rows
            order = torch.empty(max(nnz + self.n_self, 1), dtype=torch.int32, device=dev)
            L.check(self.lib.dcr_sdrf_export_order(self.handle, rowptr.data_ptr(), order.data_ptr(), L.current_stream()),
                    "dcr_sdrf_export_order")
            rp = rowptr.cpu().numpy()
            return rp, order[: int(rp[-1])].cpu().numpy()
        order = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        col = c32 = tri = None
        if with_curvature:
            col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
            c32 = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
            tri = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        L.check(self.lib.dcr_sdrf_export(self.handle, rowptr.data_ptr(), 0 if self.n_self else order.data_ptr(), L.ptr(col),
                                         L.ptr(c32), L.ptr(tri), L.current_stream()), "dcr_sdrf_export")
        out = (rowptr.cpu().numpy(), order[:nnz].cpu().numpy() if not self.n_self else None)
        if with_curvature:
            out += (col[:nnz].cpu().numpy(), c32[:nnz].cpu().numpy(), tri[:nnz].cpu().numpy())
        return out


def _raise_for_status(status: int):
    if status == L.SDRF_PROB_NAN:
        raise ValueError(_MSG_NAN)
    if status == L.SDRF_PROB_SUM:
        raise ValueError(_MSG_SUM)
    if status == L.SDRF_REMOVE_NONEDGE:
        try:
            import networkx as nx
            raise nx.NetworkXError("The edge 0-0 is not in the graph")
        except ImportError:
            raise KeyError("The edge 0-0 is not in the graph")
    if status == L.SDRF_NO_UNIFORM:
        raise L.DcrError("ran out of uniforms (one is consumed per iteration that has candidates)")
    if status == L.SDRF_ARENA_FULL:
        raise L.DcrError("SDRF adjacency arena exhausted")
    if status == L.SDRF_EMPTY_GRAPH:
        raise ValueError("min() arg is an empty sequence")      # sdrf_no_cuda.py:26 on a graph without edges
    if status == L.SDRF_TOO_MANY_CANDIDATES:
        raise L.DcrError("candidate matrix (deg x + 1)(deg y + 1) exceeds the scratch of the SDRF state")
    raise L.DcrError(f"unexpected SDRF status {status}")


CLASSICAL_MODES = {"1d": L.SDRF_MODE_1D, "augmented": L.SDRF_MODE_AUGMENTED, "haantjes": L.SDRF_MODE_HAANTJES}


def sdrf(edge_index, num_nodes: int, loops: int, remove_edges: bool, removal_bound: float, tau,
         uniforms: np.ndarray | None = None, guard: float = 1e-9, return_log: bool = False,
         state_out: list | None = None, is_undirected: bool = True, curv_type: str = "bfc"):
    """SDRF on the GPU.  Returns ``edge_index`` (int64 numpy) in the column order ``from_networkx`` produces, and with
    ``return_log`` the int32 ``[iters, 8]`` per-iteration log.

    ``curv_type='bfc'``: ``sdrf_cuda_bfc`` (rewiring/sdrf_cuda_bfc.py:14-93), undirected or — ``is_undirected=False`` —
    on the ``DiGraph`` of the input with single directed entries added / removed (:47-49, :72-73, :87-88).
    ``curv_type`` in ``'1d' | 'augmented' | 'haantjes'``: ``sdrf_no_cuda`` (rewiring/sdrf_no_cuda.py:9-68).

    ``uniforms``: the doubles ``np.random`` would produce (one per iteration that has candidates).  When omitted
    they are taken from numpy's global legacy generator exactly as the reference consumes it: the global state
    ends up advanced by the number of draws actually used.
    """
    L.require_cuda()
    directed = False
    if curv_type == "bfc":
        if is_undirected:
            rowptr, order = G.networkx_order(edge_index, num_nodes, keep_self_loops=True)    # G keeps them (:31), A does not (:29)
            state = SdrfState(rowptr, order, max_additions=max(int(loops), 0))
        else:
            directed = True
            s_rp, s_ord, p_rp, p_ord = G.digraph_order(edge_index, num_nodes, keep_self_loops=True)
            state = SdrfState(s_rp, s_ord, max_additions=max(int(loops), 0), mode=L.SDRF_MODE_BFC_DIRECTED,
                              in_rowptr=p_rp, in_order=p_ord)
    elif curv_type in CLASSICAL_MODES:
        rowptr, order = G.classical_order(edge_index, num_nodes, keep_self_loops=True)
        state = SdrfState(rowptr, order, max_additions=max(int(loops), 0), mode=CLASSICAL_MODES[curv_type])
    else:
        raise Exception(f"Method {curv_type} not available.")    # classical_curvatures.py:27-28
    own_stream = uniforms is None
    if own_stream:
        rng_state = np.random.get_state()
        uniforms = np.random.random_sample(max(int(loops), 0))
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    if uniforms.size < max(loops, 0):
        uniforms = np.concatenate([uniforms, np.full(loops - uniforms.size, np.nan)])  # NaN = "not supplied"
    n_supplied = int(np.count_nonzero(~np.isnan(uniforms)))
    u_dev = torch.from_numpy(uniforms).to(state.device)
    logs = []
    done = draws = 0
    forced = -1
    try:
        while done < loops:
            res, log = state.run(loops - done, remove_edges, removal_bound, tau, u_dev[draws:n_supplied] if n_supplied > draws
                                 else u_dev[:0], forced_choice=forced, guard=guard)
            forced = -1
            done += res["iterations_done"]
            draws += res["draws_used"]
            if res["iterations_done"]:
                logs.append(log.cpu().numpy())
            if res["status"] == L.SDRF_OK:
                break
            if res["status"] == L.SDRF_NEED_HOST:
                imp = state.pending_improvements(res["pending_n"])
                forced = host_choice(imp, tau, float(uniforms[draws]))
                continue
            _raise_for_status(res["status"])
        out_rowptr, out_order = state.export()
    finally:
        if own_stream:   # advance numpy's global generator by exactly the draws the reference would have made
            np.random.set_state(rng_state)
            if draws:
                np.random.random_sample(draws)
        if state_out is not None:
            state_out.append(state)
        else:
            state.close()
    edge_index_out = (G.from_digraph_order if directed else G.from_networkx_order)(out_rowptr, out_order)
    if return_log:
        return edge_index_out, (np.concatenate(logs) if logs else np.zeros((0, L.SDRF_LOG_INTS), dtype=np.int32))
    return edge_index_out
