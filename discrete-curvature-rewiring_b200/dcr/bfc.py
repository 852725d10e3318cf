"""Device-side BFC entry points over a sorted CSR (thin wrappers: allocate outputs, call the C ABI).

PyTorch owns every buffer; the kernels are launched on ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib as L


class DeviceCSR:
    """Symmetric, self-loop-free, sorted CSR resident in HBM: ``rowptr`` int32 ``[n+1]``, ``colidx`` int32."""

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, n: int, max_degree: int | None = None):
        self.rowptr = rowptr
        self.colidx = colidx
        self.n = int(n)
        self.nnz = int(colidx.numel())
        if max_degree is None:
            max_degree = int((rowptr[1:] - rowptr[:-1]).max().item()) if n > 0 else 0
        self.max_degree = int(max_degree)
        self._edges = None
        self._edges_aux = None
        self.asymmetric = False

    @classmethod
    def from_host(cls, rowptr: np.ndarray, colidx: np.ndarray, device="cuda") -> "DeviceCSR":
        L.require_cuda()
        rp = torch.from_numpy(np.ascontiguousarray(rowptr, dtype=np.int32)).to(device)
        ci = torch.from_numpy(np.ascontiguousarray(colidx, dtype=np.int32)).to(device)
        deg = np.diff(np.asarray(rowptr, dtype=np.int64))
        return cls(rp, ci, len(rowptr) - 1, int(deg.max()) if deg.size else 0)

    @classmethod
    def from_dense(cls, A: torch.Tensor, validate: bool = True) -> "DeviceCSR":
        """CSR of a dense fp32 ``[N,N]`` CUDA adjacency (the reference's ``A``).  One host sync (nnz)."""
        L.require_cuda()
        if not (A.is_cuda and A.dim() == 2 and A.shape[0] == A.shape[1]):
            raise ValueError("A must be a square CUDA tensor")
        if A.dtype != torch.float32 or not A.is_contiguous():
            A = A.to(torch.float32).contiguous()
        n = A.shape[0]
        lib = L.load()
        st = L.current_stream()
        counts = torch.empty(max(n, 1), dtype=torch.int32, device=A.device)
        flags = torch.zeros(1, dtype=torch.int32, device=A.device)
        L.check(lib.dcr_dense_count(A.data_ptr(), n, counts.data_ptr(), flags.data_ptr(), st), "dcr_dense_count")
        rowptr = torch.zeros(n + 1, dtype=torch.int32, device=A.device)
        rowptr[1:] = torch.cumsum(counts[:n], 0, dtype=torch.int32)
        host = torch.stack([rowptr[n], flags[0], counts[:n].max() if n else rowptr[n]]).cpu()
        nnz, fl, maxdeg = int(host[0]), int(host[1]), int(host[2])
        if validate == "directed":
            fl &= ~4                 # asymmetry is what the directed route is for
        if validate and fl:
            what = [s for b, s in ((1, "entries other than 0/1"), (2, "a non-zero diagonal (self-loops)"),
                                   (4, "asymmetry (directed graph: use DirectedCSR)")) if fl & b]
            raise NotImplementedError(
                "the B200 BFC kernels cover 0/1 adjacency without self-loops; A has " + ", ".join(what))
        self_asym = bool(int(host[1]) & 4)
        colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=A.device)[:nnz]
        L.check(lib.dcr_dense_fill(A.data_ptr(), n, rowptr.data_ptr(), colidx.data_ptr(), st), "dcr_dense_fill")
        out = cls(rowptr, colidx, n, maxdeg)
        out.asymmetric = self_asym
        return out

    def undirected_edges(self):
        """``(esrc, edst, entry)`` int32/int32/int64 device tensors: entries with row < col, in CSR order."""
        if self._edges is None:
            deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
            rows = torch.repeat_interleave(torch.arange(self.n, device=self.rowptr.device, dtype=torch.int32), deg)
            entry = torch.nonzero(rows < self.colidx, as_tuple=False).flatten()
            self._edges = (rows[entry].contiguous(), self.colidx[entry].contiguous(), entry)
        return self._edges


def edges_aux(csr: DeviceCSR) -> torch.Tensor:
    """Per-graph set-up of the edge-centric cuda-flavour kernels (cached on the CSR): entry -> edge id + hub-hub list."""
    if getattr(csr, "_edges_aux", None) is None:
        lib = L.load()
        esrc, edst, _ = csr.undirected_edges()
        e = int(esrc.numel())
        aux = torch.zeros(int(lib.dcr_bfc_cuda_edges_aux_ints(csr.nnz, e)), dtype=torch.int32, device=csr.colidx.device)
        L.check(lib.dcr_bfc_cuda_edges_prepare(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, esrc.data_ptr(),
                                               edst.data_ptr(), e, csr.nnz, aux.data_ptr(), L.current_stream()),
                "dcr_bfc_cuda_edges_prepare")
        csr._edges_aux = aux
    return csr._edges_aux


def cuda_flavour_edges(csr: DeviceCSR, e_lo: int = 0, count: int | None = None, phases: int = 3, out: dict | None = None,
                       want_fields: bool = True) -> dict:
    """cuda-flavour BFC per UNDIRECTED edge (``esrc < edst``, CSR order): ``tri, sharp, lam`` int32, ``c64``, ``c32``,
    arrays indexed by edge id.  One intersection per edge instead of one per direction; hub-hub edges get a CTA."""
    lib = L.load()
    dev = csr.colidx.device
    esrc, edst, _ = csr.undirected_edges()
    e = int(esrc.numel())
    aux = edges_aux(csr)
    if count is None:
        count = e - e_lo
    if out is None:
        alloc = lambda dt: torch.zeros(max(e, 1), dtype=dt, device=dev)[:e]
        out = {"tri": alloc(torch.int32), "c32": alloc(torch.float32),
               "sharp": alloc(torch.int32) if want_fields else None, "lam": alloc(torch.int32) if want_fields else None,
               "c64": alloc(torch.float64) if want_fields else None}
    L.check(lib.dcr_bfc_cuda_edges(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, esrc.data_ptr(), edst.data_ptr(),
                                   csr.nnz, aux.data_ptr(), int(e_lo), int(count), out["tri"].data_ptr(),
                                   L.ptr(out.get("sharp")), L.ptr(out.get("lam")), L.ptr(out.get("c64")),
                                   out["c32"].data_ptr(), int(phases), L.current_stream()), "dcr_bfc_cuda_edges")
    return out


def support(csr: DeviceCSR, out: torch.Tensor | None = None) -> torch.Tensor:
    """``A2[i,j]`` (#common neighbours) of every directed entry, int32 ``[nnz]``."""
    lib = L.load()
    tri = out if out is not None else torch.empty(max(csr.nnz, 1), dtype=torch.int32, device=csr.colidx.device)[:csr.nnz]
    L.check(lib.dcr_bfc_support(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, tri.data_ptr(), 0, csr.nnz,
                                L.current_stream()), "dcr_bfc_support")
    return tri


def support_tc(csr: DeviceCSR, out: torch.Tensor | None = None, workspace: torch.Tensor | None = None) -> torch.Tensor:
    """Same numbers as :func:`support`, computed as the dense product ``A·A`` on the tensor cores (tcgen05 int8,
    TMA, TMEM) with the edge extraction fused into the epilogue.  Dense regime only (``n <= 32768``)."""
    lib = L.load()
    dev = csr.colidx.device
    tri = out if out is not None else torch.empty(max(csr.nnz, 1), dtype=torch.int32, device=dev)[:csr.nnz]
    nbytes = int(lib.dcr_bfc_support_tc_workspace_bytes(csr.n, csr.nnz))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(lib.dcr_bfc_support_tc(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, csr.nnz, tri.data_ptr(),
                                   workspace.data_ptr(), nbytes, L.current_stream()), "dcr_bfc_support_tc")
    return tri


def cuda_flavour_tc(csr: DeviceCSR, want_fields: bool = True, workspace: torch.Tensor | None = None) -> dict:
    """cuda-flavour BFC per directed entry through the dense tensor-core route (``n <= 32768``): same outputs as
    :func:`cuda_flavour`."""
    lib = L.load()
    dev = csr.colidx.device
    nnz = csr.nnz
    alloc = lambda dt: torch.zeros(max(nnz, 1), dtype=dt, device=dev)[:nnz]
    tri, c32 = alloc(torch.int32), alloc(torch.float32)
    sharp = alloc(torch.int32) if want_fields else None
    lam = alloc(torch.int32) if want_fields else None
    c64 = alloc(torch.float64) if want_fields else None
    nbytes = int(lib.dcr_bfc_cuda_flavour_tc_workspace_bytes(csr.n, nnz))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(lib.dcr_bfc_cuda_flavour_tc(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, nnz, tri.data_ptr(),
                                        L.ptr(sharp), L.ptr(lam), L.ptr(c64), c32.data_ptr(), workspace.data_ptr(),
                                        nbytes, L.current_stream()), "dcr_bfc_cuda_flavour_tc")
    return {"tri": tri, "sharp": sharp, "lam": lam, "c64": c64, "c32": c32, "workspace": workspace}


def cuda_flavour(csr: DeviceCSR, entry_lo: int = 0, entry_hi: int | None = None, want_fields: bool = True,
                 tri: torch.Tensor | None = None) -> dict:
    """cuda-flavour BFC per directed entry: ``tri, sharp, lam`` (int32), ``c64`` (fp64), ``c32`` (fp32).

    The whole graph goes through the edge-centric kernels (every undirected edge once, :func:`cuda_flavour_edges`) and
    is expanded to both entries; an entry sub-range uses the per-entry kernels."""
    lib = L.load()
    dev = csr.colidx.device
    nnz = csr.nnz
    hi = nnz if entry_hi is None else entry_hi
    if entry_lo == 0 and hi == nnz and tri is None and nnz > 0:
        per_edge = cuda_flavour_edges(csr, want_fields=want_fields)
        eid = edges_aux(csr)[:nnz].long()
        return {k: (v[eid] if v is not None else None) for k, v in per_edge.items()}
    if tri is None:
        tri = support(csr)
    alloc = lambda dt: torch.zeros(max(nnz, 1), dtype=dt, device=dev)[:nnz]
    c32 = alloc(torch.float32)
    sharp = alloc(torch.int32) if want_fields else None
    lam = alloc(torch.int32) if want_fields else None
    c64 = alloc(torch.float64) if want_fields else None
    L.check(lib.dcr_bfc_cuda_flavour(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, tri.data_ptr(),
                                     L.ptr(sharp), L.ptr(lam), L.ptr(c64), c32.data_ptr(), entry_lo, hi,
                                     L.current_stream()), "dcr_bfc_cuda_flavour")
    return {"tri": tri, "sharp": sharp, "lam": lam, "c64": c64, "c32": c32}


SMALL_DENSE_MAX_N = 1024
_UNSUPPORTED = ((1, "entries other than 0/1"), (2, "a non-zero diagonal (self-loops)"), (4, "asymmetry (directed graph)"))


def cuda_flavour_dense_small(A: torch.Tensor, C: torch.Tensor) -> torch.Tensor:
    """cuda-flavour BFC of a small dense adjacency (``n <= 1024``) written into the dense ``C`` — two launches, no
    CSR (``dcr_dense_small.cu``).  Raises ``NotImplementedError`` for inputs outside the covered domain."""
    L.require_cuda()
    lib = L.load()
    n = A.shape[0]
    nbytes = int(lib.dcr_bfc_cuda_dense_small_workspace_bytes(n))
    ws = torch.empty(max(nbytes, 4) + 4, dtype=torch.uint8, device=A.device)    # last 4 bytes: the flags word
    flags = ws[-4:].view(torch.int32)
    flags.zero_()
    L.check(lib.dcr_bfc_cuda_dense_small(A.data_ptr(), n, C.data_ptr(), flags.data_ptr(), ws.data_ptr(), nbytes,
                                         L.current_stream()), "dcr_bfc_cuda_dense_small")
    fl = int(flags.item())       # the one host round trip of this path: the domain verdict of the pack kernel
    if fl == 4:
        return None              # asymmetric 0/1 A without self-loops: the caller takes the directed route; C is untouched
    if fl:
        raise NotImplementedError(
            "the B200 BFC kernels cover 0/1 adjacency without self-loops; A has "
            + ", ".join(s for b, s in _UNSUPPORTED if fl & b and b != 4))
    return C


class DirectedCSR:
    """Successor and predecessor CSRs (both sorted) of an asymmetric 0/1 adjacency without self-loops."""

    def __init__(self, out: DeviceCSR, inn: DeviceCSR):
        self.out, self.inn, self.n, self.nnz = out, inn, out.n, out.nnz

    @classmethod
    def from_dense(cls, A: torch.Tensor) -> "DirectedCSR":
        if A.dtype != torch.float32 or not A.is_contiguous():
            A = A.to(torch.float32).contiguous()
        return cls(DeviceCSR.from_dense(A, validate="directed"), DeviceCSR.from_dense(A.t().contiguous(), validate=False))

    @classmethod
    def from_edge_index(cls, edge_index, n: int, device="cuda") -> "DirectedCSR":
        """From a host ``[2, E]`` list of distinct directed edges without self-loops."""
        ei = np.asarray(edge_index, dtype=np.int64)

        def csr(src, dst):
            key = np.unique(src * n + dst)
            rowptr = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.bincount(key // n, minlength=n), out=rowptr[1:])
            return DeviceCSR.from_host(rowptr.astype(np.int32), (key % n).astype(np.int32), device)
        return cls(csr(ei[0], ei[1]), csr(ei[1], ei[0]))


def cuda_flavour_directed(d: DirectedCSR, want_fields: bool = True) -> dict:
    """cuda-flavour BFC of a directed simple graph per successor entry (curvature/bfc_cuda.py:11-48 by the definition)."""
    lib = L.load()
    dev = d.out.colidx.device
    nnz = d.nnz
    alloc = lambda dt: torch.zeros(max(nnz, 1), dtype=dt, device=dev)[:nnz]
    c32 = alloc(torch.float32)
    tri = alloc(torch.int32) if want_fields else None
    sharp = alloc(torch.int32) if want_fields else None
    lam = alloc(torch.int32) if want_fields else None
    c64 = alloc(torch.float64) if want_fields else None
    L.check(lib.dcr_bfc_cuda_flavour_directed(d.out.rowptr.data_ptr(), d.out.colidx.data_ptr(), d.inn.rowptr.data_ptr(),
                                              d.inn.colidx.data_ptr(), d.n, L.ptr(tri), L.ptr(sharp), L.ptr(lam),
                                              L.ptr(c64), c32.data_ptr(), 0, nnz, L.current_stream()),
            "dcr_bfc_cuda_flavour_directed")
    return {"tri": tri, "sharp": sharp, "lam": lam, "c64": c64, "c32": c32}


def post_delta_directed(d: DirectedCSR, x: int, y: int, i_nb: torch.Tensor, j_nb: torch.Tensor,
                        D: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    nb = int(lib.dcr_post_delta_workspace_bytes(d.n, int(i_nb.numel()), int(j_nb.numel())))
    ws = torch.empty(nb, dtype=torch.uint8, device=D.device)
    L.check(lib.dcr_post_delta_directed(d.out.rowptr.data_ptr(), d.out.colidx.data_ptr(), d.inn.rowptr.data_ptr(),
                                        d.inn.colidx.data_ptr(), d.n, int(x), int(y), i_nb.data_ptr(),
                                        int(i_nb.numel()), j_nb.data_ptr(), int(j_nb.numel()), D.data_ptr(),
                                        ws.data_ptr(), nb, L.current_stream()), "dcr_post_delta_directed")
    return D


class PaperWorkspace:
    """Reusable outputs + scratch of the paper-flavour kernel for one (graph, shard) shape.

    The five output arrays are views into ONE contiguous block (``bfc`` f64 first, then the four int32 arrays),
    so a rank's whole result is a single buffer for the all-gather of the multi-GPU path (no packing copy).
    """

    def __init__(self, csr: DeviceCSR, count: int, chunk: int | None = None):
        lib = L.load()
        dev = csr.colidx.device
        self.count = int(count)
        self.chunk = max(int(chunk if chunk is not None else count), 1)
        c = self.chunk
        self.block = torch.zeros(c * 24, dtype=torch.uint8, device=dev)
        self.bfc = self.block[: c * 8].view(torch.float64)
        ints = self.block[c * 8:].view(torch.int32)
        self.tri, self.sq_i, self.sq_j, self.gamma = ints[:c], ints[c:2 * c], ints[2 * c:3 * c], ints[3 * c:4 * c]
        self.scratch_bytes = int(lib.dcr_bfc_paper_scratch_bytes(csr.n, csr.max_degree, max(self.count, 1)))
        self.scratch = torch.empty(max(self.scratch_bytes, 256), dtype=torch.uint8, device=dev)


def shard_count(n_edges: int, rank: int, world: int) -> int:
    """Number of edges ``e = rank + t*world`` below ``n_edges``."""
    return max(0, (n_edges - rank + world - 1) // world)


def paper_flavour(csr: DeviceCSR, rank: int = 0, world: int = 1, ws: PaperWorkspace | None = None,
                  edges=None, events=None) -> dict:
    """Paper-flavour BFC for the undirected edges ``e = rank + t*world`` (compact outputs indexed by ``t``).

    ``edges``: optional ``(esrc, edst)`` int32 device tensors replacing the CSR's own ``row < col`` edge list.
    ``events``: optional pair of recorded-once ``torch.cuda.Event(enable_timing=True)`` bracketing the edge kernels.
    """
    lib = L.load()
    if edges is None:
        esrc, edst, _ = csr.undirected_edges()
    else:
        esrc, edst = edges
    n_edges = int(esrc.numel())
    count = shard_count(n_edges, rank, world)
    if ws is None:
        ws = PaperWorkspace(csr, count)
    ev0 = ev1 = 0
    if events is not None:
        ev0, ev1 = events[0].cuda_event, events[1].cuda_event
    if count > 0:
        L.check(lib.dcr_bfc_paper(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, csr.max_degree,
                                  esrc.data_ptr(), edst.data_ptr(), rank, world, count, ws.tri.data_ptr(),
                                  ws.sq_i.data_ptr(), ws.sq_j.data_ptr(), ws.gamma.data_ptr(), ws.bfc.data_ptr(),
                                  ws.scratch.data_ptr(), ws.scratch_bytes, ev0, ev1, L.current_stream()),
                "dcr_bfc_paper")
    return {"esrc": esrc, "edst": edst, "count": count, "tri": ws.tri[:count], "sq_i": ws.sq_i[:count],
            "sq_j": ws.sq_j[:count], "gamma": ws.gamma[:count], "bfc": ws.bfc[:count], "ws": ws}


def unshard_outputs(n_edges: int, device) -> dict:
    """Pre-allocated full-graph arrays for :func:`unshard` (one allocation per graph, not per pass)."""
    c = max(int(n_edges), 1)
    return {"tri": torch.empty(c, dtype=torch.int32, device=device), "sq_i": torch.empty(c, dtype=torch.int32, device=device),
            "sq_j": torch.empty(c, dtype=torch.int32, device=device), "gamma": torch.empty(c, dtype=torch.int32, device=device),
            "bfc": torch.empty(c, dtype=torch.float64, device=device)}


def unshard(gathered: torch.Tensor, world: int, chunk: int, n_edges: int, out: dict | None = None) -> dict:
    """Full-graph arrays (indexed by edge id) from the all-gathered per-rank blocks of the strided NCCL route."""
    lib = L.load()
    o = out if out is not None else unshard_outputs(n_edges, gathered.device)
    L.check(lib.dcr_bfc_paper_unshard(gathered.data_ptr(), int(world), int(chunk), int(n_edges), o["tri"].data_ptr(),
                                      o["sq_i"].data_ptr(), o["sq_j"].data_ptr(), o["gamma"].data_ptr(), o["bfc"].data_ptr(),
                                      L.current_stream()), "dcr_bfc_paper_unshard")
    return {k: o[k][:n_edges] for k in ("tri", "sq_i", "sq_j", "gamma", "bfc")}


def set_paper_mode(mode: str) -> str:
    """Membership structures of the paper-flavour kernels: ``"auto"`` (exact shared-memory bitmap up to 262144 nodes) or
    ``"hashed"`` (always the hashed bitmap + table of larger graphs — what the parity tests use to cover both paths).
    Returns the previous mode."""
    old = L.load().dcr_bfc_paper_set_mode(1 if mode == "hashed" else 0)
    return "hashed" if old else "auto"


def scatter_dense(csr: DeviceCSR, vals: torch.Tensor, C: torch.Tensor) -> torch.Tensor:
    """Write the legacy dense ``[N,N]`` fp32 matrix (zeros off the entries) in place."""
    lib = L.load()
    L.check(lib.dcr_scatter_dense(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, vals.data_ptr(),
                                  C.data_ptr(), L.current_stream()), "dcr_scatter_dense")
    return C


def post_delta(csr: DeviceCSR, tri: torch.Tensor, x: int, y: int, i_nb: torch.Tensor, j_nb: torch.Tensor,
               D: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    nb = int(lib.dcr_post_delta_workspace_bytes(csr.n, int(i_nb.numel()), int(j_nb.numel())))
    ws = torch.empty(nb, dtype=torch.uint8, device=D.device)
    L.check(lib.dcr_post_delta(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, tri.data_ptr(), int(x), int(y),
                               i_nb.data_ptr(), int(i_nb.numel()), j_nb.data_ptr(), int(j_nb.numel()), D.data_ptr(),
                               ws.data_ptr(), nb, L.current_stream()), "dcr_post_delta")
    return D
