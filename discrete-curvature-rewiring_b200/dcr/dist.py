"""Multi-GPU full-graph paper-flavour BFC: graph replicated, edges cut into contiguous ranges of equal WORK, results
all-gathered over NVLink (north_star: "shards naturally by edge range ... curvature shards all-gathered").

Every edge is independent given the graph.  Rank ``r`` of ``W`` computes the edges ``[bounds[r], bounds[r+1])`` — the cut
points balance the per-edge work estimate of ``dcr_bfc_paper_edge_cost`` (entries streamed + per-head + per-edge terms),
not the edge count: a hub's edges are consecutive in CSR order and cost two orders of magnitude more than the median.

Two routes for the exchange:

* ``peer`` (default on GPUs): every rank owns the full result arrays in a buffer its peers map through CUDA IPC
  (``dcr_comm``); the closing kernel of a rank stores each result at the edge's position in every rank's buffer, ordered
  by device-side flags — compute and all-gather are ONE kernel over NVLink peer memory, with no staging block, no
  collective call and no re-interleave pass.
* ``nccl``: the round-1 route, kept as the baseline and as the fallback when peer mapping is unavailable: strided shards
  ``e = r + t*W``, one ``all_gather_into_tensor`` of the compact per-rank blocks, then ``dcr_bfc_paper_unshard``.

``run_host`` is the end-to-end form (host CSR in, host results out): each rank uploads the graph and ITS slice of the edge
list, computes its range and copies only its own slice of the results into one host buffer shared by the ranks
(POSIX shared memory, page-locked by every rank) — when the consumer is the host no device-side gather is needed at all.

The SDRF loop does not shard (iteration t+1 needs the graph of iteration t): replicas only.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import numpy as np
import torch
import torch.distributed as dist

from . import bfc
from . import lib as L

FIELDS = ("bfc", "tri", "sq_i", "sq_j", "gamma")


def chunk_size(n_edges: int, world: int) -> int:
    return max(1, (n_edges + world - 1) // world)


def balanced_bounds(cost_prefix, world: int) -> list[int]:
    """Cut points ``b[0]=0 <= ... <= b[world]=E`` of a contiguous partition with equal work: ``cost_prefix`` is the
    INCLUSIVE prefix sum of the per-edge work estimates (numpy int64 array or torch tensor).  Deterministic integer
    arithmetic: every rank computes the same cuts."""
    n = int(cost_prefix.shape[0])
    if n == 0:
        return [0] * (world + 1)
    if torch.is_tensor(cost_prefix):
        total = int(cost_prefix[-1].item())
        targets = torch.tensor([(total * k) // world for k in range(1, world)], dtype=cost_prefix.dtype,
                               device=cost_prefix.device)
        cuts = torch.searchsorted(cost_prefix, targets, right=False).cpu().tolist() if world > 1 else []
    else:
        total = int(cost_prefix[-1])
        targets = np.array([(total * k) // world for k in range(1, world)], dtype=np.int64)
        cuts = np.searchsorted(cost_prefix, targets, side="left").tolist() if world > 1 else []
    b = [0] + [min(n, int(c) + 1) for c in cuts] + [n]
    for k in range(1, len(b)):
        b[k] = max(b[k], b[k - 1])
    return b


def edge_cost(csr: "bfc.DeviceCSR", esrc: torch.Tensor, edst: torch.Tensor) -> torch.Tensor:
    """Per-edge work estimate (int64, device)."""
    lib = L.load()
    dev = csr.colidx.device
    e = int(esrc.numel())
    cost = torch.empty(max(e, 1), dtype=torch.int64, device=dev)[:e]
    scratch = torch.empty(max(csr.n, 1), dtype=torch.int64, device=dev)
    L.check(lib.dcr_bfc_paper_edge_cost(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, esrc.data_ptr(),
                                        edst.data_ptr(), e, cost.data_ptr(), scratch.data_ptr(), L.current_stream()),
            "dcr_bfc_paper_edge_cost")
    return cost


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class PeerComm:
    """The ``dcr_comm`` of this rank: full result arrays in IPC-mapped device memory + views of them as torch tensors."""

    def __init__(self, n_edges: int, rank: int, world: int, group=None, device=None):
        L.require_cuda()
        lib = L.load()
        self.lib = lib
        self.rank, self.world, self.n_edges = rank, world, int(n_edges)
        h = C.c_void_p()
        L.check(lib.dcr_comm_create(rank, world, self.n_edges, C.byref(h)), "dcr_comm_create")
        self.handle = h
        if world > 1:
            mine = (C.c_ubyte * 64)()
            L.check(lib.dcr_comm_handle(self.handle, C.cast(mine, C.c_void_p)), "dcr_comm_handle")
            gathered = [None] * world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(gathered))
            L.check(lib.dcr_comm_connect(self.handle, C.cast(blob, C.c_void_p)), "dcr_comm_connect")
        self.chunk = int(lib.dcr_comm_chunk(self.handle))
        ptr = int(lib.dcr_comm_buffer(self.handle))
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.views = _views_of_device_block(ptr, self.chunk, self.n_edges, dev)

    def error(self) -> int:
        return int(self.lib.dcr_comm_error(self.handle))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.views = None
            self.lib.dcr_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _views_of_device_block(ptr: int, chunk: int, n_edges: int, dev) -> dict:
    """torch views (no copy, no ownership) of bfc f64[chunk] | tri | sq_i | sq_j | gamma int32[chunk] at ``ptr``."""

    class _Mem:       # minimal __cuda_array_interface__ carrier
        def __init__(self, p, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (p, False), "version": 2}

    block = torch.as_tensor(_Mem(ptr, chunk * 24), device=dev)
    f = block[: chunk * 8].view(torch.float64)
    ints = block[chunk * 8:].view(torch.int32)
    e = n_edges
    return {"bfc": f[:e], "tri": ints[:e], "sq_i": ints[chunk:chunk + e], "sq_j": ints[2 * chunk:2 * chunk + e],
            "gamma": ints[3 * chunk:3 * chunk + e], "_block": block}


class ShardedPaperBFC:
    """Reusable buffers for repeated sharded passes over one graph.

    ``mode``: ``"peer"`` (fused compute + all-gather over NVLink peer memory, contiguous work-balanced ranges),
    ``"nccl"`` (strided shards + NCCL all-gather + re-interleave) or ``"auto"`` = peer, falling back to nccl when the
    peers' buffers cannot be mapped (the fallback is agreed on by all ranks and reported through ``self.mode``)."""

    def __init__(self, csr: "bfc.DeviceCSR", group=None, mode: str = "auto"):
        self.csr = csr
        self.group = group
        self.rank, self.world = _world(group)
        esrc, edst, _ = csr.undirected_edges()
        self.esrc, self.edst = esrc, edst
        self.n_edges = int(esrc.numel())
        self.comm = None
        self.mode = mode
        if mode in ("auto", "peer"):
            ok, why = 1, ""
            try:
                self.comm = PeerComm(self.n_edges, self.rank, self.world, group, device=csr.colidx.device)
            except Exception as exc:          # no peer mapping on this box: every rank must take the same route
                ok, why = 0, repr(exc)[:200]
            if self.world > 1:
                flag = torch.tensor([ok], dtype=torch.int32, device=csr.colidx.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                ok = int(flag.item())
            if ok:
                self.mode = "peer"
            else:
                if mode == "peer":
                    raise L.DcrError(f"peer-memory exchange unavailable: {why}")
                if self.comm is not None:
                    self.comm.close()
                    self.comm = None
                warnings.warn(f"dcr.dist: peer-memory exchange unavailable ({why or 'on another rank'}); using the NCCL route")
                self.mode = "nccl"
        if self.mode == "peer":
            cost = edge_cost(csr, esrc, edst)
            self.bounds = balanced_bounds(torch.cumsum(cost, 0), self.world)
            self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
            self.count = self.hi - self.lo
            lib = L.load()
            self.scratch_bytes = int(lib.dcr_bfc_paper_scratch_bytes(csr.n, csr.max_degree, max(self.count, 1)))
            self.scratch = torch.empty(max(self.scratch_bytes, 256), dtype=torch.uint8, device=csr.colidx.device)
        elif self.mode == "nccl":
            self.chunk = chunk_size(self.n_edges, self.world)
            self.count = bfc.shard_count(self.n_edges, self.rank, self.world)
            self.ws = bfc.PaperWorkspace(csr, self.count, chunk=self.chunk)
            self.gathered = torch.empty(self.world * self.chunk * 24, dtype=torch.uint8, device=csr.colidx.device)
            self.out = bfc.unshard_outputs(self.n_edges, csr.colidx.device)
        else:
            raise ValueError(f"unknown mode {mode!r}")

    # ---- device-resident pass --------------------------------------------------------------------------------
    def run(self, events=None) -> dict:
        """One pass; returns the full-graph arrays (views into buffers owned by this object, overwritten by the next
        pass).  Only enqueues work on the current stream."""
        if self.mode == "nccl":
            bfc.paper_flavour(self.csr, rank=self.rank, world=self.world, ws=self.ws, events=events)
            if self.world == 1:
                ws, e = self.ws, self.n_edges
                return {"tri": ws.tri[:e], "sq_i": ws.sq_i[:e], "sq_j": ws.sq_j[:e], "gamma": ws.gamma[:e], "bfc": ws.bfc[:e]}
            dist.all_gather_into_tensor(self.gathered, self.ws.block, group=self.group)
            return bfc.unshard(self.gathered, self.world, self.chunk, self.n_edges, out=self.out)
        return self._run_peer(self.csr, self.esrc, self.edst, events)

    def _run_peer(self, csr, esrc, edst, events=None):
        lib = L.load()
        ev0 = ev1 = 0
        if events is not None:
            ev0, ev1 = events[0].cuda_event, events[1].cuda_event
        L.check(lib.dcr_bfc_paper_sharded(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, csr.max_degree,
                                          esrc.data_ptr(), edst.data_ptr(), self.lo, self.count, self.comm.handle,
                                          self.scratch.data_ptr(), self.scratch_bytes, ev0, ev1, L.current_stream()),
                "dcr_bfc_paper_sharded")
        v = self.comm.views
        return {k: v[k] for k in FIELDS}

    def check(self):
        """Synchronises; raises if a peer hand-shake timed out during the passes so far."""
        if self.comm is not None and self.comm.error():
            raise L.DcrError("dcr.dist: a peer rank never arrived at a pass (hand-shake timed out)")

    def close(self):
        if self.comm is not None:
            self.comm.close()
            self.comm = None


CUDA_FIELDS = ("c64", "tri", "sharp", "lam", "c32")


class ShardedCudaBFC:
    """Multi-GPU full-graph CUDA-flavour BFC (``balanced_forman_curvature`` of
    curvature/bfc_cuda.py:51-65 at full-graph scale): graph replicated, contiguous edge ranges balanced by the
    intersection work ``min(d_i, d_j) * log2(max(d_i, d_j))``, two passes with the all-gather of the supports between
    them fused into the kernels over CUDA-IPC peer memory (``dcr_bfc_cuda_sharded``).  ``run()`` returns the full-graph
    per-edge arrays ``c64, tri, sharp, lam, c32`` (views of the rank's ``dcr_comm`` buffer)."""

    def __init__(self, csr: "bfc.DeviceCSR", group=None):
        self.csr = csr
        self.group = group
        self.rank, self.world = _world(group)
        esrc, edst, _ = csr.undirected_edges()
        self.esrc, self.edst = esrc, edst
        self.n_edges = int(esrc.numel())
        self.aux = bfc.edges_aux(csr)
        self.comm = PeerComm(self.n_edges, self.rank, self.world, group, device=csr.colidx.device)
        deg = (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64)
        di, dj = deg[esrc.long()], deg[edst.long()]
        cost = torch.minimum(di, dj) * (torch.log2(torch.maximum(di, dj).double() + 1).ceil().to(torch.int64) + 1) + 16
        self.bounds = balanced_bounds(torch.cumsum(cost, 0), self.world)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        v = self.comm.views
        self.views = {"c64": v["bfc"], "tri": v["tri"], "sharp": v["sq_i"], "lam": v["sq_j"],
                      "c32": v["gamma"].view(torch.float32)}

    def run(self) -> dict:
        """One pass over this rank's range; only enqueues work on the current stream."""
        lib = L.load()
        csr = self.csr
        L.check(lib.dcr_bfc_cuda_sharded(csr.rowptr.data_ptr(), csr.colidx.data_ptr(), csr.n, self.esrc.data_ptr(),
                                         self.edst.data_ptr(), csr.nnz, self.aux.data_ptr(), self.lo, self.hi - self.lo,
                                         self.comm.handle, L.current_stream()), "dcr_bfc_cuda_sharded")
        return self.views

    def check(self):
        if self.comm is not None and self.comm.error():
            raise L.DcrError("dcr.dist: a peer rank never arrived at a pass (hand-shake timed out)")

    def close(self):
        if self.comm is not None:
            self.views = None
            self.comm.close()
            self.comm = None


class HostShardedPaperBFC:
    """End to end on host buffers: pinned host CSR (+ optionally the edge list) in, one host result block (shared by the
    ranks) out.

    Per pass and rank: H2D of ``rowptr``/``colidx`` (the graph is replicated; the undirected edge list is derived from it
    on the device unless the caller uploads one),
    the work-balanced range computed from the uploaded graph on the device (one 8-byte read-back of the rank's own
    bounds), the pass over that range, D2H of the rank's slice of the five result arrays into the shared block.  No
    device-side gather: the consumer is the host.  Layout of the block: bfc f64[E] | tri | sq_i | sq_j | gamma int32[E].
    """

    def __init__(self, n: int, nnz: int, n_edges: int, max_degree: int, group=None, device=None, shm_name=None):
        L.require_cuda()
        self.group = group
        self.rank, self.world = _world(group)
        self.n, self.nnz, self.n_edges, self.max_degree = int(n), int(nnz), int(n_edges), int(max_degree)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dev = dev
        self.d_rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.d_col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
        self.d_esrc = torch.empty(max(n_edges, 1), dtype=torch.int32, device=dev)[:n_edges]
        self.d_edst = torch.empty(max(n_edges, 1), dtype=torch.int32, device=dev)[:n_edges]
        self.csr = bfc.DeviceCSR(self.d_rowptr, self.d_col, n, max_degree)
        self.comm = PeerComm(n_edges, 0, 1, device=dev)           # local result arrays only (world 1: no peers mapped)
        lib = L.load()
        # the range is not known before the graph arrives: scratch for the whole edge list
        self.scratch_bytes = int(lib.dcr_bfc_paper_scratch_bytes(n, max_degree, max(n_edges, 1)))
        self.scratch = torch.empty(max(self.scratch_bytes, 256), dtype=torch.uint8, device=dev)
        self.cost = torch.empty(max(n_edges, 1), dtype=torch.int64, device=dev)[:n_edges]
        self.node_s = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        self.h_bounds = torch.empty(2, dtype=torch.int64).pin_memory()
        self.host_block = self._shared_block(shm_name)
        self.bounds = None

    def _shared_block(self, shm_name):
        nbytes = max(self.n_edges, 1) * 24
        if self.world == 1:
            return torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        name = [shm_name or f"/dev/shm/dcr_bfc_{os.getpid()}_{id(self):x}"]
        dist.broadcast_object_list(name, src=0, group=self.group)
        path = name[0]
        if self.rank == 0:
            with open(path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier(group=self.group)
        t = torch.from_file(path, shared=True, size=nbytes, dtype=torch.uint8)
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
        if int(rc) != 0:
            raise L.DcrError(f"cudaHostRegister of the shared result block failed ({rc})")
        dist.barrier(group=self.group)
        if self.rank == 0:
            os.unlink(path)              # the mappings keep it alive
        self._registered = (t.data_ptr(), nbytes)
        return t

    def host_views(self) -> dict:
        e = self.n_edges
        b = self.host_block
        ints = b[e * 8:].view(torch.int32)
        return {"bfc": b[: e * 8].view(torch.float64), "tri": ints[:e], "sq_i": ints[e:2 * e], "sq_j": ints[2 * e:3 * e],
                "gamma": ints[3 * e:4 * e]}

    def run(self, h_rowptr, h_col, h_esrc=None, h_edst=None):
        """One end-to-end pass (pinned host tensors in).  Enqueues on the current stream, with one small blocking
        read-back (the rank's bounds) in the middle; the rank's slice of ``host_block`` is complete when the stream is.
        Without ``h_esrc / h_edst`` the undirected edge list (row < col entries, CSR order) is derived on the device from
        the uploaded CSR — half the host-to-device bytes."""
        lib = L.load()
        st = L.current_stream()
        E, W = self.n_edges, self.world
        self.d_rowptr.copy_(h_rowptr, non_blocking=True)
        self.d_col.copy_(h_col, non_blocking=True)
        # every rank needs the whole edge list once to find its cut points (the cost of an edge needs both endpoints)
        self._uploaded_edges = h_esrc is not None
        if h_esrc is not None:
            self.d_esrc.copy_(h_esrc, non_blocking=True)
            self.d_edst.copy_(h_edst, non_blocking=True)
        else:
            L.check(lib.dcr_csr_upper_count(self.d_rowptr.data_ptr(), self.d_col.data_ptr(), self.n, self.node_s.data_ptr(),
                                            st), "dcr_csr_upper_count")
            self.node_s[: self.n].cumsum_(0)                       # (node_s is free until the cost kernels)
            L.check(lib.dcr_csr_upper_fill(self.d_rowptr.data_ptr(), self.d_col.data_ptr(), self.n, self.node_s.data_ptr(),
                                           self.d_esrc.data_ptr(), self.d_edst.data_ptr(), st), "dcr_csr_upper_fill")
        if W > 1:
            L.check(lib.dcr_bfc_paper_edge_cost(self.d_rowptr.data_ptr(), self.d_col.data_ptr(), self.n,
                                                self.d_esrc.data_ptr(), self.d_edst.data_ptr(), E, self.cost.data_ptr(),
                                                self.node_s.data_ptr(), st), "dcr_bfc_paper_edge_cost")
            pre = torch.cumsum(self.cost, 0)
            total = pre[-1]
            tg = torch.stack([(total * self.rank) // W, (total * (self.rank + 1)) // W])
            cut = torch.searchsorted(pre, tg, right=False) + 1
            self.h_bounds.copy_(cut, non_blocking=False)
            lo = 0 if self.rank == 0 else min(E, int(self.h_bounds[0]))
            hi = E if self.rank == W - 1 else min(E, int(self.h_bounds[1]))
            hi = max(hi, lo)
        else:
            lo, hi = 0, E
        self.bounds = (lo, hi)
        L.check(lib.dcr_bfc_paper_sharded(self.d_rowptr.data_ptr(), self.d_col.data_ptr(), self.n, self.max_degree,
                                          self.d_esrc.data_ptr(), self.d_edst.data_ptr(), lo, hi - lo, self.comm.handle,
                                          self.scratch.data_ptr(), self.scratch_bytes, 0, 0, st), "dcr_bfc_paper_sharded")
        v, hv = self.comm.views, self.host_views()
        for k in FIELDS:
            hv[k][lo:hi].copy_(v[k][lo:hi], non_blocking=True)
        return lo, hi

    def bytes_per_pass(self):
        """(h2d, d2h) bytes of this rank for the last pass."""
        lo, hi = self.bounds if self.bounds else (0, self.n_edges)
        h2d = 4 * (self.n + 1) + 4 * self.nnz + (8 * self.n_edges if getattr(self, "_uploaded_edges", True) else 0)
        return h2d, 24 * (hi - lo) + (16 if self.world > 1 else 0)

    def close(self):
        if getattr(self, "_registered", None):
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(self._registered[0])
            self._registered = None
        if self.comm is not None:
            self.comm.close()
            self.comm = None


def interleave_reference(blocks, n_edges: int):
    """Host restatement of the strided shard geometry (used by the CPU world-size-2 test): ``blocks[r][t]`` is edge
    ``r + t*W``."""
    world = len(blocks)
    out = [None] * n_edges
    for r, b in enumerate(blocks):
        for t, v in enumerate(b):
            e = r + t * world
            if e < n_edges:
                out[e] = v
    return out
