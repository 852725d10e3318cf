"""CPU suite, part 3: the N>1 path of full-graph BFC with world_size=2 over gloo.

The compute kernel needs a GPU, so here each rank fills its PaperWorkspace-shaped result block with the ORACLE values
of its own shard (edges e = rank + t*world) — the same block layout (`bfc` f64 | tri | sq_i | sq_j | gamma int32) the
CUDA path gathers — then the ranks all-gather the blocks and the host restatement of `dcr_bfc_paper_unshard`
re-interleaves them.  This exercises shard geometry, padding of the last chunk, the single-buffer collective and the
rendezvous, which are the parts of the multi-GPU path that do not depend on the device.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, REPO


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _unshard_host(gathered: np.ndarray, world: int, chunk: int, n_edges: int):
    """numpy restatement of unshard_kernel (dcr_bfc_paper.cu)."""
    out = {k: np.zeros(n_edges, dtype=np.int32) for k in ("tri", "sq_i", "sq_j", "gamma")}
    out["bfc"] = np.zeros(n_edges, dtype=np.float64)
    block = chunk * 24
    for r in range(world):
        base = gathered[r * block:(r + 1) * block]
        ids = np.arange(r, n_edges, world)
        t = np.arange(ids.size)
        out["bfc"][ids] = base[: chunk * 8].view(np.float64)[t]
        ints = base[chunk * 8:].view(np.int32)
        for q, k in enumerate(("tri", "sq_i", "sq_j", "gamma")):
            out[k][ids] = ints[q * chunk + t]
    return out


def _worker(rank, world, port, n_edges_path):
    for p in (PKG, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dcr.bfc import shard_count
        from dcr.dist import chunk_size
        from dcr.synth import named_graph
        from oracle.paper_flavour import bfc_paper
        ei, n = named_graph("cornell")
        ref = bfc_paper(ei, n)
        E = len(ref["bfc"])
        chunk = chunk_size(E, world)
        count = shard_count(E, rank, world)
        assert count == len(range(rank, E, world))
        block = np.zeros(chunk * 24, dtype=np.uint8)
        block[: chunk * 8].view(np.float64)[:count] = ref["bfc"][rank::world]
        ints = block[chunk * 8:].view(np.int32)
        for q, k in enumerate(("tri", "sq_i", "sq_j", "gamma")):
            ints[q * chunk: q * chunk + count] = ref[k][rank::world]
        gathered = torch.zeros(world * chunk * 24, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, torch.from_numpy(block))
        full = _unshard_host(gathered.numpy(), world, chunk, E)
        for k in ("tri", "sq_i", "sq_j", "gamma", "bfc"):
            assert np.array_equal(full[k], ref[k]), k
        # every rank ends with the identical full result
        digest = torch.tensor([float(full["bfc"].sum()), float(full["tri"].sum())], dtype=torch.float64)
        both = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        assert all(torch.equal(both[0], b) for b in both)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bfc_gather_world_size_n_gloo(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)


def test_shard_geometry_covers_every_edge_once():
    from dcr.bfc import shard_count
    from dcr.dist import chunk_size, interleave_reference
    for n_edges in (0, 1, 7, 8, 9, 295, 1166243):
        for world in (1, 2, 3, 4, 8):
            counts = [shard_count(n_edges, r, world) for r in range(world)]
            assert sum(counts) == n_edges
            assert max(counts) <= chunk_size(n_edges, world)
            if n_edges <= 295:
                blocks = [list(range(r, n_edges, world)) for r in range(world)]
                assert interleave_reference(blocks, n_edges) == list(range(n_edges))


# ---- contiguous work-balanced ranges (the peer-memory route): host logic over gloo --------------------------------
def _edge_cost_host(rowptr, col, esrc, edst):
    """numpy restatement of edge_cost_kernel (dcr_bfc_paper.cu)."""
    deg = np.diff(rowptr.astype(np.int64))
    rows = np.repeat(np.arange(deg.size), deg)
    S = np.bincount(rows, weights=deg[col].astype(np.float64), minlength=deg.size).astype(np.int64)
    di, dj = deg[esrc], deg[edst]
    ca, cb = S[edst] - di, S[esrc] - dj
    sw = cb < ca
    cost = np.where(sw, cb, ca) + 24 * np.where(sw, di, dj) + 160
    return np.where(np.minimum(di, dj) <= 1, 2, cost).astype(np.int64)


def _range_worker(rank, world, port, out_path):
    for p in (PKG, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dcr import graph
        from dcr.dist import balanced_bounds
        from dcr.synth import named_graph
        from oracle.paper_flavour import bfc_paper
        ei, n = named_graph("wisconsin")
        rowptr, col = graph.undirected_csr(ei, n)
        m = ei[0] < ei[1]
        esrc, edst = ei[0][m], ei[1][m]
        cost = _edge_cost_host(rowptr, col, esrc, edst)
        b = balanced_bounds(np.cumsum(cost), world)
        # every rank derives the same cuts
        got = [None] * world
        dist.all_gather_object(got, b)
        assert all(g == b for g in got)
        lo, hi = b[rank], b[rank + 1]
        # each rank computes its contiguous range with the oracle and writes it at the edges' own positions of a full
        # block (what the closing kernel does through peer memory); summing the blocks = the gather
        ref = bfc_paper(ei, n)
        E = len(ref["bfc"])
        full = torch.zeros(E, dtype=torch.float64)
        full[lo:hi] = torch.from_numpy(ref["bfc"][lo:hi])
        owner = torch.zeros(E, dtype=torch.int32)
        owner[lo:hi] = 1
        dist.all_reduce(full)
        dist.all_reduce(owner)
        assert bool((owner == 1).all())                    # the ranges tile [0, E) exactly once
        assert np.array_equal(full.numpy(), ref["bfc"])
        if rank == 0:
            work = [int(cost[b[r]:b[r + 1]].sum()) for r in range(world)]
            np.save(out_path, np.array(work))
    finally:
        dist.destroy_process_group()


def test_balanced_contiguous_ranges_world2_gloo(tmp_path):
    out = str(tmp_path / "work.npy")
    mp.spawn(_range_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    work = np.load(out)
    assert work.min() > 0 and work.max() / work.sum() < 0.6          # two ranks: neither holds more than 60 % of the work


def test_balanced_bounds_properties():
    sys.path.insert(0, PKG)
    from dcr.dist import balanced_bounds
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        cost = rng.integers(1, 1000, size=5000).astype(np.int64)
        cost[rng.integers(0, 5000, size=20)] = 200000                # hubs
        pre = np.cumsum(cost)
        b = balanced_bounds(pre, world)
        assert b[0] == 0 and b[-1] == 5000 and all(b[k] <= b[k + 1] for k in range(world))
        work = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert work.max() <= pre[-1] / world + cost.max()             # no range exceeds its share by more than one edge
        assert balanced_bounds(torch.from_numpy(pre), world) == b     # torch and numpy inputs agree
    assert balanced_bounds(np.zeros(0, dtype=np.int64), 4) == [0, 0, 0, 0, 0]
