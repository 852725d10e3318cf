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
