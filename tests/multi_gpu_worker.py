"""Two (or more) ranks under torchrun on one box: the multi-GPU paths of full-graph BFC against the single-GPU pass.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 tests/multi_gpu_worker.py

Checked on EVERY rank: the peer-memory route (contiguous work-balanced ranges, fused value + all-gather kernel over CUDA
IPC mappings) over several passes, the NCCL route (strided shards + all-gather + re-interleave), and the host form (each
rank copies its slice into one shared host block).  Prints "multi_gpu_worker ok" on rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "discrete-curvature-rewiring_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dcr import bfc, graph
    from dcr import dist as ddist
    from dcr.synth import named_graph
    for name in ("cora", "squirrel"):
        ei, n = named_graph(name)
        rowptr, col = graph.undirected_csr(ei, n)
        csr = bfc.DeviceCSR.from_host(rowptr, col, device=torch.device("cuda", local))
        full = {k: v.clone() for k, v in bfc.paper_flavour(csr).items() if k in ddist.FIELDS}
        for mode in ("peer", "nccl"):
            sh = ddist.ShardedPaperBFC(csr, mode=mode)
            assert sh.mode == mode
            for it in range(4):                         # repeated passes: the hand-shake epochs advance
                out = sh.run()
                torch.cuda.synchronize()
                for k in ddist.FIELDS:
                    assert torch.equal(out[k], full[k]), (name, mode, it, k, rank)
                if mode == "peer":                     # scribble over the local buffer: the next pass must restore it
                    for k in ddist.FIELDS:
                        out[k].zero_()
                dist.barrier()
            sh.check()
            sh.close()
        # cuda flavour: two passes with the all-gather of the supports fused between them
        want = {k: v.clone() for k, v in bfc.cuda_flavour_edges(csr).items()}
        sc = ddist.ShardedCudaBFC(csr)
        for it in range(4):
            out = sc.run()
            torch.cuda.synchronize()
            for k in ddist.CUDA_FIELDS:
                a, b = out[k], want[k]
                if a.dtype.is_floating_point:
                    a, b = a.view(torch.int64 if a.dtype == torch.float64 else torch.int32), \
                        b.view(torch.int64 if b.dtype == torch.float64 else torch.int32)
                assert torch.equal(a, b), (name, "cuda-sharded", it, k, rank)
            for k in ddist.CUDA_FIELDS:
                out[k].zero_()
            dist.barrier()
        sc.check()
        sc.close()
        m = ei[0] < ei[1]
        esrc, edst = ei[0][m].astype(np.int32), ei[1][m].astype(np.int32)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hs = ddist.HostShardedPaperBFC(n, col.size, esrc.size, csr.max_degree)
        h = (pin(rowptr.astype(np.int32)), pin(col), pin(esrc), pin(edst))
        for it in range(2):
            lo, hi = hs.run(*(h if it == 0 else h[:2]))     # second pass: CSR only, edge list derived on the device
            torch.cuda.synchronize()
            dist.barrier()
            hv = hs.host_views()
            for k in ddist.FIELDS:                      # every rank sees the whole shared block
                assert np.array_equal(hv[k].numpy(), full[k].cpu().numpy()), (name, "host", it, k, rank)
            dist.barrier()
        hs.close()
    dist.barrier()
    if rank == 0:
        print("multi_gpu_worker ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
