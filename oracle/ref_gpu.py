"""Oracle (TEST INFRASTRUCTURE ONLY): the reference's OWN compiled numba kernels, run on the GPU.

``oracle/build_ref.py`` compiles ``curvature/bfc_cuda.py``'s two ``@cuda.jit`` kernels (unmodified, through
numba's NVVM pipeline) to PTX under ``oracle/_ref/``.  This module loads that PTX with the CUDA driver API
(``cuda.bindings.driver`` — no numba, no ``/root/reference`` needed at run time) and launches it exactly the way
the reference's host wrappers do.  It serves two purposes:

  * **pinning by execution**: the fp32 bit patterns of the compiled kernel (two roundings of fp64 arithmetic,
    SURVEY.md App. A.3) decide SDRF's argmin/argmax ties; ``tests/test_gpu_ref_kernels.py`` checks the
    ``"compiled"`` rounding model of ``oracle/cuda_flavour.py`` and the CUDA product path against these kernels
    bit for bit;
  * **the "numba bfc_cuda on the same B200" baseline** of ``bench.py`` (dense full-graph BFC and the SDRF loop
    driven the reference's way: two ``A @ A`` per iteration and one ``.item()`` per candidate).

Host-side statements restated here (the kernels themselves are the reference's):
  ``balanced_forman_curvature``  ``curvature/bfc_cuda.py:51-65``,
  ``balanced_forman_post_delta`` ``curvature/bfc_cuda.py:144-159``,
  ``sdrf_reference_gpu``         ``rewiring/sdrf_cuda_bfc.py:14-93`` (graph bookkeeping shared with ``oracle/sdrf.py``;
                                 both modes — ``is_undirected=False`` is there for the directed row of SURVEY.md §8f).

numba kernel ABI (numba ``Array`` data model, flattened): an ``ndim``-d array is passed as
``meminfo*, parent*, nitems:i64, itemsize:i64, data*, shape[ndim]:i64, strides[ndim]:i64`` (strides in bytes).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_state = {}


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "manifest.json"))


def _drv():
    from cuda.bindings import driver
    return driver


def _ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _load():
    if _state:
        return _state
    import torch
    torch.cuda.init()
    torch.zeros(1, device="cuda")                      # make torch's primary context current on this thread
    drv = _drv()
    with open(os.path.join(REF_DIR, "manifest.json")) as f:
        manifest = json.load(f)
    for name, k in manifest["kernels"].items():
        with open(os.path.join(REF_DIR, k["file"]), "rb") as f:
            ptx = f.read() + b"\0"
        mod = _ck(drv.cuModuleLoadData(ptx))           # driver JIT: .target sm_90 PTX -> sm_100 SASS
        fn = _ck(drv.cuModuleGetFunction(mod, k["entry"].encode()))
        _state[name] = (mod, fn, k["n_params"])
    _state["manifest"] = manifest
    return _state


def _array_params(t, itemsize=None, strides=None):
    """Flattened numba ``Array`` argument for a torch tensor (strides in bytes)."""
    es = t.element_size() if itemsize is None else itemsize
    st = [s * t.element_size() for s in t.stride()] if strides is None else strides
    vals = [0, 0, t.numel(), es, t.data_ptr()] + list(t.shape) + st
    types = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p] + [C.c_int64] * (2 * t.dim())
    return vals, types


def _launch(name, grid, block, args):
    import torch
    drv = _drv()
    _, fn, n_params = _load()[name]
    vals, types = [], []
    for a in args:
        if isinstance(a, tuple) and len(a) == 2 and isinstance(a[1], list):
            vals += a[0]
            types += a[1]
        else:
            v, ty = a
            vals.append(v)
            types.append(ty)
    assert len(vals) == n_params, (len(vals), n_params)
    stream = torch.cuda.current_stream().cuda_stream
    _ck(drv.cuLaunchKernel(fn, grid[0], grid[1], 1, block[0], block[1], 1, 0, stream,
                           (tuple(vals), tuple(types)), 0))


def balanced_forman_curvature(A, C_out=None):
    """``curvature/bfc_cuda.py:51-65`` with the reference's compiled kernel.  ``A``: fp32 CUDA ``[N,N]``."""
    import torch
    N = A.shape[0]
    A2 = torch.matmul(A, A)                                               # :53
    d_in = A.sum(axis=0)                                                  # :54
    d_out = A.sum(axis=1)                                                 # :55
    if C_out is None:
        C_out = torch.zeros(N, N, device=A.device)                        # :56-57
    tpb = (16, 16)                                                        # :59
    grid = (math.ceil(N / tpb[0]), math.ceil(N / tpb[1]))                 # :60-62
    _launch("_balanced_forman_curvature", grid, tpb,
            [_array_params(A), _array_params(A2), _array_params(d_in), _array_params(d_out),
             (N, C.c_int32), _array_params(C_out)])                       # :64
    return C_out


def balanced_forman_post_delta(A, x, y, i_neighbors, j_neighbors, D=None, A2=None):
    """``curvature/bfc_cuda.py:144-159``.  The neighbour lists go in as int64 arrays read through the kernel's
    ``int32[:]`` view with 8-byte strides, as the reference passes ``np.array(list)`` (:157-158)."""
    import torch
    N = A.shape[0]
    if A2 is None:
        A2 = torch.matmul(A, A)                                           # :146
    d_in = float(A[:, x].sum())                                           # :147  (c_float(tensor) -> .item())
    d_out = float(A[y].sum())                                             # :148
    if D is None:
        D = torch.zeros(len(i_neighbors), len(j_neighbors), device=A.device)   # :149-150
    tpb = (16, 16)
    grid = (math.ceil(D.shape[0] / tpb[0]), math.ceil(D.shape[1] / tpb[1]))
    i_nb = torch.as_tensor(np.array(i_neighbors, dtype=np.int64)).to(A.device)  # numba copies host arrays H2D
    j_nb = torch.as_tensor(np.array(j_neighbors, dtype=np.int64)).to(A.device)
    _launch("_balanced_forman_post_delta", grid, tpb,
            [_array_params(A), _array_params(A2), (d_in, C.c_float), (d_out, C.c_float), (N, C.c_int32),
             _array_params(D), (int(x), C.c_int32), (int(y), C.c_int32),
             _array_params(i_nb, itemsize=8, strides=[8]), _array_params(j_nb, itemsize=8, strides=[8]),
             (D.shape[0], C.c_int32), (D.shape[1], C.c_int32)])
    return D


def sdrf_reference_gpu(edge_index, num_nodes, loops, remove_edges, removal_bound, tau, uniforms,
                       time_budget_s: float | None = None, is_undirected: bool = True, batched_items: bool = False):
    """``rewiring/sdrf_cuda_bfc.py:14-93`` driven the reference's way on the GPU (``is_undirected=True``):
    full ``balanced_forman_curvature`` every iteration, ``C.argmin().item()``, a second ``A @ A`` inside
    ``balanced_forman_post_delta`` and one ``.item()`` per candidate.  Returns ``(edge_index_out, log)`` with
    the log records of :func:`oracle.sdrf.sdrf_oracle`; stops early (log shorter than ``loops``) once
    ``time_budget_s`` of wall clock is spent.  ``batched_items=True`` (long parity runs only, never the timed baseline)
    reads the improvements with ONE device-to-host copy of ``D - C[x,y]`` instead of one ``.item()`` per candidate —
    the same fp32 subtraction, the same values, ~40x less wall clock on hub candidates."""
    import time

    import torch

    from .sdrf import (choice_index, dense_adjacency, edge_index_from_adjacency, edge_index_from_digraph,
                       networkx_adjacency, networkx_digraph, softmax)

    A = torch.from_numpy(dense_adjacency(edge_index, is_undirected)).cuda()   # :26-29, :34
    N = A.shape[0]                                                        # :30
    if is_undirected:
        adj, pred = networkx_adjacency(edge_index, max(num_nodes, N)), None   # :31-33
    else:
        adj, pred = networkx_digraph(edge_index, max(num_nodes, N))       # :31 (adj = successors)
    Cm = torch.zeros(N, N, device="cuda")                                 # :35
    n_draws = 0
    log = []
    t0 = time.perf_counter()
    for _ in range(loops):                                                # :37
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            break
        can_add = True                                                    # :38
        balanced_forman_curvature(A, C_out=Cm)                            # :39
        ix_min = Cm.argmin().item()                                       # :40
        x, y = ix_min // N, ix_min % N                                    # :41-42
        x_neighbors = list(adj[x]) + [x]                                  # :45 / :48
        y_neighbors = list(adj[y] if is_undirected else pred[y]) + [y]    # :46 / :49
        cand_pos = [(I, J) for I, i in enumerate(x_neighbors) for J, j in enumerate(y_neighbors)
                    if (i != j) and (j not in adj[i])]                    # :50-54 (positions: the lists hold no repeats)
        candidates = [(x_neighbors[I], y_neighbors[J]) for I, J in cand_pos]
        rec = {"x": x, "y": y, "n_candidates": len(candidates), "k": -1, "l": -1, "choice": -1,
               "removed": None, "improvements": np.zeros(0)}
        stop = False
        if len(candidates):                                               # :56
            D = balanced_forman_post_delta(A, x, y, x_neighbors, y_neighbors)     # :57
            if batched_items:
                diff = (D - Cm[x, y]).cpu().numpy()
                pos = np.array(cand_pos)
                improvements = diff[pos[:, 0], pos[:, 1]].astype(np.float64)
            else:
                improvements = []
                for (i, j) in candidates:                                 # :58-62
                    improvements.append((D - Cm[x, y])[x_neighbors.index(i), y_neighbors.index(j)].item())
                improvements = np.array(improvements)
            choice = choice_index(softmax(improvements, tau=tau), float(uniforms[n_draws]))   # :64-68
            n_draws += 1
            k, l = candidates[choice]
            adj[k][l] = None                                              # :69
            if is_undirected:
                adj[l][k] = None
                A[k, l] = A[l, k] = 1                                     # :70-71
            else:
                pred[l][k] = None
                A[k, l] = 1                                               # :72-73
            rec.update(k=k, l=l, choice=choice, improvements=improvements)
        else:
            can_add = False                                               # :75
            if not remove_edges:                                          # :76-77
                stop = True
        if remove_edges and not stop:                                     # :79
            ix_max = Cm.argmax().item()                                   # :80
            xr, yr = ix_max // N, ix_max % N                              # :81-82
            if Cm[xr, yr] > removal_bound:                                # :83
                if yr not in adj[xr]:
                    raise KeyError(f"The edge {xr}-{yr} is not in the graph")
                del adj[xr][yr]                                           # :84
                if is_undirected:
                    if xr != yr:
                        del adj[yr][xr]
                    A[xr, yr] = A[yr, xr] = 0                             # :85-86
                else:
                    del pred[yr][xr]
                    A[xr, yr] = 0                                         # :87-88
                rec["removed"] = (xr, yr)
            elif can_add is False:                                        # :89-91
                stop = True
        log.append(rec)
        if stop:
            break
    return (edge_index_from_adjacency(adj) if is_undirected else edge_index_from_digraph(adj)), log   # :93
