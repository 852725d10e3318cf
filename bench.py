#!/usr/bin/env python
"""bench.py — BFC edges/s (full-graph paper-flavour BFC, arxiv-shaped synthetic graph) and SDRF iterations/s.

    python bench.py --gpus N --steps K --warmup W            # our arm (libdcr.so, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle's C port on the host cores

One JSON line on stdout (rank 0).  A "step" is one full pass of the hot path over the whole graph: plan + edge
kernels (+ all-gather and re-interleave when N > 1).  Timing: CUDA events per step on the launching stream, an L2
flush (256 MiB memset) between steps outside the events, barrier + synchronize around the timed region, max over
ranks.  See DESIGN.md §Measurement for the definitions of `value`, `e2e`, `roofline` and `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "discrete-curvature-rewiring_b200")
for _p in (PKG, REPO):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

DENSE_WORKLOAD = ("squirrel-shaped synthetic graph (N=5201, E=198000): full-graph cuda-flavour BFC in the dense regime — "
                  "supports A2[i,j] on the edges via tcgen05 int8 A·A (TMA, TMEM, fused epilogue), then the per-entry "
                  "closing pass; the sparse sorted-list support kernel is timed beside it")

WORKLOADS = {
    "arxiv": "arxiv-shaped synthetic graph (N=169343, E=1166243, Chung-Lu alpha=0.6, p_tri=0.1, seed 169343): "
             "full-graph paper-flavour BFC (deg, #tri, #sq_i, #sq_j, gamma_max, fp64 value per undirected edge)",
    "squirrel": "squirrel-shaped synthetic graph (N=5201, E=198000): full-graph paper-flavour BFC",
    "cora": "cora-shaped synthetic graph (N=2708, E=5278): full-graph paper-flavour BFC",
}


# DRAM traffic of the edge kernels of ONE pass over the arxiv-shaped graph, from the committed ncu capture
# (17.64 + 22.60 + 26.82 MB read, 0.23 + 1.58 + 3.26 MB written): the CSR is L2-resident, so this is ~1000x below the
# algorithmic gather bytes.
NCU_DRAM_BYTES_PER_PASS = 78.37e6


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(rowptr: np.ndarray, col: np.ndarray, esrc: np.ndarray, edst: np.ndarray):
    """B_gather(e) = 16 + 4(d_i+d_j) + 4((S_i-d_j)+(S_j-d_i)) + 24, S_v = sum of neighbour degrees.
    Also B_compulsory = 4(N+1) + 4*2E + 24E (CSR once + results once)."""
    deg = np.diff(rowptr.astype(np.int64))
    n = deg.size
    rows = np.repeat(np.arange(n), deg)
    S = np.bincount(rows, weights=deg[col].astype(np.float64), minlength=n).astype(np.int64)
    di, dj = deg[esrc], deg[edst]
    per_edge = 16 + 4 * (di + dj) + 4 * ((S[esrc] - dj) + (S[edst] - di)) + 24
    compulsory = 4 * (n + 1) + 4 * col.size + 24 * esrc.size
    return per_edge, int(compulsory)


class ClockSampler:
    """SM clock and throttle reasons for the clocks record of the JSON line.

    * SM clock DURING the timed steps: an in-band probe (`dcr_sm_clock_probe`: one thread, clock64 / globaltimer)
      enqueued on a side stream between the timed steps, so it runs while their kernels run.
    * Throttle reasons: NVML, in-process — but NOT while the timed steps execute.  On these shared hosts every form of
      NVML / nvidia-smi polling tried (nvidia-smi -lms, nvmlDeviceGetClockInfo, nvmlDeviceGetCurrentClocksEventReasons)
      sporadically stalled GPU work for 20-200 ms, turning one 7 ms step into a 170 ms one (step_ms lists in
      profiles/r01_clock_sampling_perturbation.txt; without polling every step is within 0.02 ms of the others).
      The reasons are therefore sampled every 20 ms over an UNTIMED batch of the same steps run back to back right
      after the timed region (`under_load`), i.e. under the identical load."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.masks = []
        self.stop_flag = False
        self.handle = None
        self.nvml = None
        self.max_mhz = None
        self.probe_out = None
        self.probe_stream = None
        self.n_probes = 0
        self.probe_where = "during the timed steps"

    def start(self):
        import torch
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if visible:
                try:
                    idx = int(visible.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.probe_out = torch.zeros(64, dtype=torch.float32, device=f"cuda:{self.gpu}")
        self.probe_stream = torch.cuda.Stream(device=self.gpu)

    def _poll(self):
        while not self.stop_flag:
            try:
                try:
                    m = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    m = self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.masks.append(int(m))
            except Exception:
                pass
            time.sleep(0.02)

    def probe(self):
        """Enqueue one SM-clock probe on the side stream (call between the enqueues of the timed steps)."""
        if self.probe_out is None or self.n_probes >= self.probe_out.numel():
            return
        from dcr import lib as L
        L.check(L.load().dcr_sm_clock_probe(self.probe_out[self.n_probes:].data_ptr(), self.probe_stream.cuda_stream),
                "dcr_sm_clock_probe")
        self.n_probes += 1

    def under_load(self, step_fn, steps: int, exact: bool = False):
        """Poll the throttle reasons while untimed repetitions of the timed step execute (`exact`: exactly `steps`
        of them — needed when the step contains collectives and every rank must run the same count)."""
        import torch
        if self.nvml is None:
            for _ in range(steps if exact else 0):
                step_fn()
            return
        self.stop_flag = False
        th = threading.Thread(target=self._poll, daemon=True)
        th.start()
        t_end = time.perf_counter() + (0.0 if exact else 0.25)
        done = 0
        while done < steps or time.perf_counter() < t_end:       # single GPU: at least a quarter of a second of load
            step_fn()
            done += 1
            if done % 4 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        self.stop_flag = True
        th.join(timeout=3)

    def result(self) -> dict:
        import torch
        torch.cuda.synchronize()
        mhz = [float(v) for v in self.probe_out[: self.n_probes].cpu().tolist() if v > 0] if self.n_probes else []
        mask = 0
        for m in self.masks:
            mask |= m
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": self.max_mhz,
                "reasons": [name for name, bit in self.REASONS if mask & bit],
                "samples": len(mhz), "reason_samples": len(self.masks),
                "source": f"sm_mhz: in-band clock64/globaltimer probes on a side stream {self.probe_where}; reasons: NVML "
                          "every 20 ms over an untimed batch of the same steps run right after the timed region (NVML polling "
                          "inside the region stalls GPU work on these hosts, see profiles/)"}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_bfc_rate(rowptr, col, esrc, edst, threads: int, budget_s: float, seed: int = 0):
    """Time the oracle's C port (bfc_naive restated, pthreads) on a seeded uniform sample of the edges sized for
    about `budget_s` seconds.  Returns (edges/s, sample size, seconds)."""
    from oracle.c_port import bfc_paper_c
    rng = np.random.default_rng(seed)
    perm = rng.permutation(esrc.size)
    probe = perm[: min(esrc.size, 4096)]
    t0 = time.perf_counter()
    bfc_paper_c(rowptr, col, esrc[probe], edst[probe], threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    m = int(min(esrc.size, max(4096, budget_s * probe.size / dt)))
    pick = perm[:m]
    t0 = time.perf_counter()
    bfc_paper_c(rowptr, col, esrc[pick], edst[pick], threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    return m / dt, m, dt


def build_graph(workload: str):
    from dcr import graph
    from dcr.synth import named_graph
    ei, n = named_graph(workload)
    rowptr, col = graph.undirected_csr(ei, n)
    m = ei[0] < ei[1]
    esrc = ei[0][m].astype(np.int32)
    edst = ei[1][m].astype(np.int32)
    return ei, n, rowptr, col, esrc, edst


# ------------------------------------------------------------------------------------------------------------
# reference arm: the CPU implementation of the path on the host cores
# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__
    __graft_entry__.build()
    ei, n, rowptr, col, esrc, edst = build_graph(args.workload)
    threads = host_threads()
    steps = max(1, args.steps)
    budget = min(20.0, 120.0 / (steps + args.warmup))       # whole run stays within a few minutes
    rates, m = [], 0
    for s in range(args.warmup + steps):
        rate, m, dt = cpu_bfc_rate(rowptr, col, esrc, edst, threads, budget, seed=s)
        if s >= args.warmup:
            rates.append((m, dt))
    tot_m = sum(a for a, _ in rates)
    tot_t = sum(b for _, b in rates)
    value = tot_m / tot_t
    sample = (f"oracle C port of curvature/bfc_naive.py (oracle/c/bfc_paper_csr.c), {threads} pthreads, per step a "
              f"seeded uniform sample of {m} of the {esrc.size} undirected edges")
    line = {
        "impl": "reference", "metric": "bfc_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32 counts + f64 value", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "sample_edges_per_step": m},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def sdrf_bench(args, torch):
    """SDRF iterations/s on the cora-shaped graph (config 3), device loop vs the dense CPU restatement."""
    from dcr import graph, sdrf
    from dcr.synth import named_graph
    from oracle.sdrf import sdrf_oracle
    ei, n = named_graph("cora")
    loops, tau, bound = args.sdrf_loops, 163, 0.95
    uni = np.random.RandomState(3).random_sample(loops)
    out = {"workload": f"cora-shaped synthetic graph (N=2708, E=5278), {loops} iterations, tau={tau}, "
                       f"removal_bound={bound} (utils/hyperparams.py Cora), stochastic draw with host uniforms"}
    # device-resident loop only (state already built): CUDA events around the single persistent-kernel launch
    rowptr_o, order = graph.networkx_order(ei, n)
    best = None
    for rep in range(3):
        st = sdrf.SdrfState(rowptr_o, order, max_additions=loops)
        u_dev = torch.from_numpy(uni).cuda()
        log = torch.empty((loops, 8), dtype=torch.int32, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        res, _ = st.run(loops, True, bound, tau, u_dev, log=log)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        done = res["iterations_done"]
        st.close()
        if res["status"] == 0 and (best is None or ms < best[0]):
            best = (ms, done)
    if best is not None:
        out["iters_per_s"] = best[1] / (best[0] * 1e-3)
        out["iterations"] = best[1]
        out["ms_total"] = best[0]
    # end to end through the public entry point: host edge_index in, host edge_index out (set-up + loop + export)
    t0 = time.perf_counter()
    got, glog = sdrf.sdrf(ei, n, loops, True, bound, tau, uniforms=uni, return_log=True)
    dt = time.perf_counter() - t0
    out["e2e_iters_per_s"] = len(glog) / dt
    out["e2e_s"] = dt
    # CPU: dense restatement of bfc_cuda.py + sdrf_cuda_bfc.py (two A@A per iteration, like the reference)
    cpu_iters = args.sdrf_cpu_iters
    t0 = time.perf_counter()
    want, wlog = sdrf_oracle(ei, n, cpu_iters, True, bound, tau, uni, rounding="compiled", incremental_a2=False)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": len(wlog) / dt, "unit": "iterations/s", "cores": host_threads(), "kind": "port",
                           "sample": f"first {cpu_iters} iterations of the same run, dense numpy restatement "
                                     "(oracle/sdrf.py, BLAS A@A + per-entry numpy loops)"}
    same = [tuple(int(v) for v in r) for r in glog[:cpu_iters]] == [
        (r["x"], r["y"], r["n_candidates"], r["k"], r["l"], r["choice"],
         -1 if r["removed"] is None else r["removed"][0], -1 if r["removed"] is None else r["removed"][1])
        for r in wlog]
    out["prefix_matches_cpu"] = bool(same)
    if "iters_per_s" in out:
        out["speedup_vs_cpu"] = out["iters_per_s"] / out["cpu_baseline"]["value"]
    # The reference's OWN numba kernels on this GPU (oracle/_ref PTX, built from /root/reference by oracle/build_ref.py,
    # loaded with the driver API) driven the reference's way: two A@A + an N^2 x N kernel per iteration, one .item()
    # per candidate.  A reported baseline ("the repo's numba bfc_cuda on the same B200"), present when oracle/_ref is.
    try:
        from oracle import ref_gpu
        if ref_gpu.available():
            budget = 6.0
            ref_gpu.sdrf_reference_gpu(ei, n, 2, True, bound, tau, uni)                       # JIT + warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, rlog = ref_gpu.sdrf_reference_gpu(ei, n, loops, True, bound, tau, uni, time_budget_s=budget)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            k = len(rlog)
            same_ref = [tuple(int(v) for v in r) for r in glog[:k]] == [
                (r["x"], r["y"], r["n_candidates"], r["k"], r["l"], r["choice"],
                 -1 if r["removed"] is None else r["removed"][0], -1 if r["removed"] is None else r["removed"][1])
                for r in rlog]
            from curvature.bfc_cuda import balanced_forman_curvature as ours_dense
            def timed(fn, reps):
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps
            dense = []          # balanced_forman_curvature(A) on dense A: configs 1-4 (SURVEY.md §8d baseline 3)
            for shape in ("cornell", "wisconsin", "cora", "squirrel"):
                gei, gn = named_graph(shape)
                A = torch.zeros(gn, gn, device="cuda")
                A[torch.from_numpy(gei[0]).cuda(), torch.from_numpy(gei[1]).cuda()] = 1
                big = gn > 4000
                r_ms = timed(lambda: ref_gpu.balanced_forman_curvature(A), 1 if big else 3)
                o_ms = timed(lambda: ours_dense(A), 10)
                same_bits = bool(torch.equal(ref_gpu.balanced_forman_curvature(A).view(torch.int32),
                                             ours_dense(A).view(torch.int32)))
                dense.append({"shape": shape, "n": gn, "reference_ms": r_ms, "ours_dropin_ms": o_ms,
                              "bit_identical": same_bits})
                del A
            # config 2: the WebKB shapes with the reference's own hyper-parameters, whole rewiring, both sides end to end
            from dcr.synth import SDRF_PARAMS
            config2 = []
            for shape in ("wisconsin", "texas"):
                gei, gn = named_graph(shape)
                lp, tu, bd = SDRF_PARAMS[shape]
                for t_ in (tu, float("inf")):
                    u2 = np.random.RandomState(11).random_sample(lp)
                    sdrf.sdrf(gei, gn, lp, True, bd, t_, uniforms=u2)                     # warm-up
                    t0 = time.perf_counter()
                    g_out, g_log = sdrf.sdrf(gei, gn, lp, True, bd, t_, uniforms=u2, return_log=True)
                    ours_s = time.perf_counter() - t0
                    t0 = time.perf_counter()
                    r_out, r_log = ref_gpu.sdrf_reference_gpu(gei, gn, lp, True, bd, t_, u2)
                    torch.cuda.synchronize()
                    ref_s = time.perf_counter() - t0
                    config2.append({"shape": shape, "loops": lp, "tau": "inf" if t_ == float("inf") else t_,
                                    "removal_bound": bd, "iterations": len(r_log), "ours_e2e_s": ours_s,
                                    "reference_numba_s": ref_s, "identical_edge_index": bool(np.array_equal(g_out, r_out))})
            ref_ms = [d["reference_ms"] for d in dense if d["shape"] == "cora"][0]
            our_ms = [d["ours_dropin_ms"] for d in dense if d["shape"] == "cora"][0]
            out["reference_numba_on_this_gpu"] = {
                "sdrf_iters_per_s": k / dt, "iterations": k, "wall_s": dt, "sequence_prefix_matches_ours": bool(same_ref),
                "dense_bfc_ms_reference": ref_ms, "dense_bfc_ms_ours_dropin": our_ms, "dense_bfc": dense, "sdrf_config2": config2,
                "what": "unmodified numba kernels of curvature/bfc_cuda.py (PTX via numba.cuda.compile_ptx, driver JIT to "
                        "sm_100) + the reference's host statements (oracle/ref_gpu.py); dense_bfc = "
                        "balanced_forman_curvature(A) on the same cora-shaped dense A, ours through the drop-in module"}
            if "iters_per_s" in out:
                out["speedup_vs_reference_numba"] = out["iters_per_s"] / (k / dt)
    except Exception as exc:                     # a reported baseline must never take the bench down
        out["reference_numba_on_this_gpu"] = {"unavailable": repr(exc)[:200]}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        __graft_entry__.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libdcr has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL prints its version banner on stdout; stdout carries one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from dcr import bfc
    from dcr import lib as L
    from dcr.dist import ShardedPaperBFC
    L.load()

    ei, n, rowptr, col, esrc, edst = build_graph(args.workload)
    E = int(esrc.size)
    per_edge_bytes, b_compulsory = algorithmic_bytes(rowptr, col, esrc, edst)
    b_gather_total = int(per_edge_bytes.sum())
    b_gather_rank = int(per_edge_bytes[rank::world].sum())
    peak, peak_src = load_peaks()

    dev = torch.device("cuda", local)
    csr = bfc.DeviceCSR.from_host(rowptr, col, device=dev)
    csr._edges = (torch.from_numpy(esrc).to(dev), torch.from_numpy(edst).to(dev), None)
    sh = ShardedPaperBFC(csr)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def new_events(k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for a, b in evs:
            a.record()
            b.record()      # materialise the cudaEvent handles (the edge-kernel pair is recorded by the library)
        return evs

    # ---- device-resident throughput -------------------------------------------------------------------------
    # All K steps are ENQUEUED without a host sync in between (per-step CUDA events on the stream, L2 flush between
    # steps outside the events): the GPU never waits for the host, so host-side hiccups (e.g. the NVML sampler
    # thread) cannot leak into the device times.
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    # N > 1: 2-GPU runs intermittently showed one or two 7-120 ms all-gathers among the first timed steps (never at
    # N = 1, never in the run without the clock probes).  Extra untimed steps and a host head start did not remove
    # them; the only rank-asymmetric work inside the timed region was rank 0's in-band clock probe, so at N > 1 the
    # probes now ride on the untimed batch of the same steps that is run right after (where NVML is polled too).
    settle = 5 if world > 1 else 0
    for _ in range(args.warmup + settle):
        flush.zero_()
        sh.run()
    step_ev, edge_ev = new_events(args.steps), new_events(args.steps)
    barrier()
    wall0 = time.perf_counter()
    res = None
    for k in range(args.steps):
        flush.zero_()
        step_ev[k][0].record()
        res = sh.run(events=edge_ev[k])
        step_ev[k][1].record()
        if world == 1 and not args.no_clocks and k % max(1, args.steps // 8) == 0:
            sampler.probe()         # runs on a side stream while this step's kernels execute
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in step_ev]
    edge_ms = [a.elapsed_time(b) for a, b in edge_ev]
    # where a step's time goes on this rank: planning (step begin -> first edge kernel), the edge kernels + value
    # kernel, and what follows them (all-gather + re-interleave at N > 1)
    plan_ms = [s0.elapsed_time(e0) for (s0, _), (e0, _) in zip(step_ev, edge_ev)]
    tail_ms = [e1.elapsed_time(s1) for (_, s1), (_, e1) in zip(step_ev, edge_ev)]

    def untimed_step():
        flush.zero_()
        sh.run()
        if world > 1 and rank == 0 and not args.no_clocks:
            sampler.probe()         # N > 1: SM-clock probes ride on the untimed batch of the same steps (see below)

    if world > 1:
        sampler.probe_where = ("during an untimed batch of the same steps run right after the timed region (at N > 1 the "
                               "probe is kept out of the timed steps: it was the only rank-asymmetric work in them)")
        n_untimed = max(args.steps, 16)              # every rank runs the same untimed batch (collectives inside)
        if rank == 0 and not args.no_clocks:
            sampler.under_load(untimed_step, n_untimed, exact=True)
        else:
            for _ in range(n_untimed):
                untimed_step()
        barrier()
        clocks = sampler.result() if (rank == 0 and not args.no_clocks) else None
    else:
        clocks = None
        if not args.no_clocks:
            sampler.under_load(untimed_step, args.steps)
            clocks = sampler.result()
    total_ms = max_over_ranks(float(np.sum(step_ms)))
    ms_per_step = total_ms / args.steps
    value = E / (ms_per_step * 1e-3)
    edge_ms_avg = float(np.mean(edge_ms))

    # ---- end to end: host CSR in (pinned), host results out, copies inside the timed region -------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_rowptr, h_col, h_esrc, h_edst = pin(rowptr), pin(col), pin(esrc), pin(edst)
    h_out = torch.empty(E * 24, dtype=torch.uint8).pin_memory()
    d_rowptr = torch.empty_like(h_rowptr, device=dev)
    d_col = torch.empty_like(h_col, device=dev)
    d_esrc = torch.empty_like(h_esrc, device=dev)
    d_edst = torch.empty_like(h_edst, device=dev)
    csr2 = bfc.DeviceCSR(d_rowptr, d_col, n, csr.max_degree)
    csr2._edges = (d_esrc, d_edst, None)
    sh2 = ShardedPaperBFC(csr2)
    h2d = sum(int(t.numel() * t.element_size()) for t in (h_rowptr, h_col, h_esrc, h_edst))
    d2h = E * 24

    def e2e_step():
        d_rowptr.copy_(h_rowptr, non_blocking=True)
        d_col.copy_(h_col, non_blocking=True)
        d_esrc.copy_(h_esrc, non_blocking=True)
        d_edst.copy_(h_edst, non_blocking=True)
        r = sh2.run()
        o = 0
        for key in ("bfc", "tri", "sq_i", "sq_j", "gamma"):
            src = r[key].view(torch.uint8)
            h_out[o:o + src.numel()].copy_(src, non_blocking=True)
            o += src.numel()

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    e2e_ev = new_events(args.steps)
    barrier()
    for k in range(args.steps):
        flush.zero_()
        e2e_ev[k][0].record()
        e2e_step()
        e2e_ev[k][1].record()
    barrier()
    e2e_total = max_over_ranks(float(np.sum([a.elapsed_time(b) for a, b in e2e_ev])))
    e2e_value = E / (e2e_total / args.steps * 1e-3)

    # parity spot check of what was just timed (full check lives in tests/): first 2000 edges vs the C oracle
    checked = None
    if rank == 0:
        from oracle.c_port import bfc_paper_c
        k = min(E, 2000)
        ref = bfc_paper_c(rowptr, col, esrc[:k], edst[:k], 1)
        checked = all(np.array_equal(res[key][:k].cpu().numpy(), ref[key]) for key in ("tri", "sq_i", "sq_j", "gamma", "bfc"))

    # kernels of one pass (dense mode, n <= 262144): degree, node_s, classify, split_zero, plan_ranges, plan_groups, order;
    # paper_group_kernel + paper_light_warp_kernel; paper_value_kernel; unshard at N > 1
    launches_per_step = 7 + 2 + 1 + (1 if world > 1 else 0)
    line = {
        "metric": "bfc_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32 counts + f64 value", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "nodes": n, "undirected_edges": E,
                   "sharding": f"edge e -> rank e % {world}, graph replicated, one all-gather" if world > 1 else "single GPU",
                   "l2": "256 MiB memset between steps (outside the per-step CUDA events)",
                   "timing": "K steps enqueued back to back, per-step CUDA events, sum over steps, max over ranks", "wall_s_timed_region": wall,
                   "parity_spot_check_vs_c_oracle": checked, "step_ms": [round(x, 3) for x in step_ms],
                   "extra_untimed_steps_after_warmup": settle,
                   "step_ms_median_rank0": round(float(np.median(step_ms)), 4),
                   "phase_ms_rank0": {"plan": round(float(np.mean(plan_ms)), 4), "edge_kernels": round(edge_ms_avg, 4),
                                      "gather_unshard": round(float(np.mean(tail_ms)), 4)}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_total / args.steps,
                "what": "dcr.dist.ShardedPaperBFC.run on host (pinned) CSR + edge list; results copied back to pinned host memory"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {
            "bound": "hbm", "kernel": "paper_group_kernel + paper_light_warp_kernel (the two edge kernels of one step, concurrent)",
            "achieved": b_gather_rank / (edge_ms_avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": b_gather_rank / (edge_ms_avg * 1e-3) / 1e9 / peak, "peak_source": peak_src,
            "traffic": NCU_DRAM_BYTES_PER_PASS if (world == 1 and args.workload == "arxiv") else None,
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum summed over the two edge-kernel "
                              "launches of one pass (profiles/r01_v20_ncu_full_edge_kernels.csv)",
            "edge_kernels_ms": edge_ms_avg, "algorithmic_bytes": b_gather_rank,
            "b_gather_total": b_gather_total, "b_compulsory": b_compulsory,
            "note": "algorithmic bytes = SURVEY.md §8d B_gather of this rank's edges (every 2-hop list once per edge, "
                    "no cross-edge reuse charged); the CSR is L2-resident so DRAM traffic is far below it — see "
                    "profiles/ for dram__bytes and lts__t_bytes",
        },
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        thr = host_threads()
        rate, m, dt = cpu_bfc_rate(rowptr, col, esrc, edst, thr, args.cpu_budget)
        line["cpu_baseline"] = {"value": rate, "unit": "edges/s", "cores": thr, "kind": "port",
                                "sample": f"oracle C port of curvature/bfc_naive.py (oracle/c/bfc_paper_csr.c), {thr} "
                                          f"pthreads, seeded uniform sample of {m} of {E} edges, {dt:.1f} s"}
    if rank == 0 and not args.no_sdrf:
        line["sdrf"] = sdrf_bench(args, torch)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries exactly one JSON line: anything libraries print to fd 1 (e.g. NCCL's version banner) goes to
    stderr instead; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_dense(args):
    """Config 4: the dense-regime tensor path (1 GPU).  value = undirected edges/s of the full cuda-flavour pass."""
    import torch

    import __graft_entry__
    __graft_entry__.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    torch.cuda.set_device(0)
    from dcr import bfc
    ei, n, rowptr, col, esrc, edst = build_graph("squirrel")
    E = int(esrc.size)
    csr = bfc.DeviceCSR.from_host(rowptr, col)
    n_pad = (n + 127) // 128 * 128
    ws = torch.empty(int(bfc.L.load().dcr_bfc_support_tc_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
    tri = torch.empty(csr.nnz, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        out = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return float(np.mean(out))

    sampler = ClockSampler(0)
    sampler.start()
    timed(lambda: bfc.support_tc(csr, out=tri, workspace=ws), 1, args.warmup)
    for _ in range(4):
        sampler.probe()
        bfc.support_tc(csr, out=tri, workspace=ws)
    ms_tc = timed(lambda: bfc.support_tc(csr, out=tri, workspace=ws), args.steps, args.warmup)
    ws2 = torch.empty(int(bfc.L.load().dcr_bfc_cuda_flavour_tc_workspace_bytes(n, csr.nnz)), dtype=torch.uint8,
                      device="cuda")
    ms_full = timed(lambda: bfc.cuda_flavour_tc(csr, want_fields=False, workspace=ws2), args.steps, args.warmup)
    sampler.under_load(lambda: bfc.cuda_flavour_tc(csr, want_fields=False, workspace=ws2), args.steps)
    clocks = sampler.result()
    ms_sparse = timed(lambda: bfc.support(csr, out=tri), args.steps, args.warmup)
    ms_full_sparse = timed(lambda: bfc.cuda_flavour(csr, want_fields=False, tri=bfc.support(csr, out=tri)),
                           args.steps, args.warmup)
    same = bool(torch.equal(bfc.support_tc(csr), bfc.support(csr)))
    c_tc = bfc.cuda_flavour_tc(csr, want_fields=False)["c32"]
    c_sp = bfc.cuda_flavour(csr, want_fields=False)["c32"]
    same = same and bool(torch.equal(c_tc.view(torch.int32), c_sp.view(torch.int32)))
    ops = 2.0 * n_pad ** 3
    line = {
        "metric": "bfc_edges_per_sec", "value": E / (ms_full * 1e-3), "unit": "edges/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_full, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int8 x int8 -> int32 (tcgen05 kind::i8), f64 closing formula",
        "data": "synthetic", "config": {"workload": DENSE_WORKLOAD, "nodes": n, "undirected_edges": E, "n_pad": n_pad,
                                        "l2": "256 MiB memset between steps", "outputs_bit_identical_to_sparse_path": same},
        "clocks": clocks, "gpu_launches": 5 * args.steps,   # fill, GEMM, fill_q, GEMM, closing (+ one memset)
        "roofline": {"bound": "tensor", "kernel": "tc_support_kernel (+ memset and tc_fill_kernel of the same call)",
                     "achieved": ops / (ms_tc * 1e-3) / 1e12, "peak": 4500.0, "unit": "TOP/s (int8)",
                     "frac": ops / (ms_tc * 1e-3) / 1e12 / 4500.0,
                     "peak_source": "nominal dense int8 (no measured int8 peak in MEASURED_PEAKS.json)",
                     "traffic": None, "support_tc_ms": ms_tc, "ops": ops},
        "sparse_path": {"support_ms": ms_sparse, "full_ms": ms_full_sparse,
                        "note": "sorted-list intersection kernels on the same CSR (dcr_bfc_support + dcr_bfc_cuda_flavour)"},
    }
    emit(line)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="arxiv", choices=sorted(WORKLOADS) + ["squirrel-dense"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--sdrf-loops", type=int, default=1000)
    ap.add_argument("--sdrf-cpu-iters", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sdrf", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="debug: do not sample clocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "squirrel-dense":
        if args.impl == "reference":
            args.workload = "squirrel"
            run_reference(args)
        else:
            run_dense(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
