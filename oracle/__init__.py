"""CPU oracle for the BFC / SDRF hot path — TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms of the reference (jakubbober/discrete-curvature-rewiring)
that the CUDA library replaces.  Every function cites the reference ``file:line`` it follows.

Rules (DESIGN.md §"Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
    legs may import, call, link or execute anything under ``oracle/`` — as the checker or the timed CPU
    baseline, never as the product path.  Nothing under ``discrete-curvature-rewiring_b200/`` imports it.
  * pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself,
    SURVEY.md §4/§8c).  The oracle is instead pinned against outputs of the UNMODIFIED reference modules run in
    the build container (``curvature/bfc_naive.py`` directly; ``curvature/bfc_cuda.py`` and
    ``rewiring/sdrf_cuda_bfc.py`` under ``NUMBA_ENABLE_CUDASIM=1``), committed as ``tests/golden/*.npz`` with
    the generating script ``tests/golden/generate_golden.py``, and against the known answers of SURVEY.md
    Appendix G.  The fp32 bit patterns of the compiled (non-simulator) numba kernel follow its PTX dataflow
    (``numba.cuda.compile_ptx``) — rounding model ``"compiled"`` — and are pinned BY
    EXECUTION: ``oracle/build_ref.py`` compiles the reference's own kernels to PTX (``oracle/_ref/``, derived from
    /root/reference, not committed), ``oracle/ref_gpu.py`` runs them on the GPU through the driver API, and
    ``tests/test_gpu_ref_kernels.py`` checks this oracle and the CUDA product path against them bit for bit.
    The simulator's all-fp32 arithmetic is rounding model ``"sim32"`` and is what the golden files generated
    under the simulator are compared with bit for bit.
"""
