"""discrete-curvature-rewiring_b200 — B200 (sm_100a) implementation of the BFC / SDRF hot path.

The directory name is not a Python identifier; it is meant to be put on ``sys.path`` (in front of the reference
checkout) so that ``curvature.bfc_cuda``, ``curvature.bfc_naive``, ``rewiring.sdrf_cuda_bfc``,
``rewiring.rewire`` and ``utils.softmax`` resolve to the modules in here, and ``dcr`` to the core package
(ctypes binding of ``libdcr.so``).  See INTEGRATION.md.
"""
