"""Minimal stand-in for the parts of ``torch_geometric`` the reference's hot path touches.

PyTorch Geometric is not installed in the build / GPU images, and the reference's
``rewiring/sdrf_cuda_bfc.py:6-7``, ``models/gcn.py:8-9`` and ``experiment/training_loop.py:7`` import it.
This package provides exactly the names those files use, with the semantics of torch-geometric 2.0.3
(SURVEY.md App. E.1).  It is only put on ``sys.path`` when the real package is absent
(``dcr.compat.ensure_torch_geometric()``); a real installation always wins.
"""
from . import data, nn, utils  # noqa: F401

__version__ = "2.0.3+dcr.standin"
