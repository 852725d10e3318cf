"""Environment shims: make ``torch_geometric`` importable (stand-in) when the real package is absent."""
from __future__ import annotations

import importlib.util
import os
import sys

_STANDIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "standin")


def ensure_torch_geometric() -> str:
    """Returns ``"real"`` or ``"standin"``.  A real installation always wins."""
    if "torch_geometric" in sys.modules:
        return "standin" if "standin" in getattr(sys.modules["torch_geometric"], "__version__", "") else "real"
    if importlib.util.find_spec("torch_geometric") is not None:
        return "real"
    if _STANDIN not in sys.path:
        sys.path.append(_STANDIN)
    return "standin"
