"""Recipe (TEST INFRASTRUCTURE ONLY): compile the reference's two numba kernels to PTX, into ``oracle/_ref/``.

    python oracle/build_ref.py            # build container only: needs /root/reference; no GPU needed

What it does — no reference source is copied or edited:
  * ``numba.cuda.jit`` is replaced, for the duration of ``import curvature.bfc_cuda`` from ``/root/reference``, by
    a decorator that only records ``(python function, signature string)``: the module's signature-typed
    ``@cuda.jit(...)`` decorators (``curvature/bfc_cuda.py:11,68-69``) otherwise compile at import time and need
    a CUDA driver, which this container does not have;
  * ``numba.cuda.compile_ptx(function, signature, cc=(9, 0))`` — numba's own NVVM pipeline, i.e. the code a user
    of the reference gets — writes ``oracle/_ref/bfc_cuda_ref.ptx`` (both entry points in one file is not
    possible with ``compile_ptx``: two files) and ``oracle/_ref/manifest.json`` (entry names, numba version).
    numba 0.65 tops out at compute capability 9.0; ``.target sm_90`` PTX is JIT-compiled by the driver for
    sm_100 on the GPU box (``oracle/ref_gpu.py`` loads it with ``cuModuleLoadData``).

``oracle/_ref/`` is git-ignored (derived from the reference) but travels to the GPU box with ``gpurun``.
"""
from __future__ import annotations

import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("DCR_REFERENCE", "/root/reference")


def build(force: bool = False) -> bool:
    """Returns True when ``oracle/_ref`` holds both PTX files afterwards."""
    manifest = os.path.join(OUT, "manifest.json")
    src = os.path.join(REFERENCE, "curvature", "bfc_cuda.py")
    if not os.path.exists(src):
        return os.path.exists(manifest)
    if (not force and os.path.exists(manifest)
            and os.path.getmtime(manifest) >= max(os.path.getmtime(src), os.path.getmtime(__file__))):
        return True
    os.environ.pop("NUMBA_ENABLE_CUDASIM", None)
    import numba
    from numba import cuda

    captured = []
    real_jit = cuda.jit

    def recording_jit(signature):
        def deco(fn):
            captured.append((fn, signature))
            return fn
        return deco

    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "curvature" or k.startswith("curvature.")}
    for k in saved_mods:
        del sys.modules[k]
    cuda.jit = recording_jit
    try:
        sys.path.insert(0, REFERENCE)
        import curvature.bfc_cuda  # noqa: F401  (the unmodified reference module)
    finally:
        cuda.jit = real_jit
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "curvature" or k.startswith("curvature.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)

    os.makedirs(OUT, exist_ok=True)
    entries = {}
    for fn, signature in captured:
        ptx, _ = cuda.compile_ptx(fn, signature, cc=(9, 0))
        (entry,) = re.findall(r"\.visible \.entry (\w+)", ptx)
        fname = f"{fn.__name__.lstrip('_')}.ptx"
        with open(os.path.join(OUT, fname), "w") as f:
            f.write(ptx)
        entries[fn.__name__] = {"file": fname, "entry": entry, "signature": signature,
                                "n_params": len(re.findall(r"^\s*\.param ", ptx, flags=re.M))}
    with open(manifest, "w") as f:
        json.dump({"numba": numba.__version__, "cc": [9, 0], "source": "curvature/bfc_cuda.py (unmodified)",
                   "kernels": entries}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference not available")
