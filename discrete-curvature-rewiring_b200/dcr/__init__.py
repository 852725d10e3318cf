"""Core host package: ctypes binding of libdcr.so (hand-written sm_100a kernels) + graph set-up helpers.

There is no CPU fallback: importing :mod:`dcr.lib` raises when ``libdcr.so`` has not been built, and every
compute entry point requires a CUDA device.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
