"""lobpcg_b200 — B200-native (sm_100a) LOBPCG hot path behind the reference's solver/operator interface.

Layout: ``csrc/`` CUDA kernels + C ABI (built in-tree by ``lobpcg_b200.build``), ``api.py`` ctypes mirror
of the reference interface, ``problems.py`` synthetic problem generators, ``dist.py`` multi-GPU plumbing.
"""
from . import problems  # noqa: F401

__all__ = ["problems"]
