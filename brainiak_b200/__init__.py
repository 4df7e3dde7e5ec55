"""brainiak_b200 — Blackwell-native engine for the FCMA correlation hot path of brainiak/brainiak.

Only what that path needs: ``csrc/`` (CUDA kernels + C ABI -> libfcma_b200.so), ``_lib`` (ctypes
binding) and ``fcma/`` (host-side mirror of ``brainiak.fcma``).  No CPU fallback exists.
"""
__version__ = "0.1.0"
