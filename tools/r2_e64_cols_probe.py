#!/usr/bin/env python
"""32 < E <= 64: column pass over the stored block (k_norm_syrk_cols64) against the transposed copy + row pass, per-kernel
times of one symmetric step (CUDA events inside the C pipeline).   python tools/r2_e64_cols_probe.py [V T E]"""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine
V, T, E = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (40000, 64, 64)
eps = 8
dev = torch.device("cuda:0")
lib = _lib.load()
ep = bench.device_epochs(V, T, E, dev, seed=7)
op = engine.pack_epochs(ep, None, "fp16x3")
K = torch.zeros((V, E, E), device=dev)
ref = None
for name, flags, rows, transposed in (("transposed copy, 1280 rows", 0, 1280, True),
                                      ("transposed copy, 2560 rows", 0, 2560, True),
                                      ("column pass, 1280 rows", _lib.FLAG_COLS_WIDE, 1280, False),
                                      ("column pass, 2560 rows", _lib.FLAG_COLS_WIDE, 2560, False),
                                      ("column pass, 4096 rows", _lib.FLAG_COLS_WIDE, 4096, False)):
    work = engine.SymWorkspace(E, V, rows, dev, transposed_copy=transposed)
    for rep in range(2):
        K.zero_()
        if rep == 1:
            lib.fcma_timing_enable(1)
        engine.voxel_kernels_sym(op, 0, V, eps, flags=flags, work=work, out=K)
        torch.cuda.synchronize()
    g, s, s2 = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
    npass = lib.fcma_timing_read3(ctypes.byref(g), ctypes.byref(s), ctypes.byref(s2))
    lib.fcma_timing_enable(0)
    if ref is None:
        ref = K.clone()
    err = float((K - ref).abs().max() / ref.abs().max())
    print(f"{name:36s}: passes {npass:3d}  GEMM {g.value:8.1f} ms  rows {s.value:8.1f} ms  columns {s2.value:8.1f} ms  "
          f"total {g.value + s.value + s2.value:8.1f} ms   max|dK|/max|K| vs first {err:.1e}", flush=True)
    del work
