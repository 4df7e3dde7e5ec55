#!/usr/bin/env python
"""Per-kernel times of the fused pipeline (timing hook of the C library): python tools/ab_pipeline.py [nb]"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine
lib = _lib.load()
V, T, E, eps = 50000, 200, 32, 8
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
work = engine.Workspace(E, V, nb, dev)
K = torch.empty((nb, E, E), device=dev)
for prec in ("bf16", "fp16x3", "tf32x3"):
    op = engine.pack_epochs(ep, None, prec)
    for _ in range(3):
        engine.voxel_kernels(op, op, 0, nb, eps, work=work, out=K)
    torch.cuda.synchronize()
    lib.fcma_timing_enable(1)
    for _ in range(6):
        engine.voxel_kernels(op, op, 0, nb, eps, work=work, out=K)
    torch.cuda.synchronize()
    a, b = ctypes.c_double(0), ctypes.c_double(0)
    n = lib.fcma_timing_read(ctypes.byref(a), ctypes.byref(b))
    lib.fcma_timing_enable(0)
    print("%-7s nb=%d  gemm %.3f ms  syrk %.3f ms  total %.3f ms" % (prec, nb, a.value / n, b.value / n, (a.value + b.value) / n), flush=True)
