#!/usr/bin/env python
"""BASELINE configs[3] shape on one GPU (one row block): V=100 000, T=500, E=64, eps=8."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine
lib = _lib.load()
V, T, E, eps, nb = 100000, 500, 64, 8, 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.empty((E, T, V), device=dev)
for e in range(E):
    ep[e] = torch.randn((T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
op = engine.pack_epochs(ep, None, "fp16x3")
del ep
work = engine.Workspace(E, V, nb, dev)
K = torch.empty((nb, E, E), device=dev)
fl = _lib.FLAG_MASK_SELF
for _ in range(2):
    engine.voxel_kernels(op, op, 1000, nb, eps, flags=fl, work=work, out=K)
torch.cuda.synchronize()
lib.fcma_timing_enable(1)
for _ in range(3):
    engine.voxel_kernels(op, op, 1000, nb, eps, flags=fl, work=work, out=K)
torch.cuda.synchronize()
a, b = ctypes.c_double(0), ctypes.c_double(0)
n = lib.fcma_timing_read(ctypes.byref(a), ctypes.byref(b))
lib.fcma_timing_enable(0)
corr = nb * V * E
tot = (a.value + b.value) / n
print("config4 block: gemm %.2f ms  syrk %.2f ms  -> %.3e corr/s (1 GPU);  HBM: gemm %.0f GB/s write, syrk %.0f GB/s read"
      % (a.value / n, b.value / n, corr / tot * 1e3, corr * 4 / (a.value / n) / 1e6, corr * 4 / (b.value / n) / 1e6))
Kd = K.double()
tr = torch.diagonal(Kd, dim1=1, dim2=2).sum(1)
print("invariants: trace/E/(V-1) in [%.6f, %.6f]; subject-block row sums max %.3e; symmetric %s"
      % (float((tr / (E * (V - 1.0))).min()), float((tr / (E * (V - 1.0))).max()),
         float(Kd.view(nb, E // eps, eps, E).sum(2).abs().max()), bool((K == K.transpose(1, 2)).all())))
