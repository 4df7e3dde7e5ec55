#!/usr/bin/env python
"""Generic in-process A/B of the fused pipeline under environment knobs (all knobs are read per launch):
   python tools/ab_env.py <nb> <prec> "KNOB=1,OTHER=0" "KNOB=0" ...
Prints per-kernel CUDA-event times (library timing hook) for each variant, alternating, 3 rounds."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib
from brainiak_b200 import build as _build  # noqa: E402
_build.build(diag=True)      # the FCMA_* knobs exist only in the diagnostic build (-DFCMA_DIAG)
_lib.use_diag_build()
from brainiak_b200.fcma import engine
lib = _lib.load()
V, T, E, eps = 50000, 200, 32, 8
nb = int(sys.argv[1]); prec = sys.argv[2]
variants = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[3:]] or [{}]
keys = sorted({k for v in variants for k in v})
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
work = engine.Workspace(E, V, nb, dev)
K = torch.empty((nb, E, E), device=dev)
op = engine.pack_epochs(ep, None, prec)
ref = None
for rep in range(3):
    for v in variants:
        for k in keys:
            os.environ[k] = v.get(k, "0")
        for _ in range(2):
            engine.voxel_kernels(op, op, 40000, nb, eps, work=work, out=K)
        torch.cuda.synchronize()
        if ref is None: ref = K.clone()
        d = (K - ref).abs().max().item()
        lib.fcma_timing_enable(1)
        for _ in range(5):
            engine.voxel_kernels(op, op, 40000, nb, eps, work=work, out=K)
        torch.cuda.synchronize()
        a, b = ctypes.c_double(0), ctypes.c_double(0)
        n = lib.fcma_timing_read(ctypes.byref(a), ctypes.byref(b))
        lib.fcma_timing_enable(0)
        print("%-7s nb=%d %-44s gemm %.3f ms  syrk %.3f ms  total %.3f ms  max|dK| %.2g" %
              (prec, nb, ",".join(f"{k}={v.get(k,'0')}" for k in keys), a.value / n, b.value / n, (a.value + b.value) / n, d), flush=True)
for k in keys: os.environ.pop(k, None)
