#!/usr/bin/env python
"""Round-2 A/B of the symmetric step at the bench shape (diagnostic build): per-kernel CUDA-event times of
fcma_voxel_kernels_sym for flag / environment variants, alternating, 2 rounds.
   python tools/r2_ab.py [V] [variant ...]        variant = name:flags:ENV=val,ENV=val"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib, build as _build
_build.build(diag=True)
_lib.use_diag_build()
from brainiak_b200.fcma import engine
lib = _lib.load()
V = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
T, E, eps, rows = 200, 32, 8, 4096
default = ["base:0:", "cols_v2:%d:" % _lib.FLAG_COLS_V2, "cols_tma:%d:" % _lib.FLAG_COLS_TMA, "gemm_tma_transposed:0:FCMA_GEMM_DEBUG=128",
           "gemm_no_epilogue:0:FCMA_GEMM_DEBUG=4", "f16_block:%d:" % _lib.FLAG_F16_INTERMEDIATE]
variants = []
for a in (sys.argv[2:] or default):
    name, fl, env = a.split(":")
    variants.append((name, int(fl), dict(kv.split("=") for kv in env.split(",") if kv)))
keys = sorted({k for v in variants for k in v[2]})
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
op = engine.pack_epochs(ep, None, "fp16x3")
work = engine.SymWorkspace(E, V, rows, dev, transposed_copy=False)
K = torch.zeros((V, E, E), device=dev)


def step(fl):
    K.zero_()
    engine.voxel_kernels_sym(op, 0, V, eps, flags=fl, work=work, out=K)


ref = None
for rep in range(2):
    for name, fl, env in variants:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        step(fl)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            step(fl)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        lib.fcma_timing_enable(1)
        step(fl)
        torch.cuda.synchronize()
        x, y, z = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        lib.fcma_timing_read3(ctypes.byref(x), ctypes.byref(y), ctypes.byref(z))
        lib.fcma_timing_enable(0)
        if ref is None:
            ref = K.clone()
        err = float((K - ref).abs().max() / ref.abs().max())
        print("%-22s step %.1f ms | gemm %.1f  rows %.1f  cols %.1f ms (timed per pass) | max|dK|/max|K| vs first %.1e" % (name, ms, x.value, y.value, z.value, err), flush=True)
for k in keys:
    os.environ.pop(k, None)
