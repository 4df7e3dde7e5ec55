#!/usr/bin/env python
"""Where does the pair GEMM's time go?  Times one launch (4096 rows x 50 000 columns x 32 epochs) with parts
switched off through FCMA_GEMM_DEBUG (outputs are WRONG while a bit is set -- timing only):
   0  full kernel            4  main loop only (TMA loads + MMAs, epilogue warps just hand TMEM back)
  20  MMAs only (stale shared-memory stages are re-read: no loads, no epilogue)
and for T = 64 .. 256 (1 .. 4 k-blocks of 64) to expose per-stage costs.   python tools/gemm_debug.py [prec ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib, build as _build
_build.build(diag=True)      # the FCMA_* knobs exist only in the diagnostic build (-DFCMA_DIAG)
_lib.use_diag_build()
from brainiak_b200.fcma import engine
V, E, nb = 50000, 32, 4096
dev = torch.device("cuda:0")
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
work = engine.Workspace(E, V, nb, dev)
ld = ((V + 31) // 32) * 32
cbuf = work.buf.view(torch.float32)[: nb * E * ld].view(nb, E, ld)
for T in (200, 64, 128, 192, 256):
    g = torch.Generator(device=dev).manual_seed(0)
    ep = torch.randn((E, T, V), device=dev, generator=g)
    engine.epoch_normalize_(ep)
    for prec in sys.argv[1:] or ("fp16x3", "bf16"):
        rows = engine.pack_epochs(ep, None, prec)
        out = []
        for dbg in (0, 4, 20):
            os.environ["FCMA_GEMM_DEBUG"] = str(dbg)
            out.append(min(timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld, fisher_epochs=E)) for _ in range(2)))
        print(f"T={T:3d} {prec:7s}: full(+fisher) {out[0]:7.3f} ms   main loop only {out[1]:7.3f} ms   MMAs only {out[2]:7.3f} ms", flush=True)
        del rows
    del ep
os.environ["FCMA_GEMM_DEBUG"] = "0"
