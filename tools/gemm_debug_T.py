#!/usr/bin/env python
"""Per-stage overhead of the pair GEMM main loop: time it for T = 192 / 200 / 256 (3, 3+tail, 4 k-blocks
of 64) with and without the epilogue (FCMA_GEMM_DEBUG=4).  Timing only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for T in (192, 200, 256, 128, 64):
    g = torch.Generator(device=dev).manual_seed(0)
    ep = torch.randn((E, T, V), device=dev, generator=g)
    engine.epoch_normalize_(ep)
    for prec in ("bf16", "fp16x3"):
        rows = engine.pack_epochs(ep, None, prec)
        for nores in ("1", "0"):
            os.environ["FCMA_GEMM_RESIDENT"] = nores
            out = []
            for dbg in (0, 4):
                os.environ["FCMA_GEMM_DEBUG"] = str(dbg)
                out.append(min(timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld)) for _ in range(2)))
            print(f"T={T:3d} {prec:7s} resident={'yes' if nores=='1' else 'no '}: full {out[0]:7.3f} ms   mainloop-only {out[1]:7.3f} ms", flush=True)
        del rows
    del ep
os.environ["FCMA_GEMM_DEBUG"] = "0"
