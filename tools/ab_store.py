#!/usr/bin/env python
"""GEMM epilogue A/B: TMA-store (default) vs direct STG (FCMA_GEMM_NO_TMA_STORE=1), strided block."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200.fcma import engine
V, T, E, nb = 50000, 200, 32, 4096
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
work = engine.Workspace(E, V, nb, dev)
ld = ((V + 31) // 32) * 32
cbuf = work.buf.view(torch.float32)[: nb * E * ld].view(nb, E, ld)
def timeit(fn, n=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for prec in sys.argv[1:] or ("fp16x3", "bf16"):
    rows = engine.pack_epochs(ep, None, prec)
    res = {}
    for rep in range(3):
        for v in ("0", "1"):
            os.environ["FCMA_GEMM_NO_TMA_STORE"] = v
            ms = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld))
            msf = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld, fisher_epochs=E))
            res.setdefault(v, []).append((ms, msf))
    for v, r in res.items():
        print(prec, "STG      " if v == "1" else "TMA store", " gemm ms:", ["%.3f" % a for a, _ in r], " +fisher:", ["%.3f" % b for _, b in r], flush=True)
    del rows
os.environ["FCMA_GEMM_NO_TMA_STORE"] = "0"
