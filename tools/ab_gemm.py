#!/usr/bin/env python
"""A/B timing of GEMM scheduling variants inside one process (same GPU, alternating)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib
from brainiak_b200 import build as _build  # noqa: E402
_build.build(diag=True)      # the FCMA_* knobs exist only in the diagnostic build (-DFCMA_DIAG)
_lib.use_diag_build()
from brainiak_b200.fcma import engine
V, T, E, eps, nb = 50000, 200, 32, 8, 2048
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
work = engine.Workspace(E, V, nb, dev)
ld = ((V + 31) // 32) * 32
cbuf = work.buf.view(torch.float32)[: nb * E * ld].view(nb, E, ld)
K = torch.empty((nb, E, E), device=dev)
def timeit(fn, n=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
variants = [a.split("=") for a in sys.argv[1:]] or [["FCMA_GEMM_SCHED", "0"], ["FCMA_GEMM_SCHED", "1"]]
for prec in ("tf32x3", "bf16x3", "bf16"):
    rows = engine.pack_epochs(ep, None, prec)
    res = {}
    for rep in range(3):
        for k, v in variants:
            os.environ[k] = v
            ms = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld))
            msf = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld, fisher_epochs=E))
            res.setdefault((k, v), []).append((ms, msf))
    for kv, r in res.items():
        print(prec, kv, " gemm ms:", ["%.3f" % a for a, _ in r], " +fisher:", ["%.3f" % b for _, b in r], flush=True)
