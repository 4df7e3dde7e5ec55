#!/usr/bin/env python
"""Bottleneck diagnosis of the pair GEMM: time it with parts of the epilogue switched off
(FCMA_GEMM_DEBUG bit mask: 1 no stores, 2 no TMEM load/math/fill, 4 no epilogue work,
8 one extra tcgen05.commit per MMA segment) in the
resident and the streaming variant.  Outputs are WRONG while a bit is set -- timing only."""
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
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for prec in sys.argv[1:] or ("bf16", "fp16x3"):
    rows = engine.pack_epochs(ep, None, prec)
    for nores in ("0", "1"):
        os.environ["FCMA_GEMM_RESIDENT"] = nores
        for dbg in (0, 4, 12):
            os.environ["FCMA_GEMM_DEBUG"] = str(dbg)
            ms = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld))
            msf = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld, fisher_epochs=E))
            print(f"{prec:7s} resident={'yes' if nores=='1' else 'no '} debug={dbg}: gemm {ms:7.3f} ms   +fisher {msf:7.3f} ms", flush=True)
    del rows
os.environ["FCMA_GEMM_DEBUG"] = "0"
