#!/usr/bin/env python
"""Small tiled-pipeline run for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine, synthetic
dev = torch.device("cuda:0")
for (V, V2, T, E, eps, start, nb, rows) in ((700, None, 40, 8, 4, 77, 600, 768), (300, 530, 24, 16, 8, 0, 300, 512), (200, None, 30, 48, 16, 3, 150, 150)):
    raw, _ = synthetic.make_epochs(V, T, E, seed=1)
    raw2 = synthetic.make_epochs(V2, T, E, seed=2)[0] if V2 else None
    ep, T_e = engine.stack_epochs(raw, dev)
    r = engine.pack_epochs(ep, T_e, "fp32")
    c = engine.pack_epochs(engine.stack_epochs(raw2, dev)[0], T_e, r.precision) if V2 else r
    work = engine.Workspace(E, V2 or V, rows, dev)
    K = engine.voxel_kernels(r, c, start, nb, eps, flags=0 if V2 else _lib.FLAG_MASK_SELF, work=work)
    torch.cuda.synchronize()
    print("ok", V, V2, E, float(K.abs().max()))
