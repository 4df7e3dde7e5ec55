#!/usr/bin/env python
"""One pass of the symmetric pipeline at the headline shape (for ncu): rows [start, start+nb) against columns
[start, V): one k_corr_umma2 launch (symmetric mode) + k_norm_syrk over the block and over its transposed copy.
python tools/prof_target_sym.py [prec] [nb] [flags] [start] [V T E]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib  # noqa: E402,F401
from brainiak_b200.fcma import engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
start = int(sys.argv[4]) if len(sys.argv) > 4 else 0
V, T, E, eps = 50000, 200, 32, 8
if len(sys.argv) > 7:
    V, T, E = int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
op = engine.pack_epochs(ep, None, prec)
work = engine.SymWorkspace(E, V, nb, dev, start=start)
K = torch.zeros((V, E, E), device=dev)
for _ in range(2):
    engine.voxel_kernels_sym(op, start, nb, eps, flags=flags, work=work, out=K)
torch.cuda.synchronize()
print("done", prec, nb, flags, start, float(K.abs().max()))
