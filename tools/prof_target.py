#!/usr/bin/env python
"""One launch of each hot kernel at the headline shape (for ncu):  python tools/prof_target.py [prec] [nb]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib  # noqa: E402
from brainiak_b200.fcma import engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
V, T, E, eps = 50000, 200, 32, 8
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
rows = engine.pack_epochs(ep, None, prec)
work = engine.Workspace(E, V, nb, dev)
K = torch.empty((nb, E, E), device=dev)
for _ in range(2):
    engine.voxel_kernels(rows, rows, 0, nb, eps, flags=flags, work=work, out=K)
torch.cuda.synchronize()
print("done", prec, nb, float(K.abs().max()))
