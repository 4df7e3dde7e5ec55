#!/usr/bin/env python
"""Plain pipeline (row kernel) against the oracle for 32 < E <= 64 and every eps: where do the kernels differ?"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine, synthetic
from oracle import fcma_oracle as orc
dev = torch.device("cuda:0")
V, T = 1300, 24
for E, eps in ((64, 2), (64, 4), (64, 8), (48, 2), (40, 4), (32, 2), (64, 64), (64, 32)):
    raw, _ = synthetic.make_epochs(V, T, E, seed=2000 + 41 * E + eps)
    ep, T_e = engine.stack_epochs(raw, dev)
    op = engine.pack_epochs(ep, T_e, "fp32")
    plain = engine.voxel_kernels(op, op, 700, 30, eps, flags=_lib.FLAG_MASK_SELF).cpu().numpy()
    _, z, _ = orc.voxel_block(raw, None, 700, 30, eps, shrink=False)
    for i in range(30):
        z[i, :, 700 + i] = 0
    Kref = orc.kernel_matrices(z, f64=True)
    d = np.abs(plain - Kref)
    v, a, b = np.unravel_index(np.argmax(d), d.shape)
    # how many columns of that voxel have z == 0 in the oracle for epochs a, b
    nz = int(np.sum(z[v, a] == 0)), int(np.sum(z[v, b] == 0))
    print(f"E={E} eps={eps}: max|dK| {d.max():.3f} at voxel {700 + v} entry ({a},{b}): gpu {plain[v, a, b]:.3f} oracle {Kref[v, a, b]:.3f}; "
          f"oracle zeros in those epochs {nz}; per-voxel max {np.round(d.reshape(30, -1).max(1)[:8], 2)}", flush=True)
