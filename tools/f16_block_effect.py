#!/usr/bin/env python
"""What the opt-in fp16 Fisher-z block costs in the RESULT: cross-validation accuracies of all V voxels (GPU SVM CV)
from the fp32-block symmetric pipeline, the plain pipeline (same values, other summation order) and the fp16-block
pipeline; counts voxels whose accuracy differs and the overlap of the top-1% selections.
python tools/f16_block_effect.py [V]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib  # noqa: E402
from brainiak_b200.fcma import engine  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
T, E, eps, folds = 200, 32, 8, 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
ep = torch.randn((E, T, V), device=dev, generator=g)
ep[1::2, :, : V // 100] += 0.6 * torch.randn((E // 2, T, 1), device=dev, generator=g)
engine.epoch_normalize_(ep)
op = engine.pack_epochs(ep, None, "fp16x3")
del ep
labels = [e % 2 for e in range(E)]
fold_desc = engine.make_svm_folds(labels, folds)


def accuracies(K):
    K = K.clone()
    engine.shrink_kernels_(K)
    return engine.svm_cv_precomputed(K, labels, folds, C=1.0, tol=1e-3, folds=fold_desc)


work = engine.SymWorkspace(E, V, 4096, dev)
K32 = engine.voxel_kernels_sym(op, 0, V, eps, work=work)
a32 = accuracies(K32)
K16 = engine.voxel_kernels_sym(op, 0, V, eps, flags=_lib.FLAG_F16_INTERMEDIATE, work=work)
a16 = accuracies(K16)
Kp = torch.empty((V, E, E), device=dev)
engine.voxel_kernels(op, op, 0, V, eps, work=work, out=Kp)
ap = accuracies(Kp)
sc = float(Kp.abs().max())
top = max(1, V // 100)


def report(name, K, a):
    d = float((K - Kp).abs().max()) / sc
    flips = int(np.sum(a != ap))
    t1, t2 = set(np.argsort(-a, kind="stable")[:top]), set(np.argsort(-ap, kind="stable")[:top])
    print("%-28s max|dK|/max|K| vs plain %.3g; accuracies differing from plain: %d of %d (%.3f %%), max |d acc| %.4f; "
          "top-1%% overlap %d / %d" % (name, d, flips, V, 100.0 * flips / V, float(np.max(np.abs(a - ap))), len(t1 & t2), top),
          flush=True)


report("symmetric, fp32 block", K32, a32)
report("symmetric, fp16 block", K16, a16)
print("planted voxels in the top 1%%: plain %d, fp32 block %d, fp16 block %d of %d" % tuple(
    [int(np.sum(np.argsort(-a, kind="stable")[:top] < top)) for a in (ap, a32, a16)] + [top]))
