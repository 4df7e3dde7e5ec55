#!/usr/bin/env python
"""VoxelSelector.run at the bench shape for the three kinds of SVC the GPU cross-validation takes: shrinking=False (the
reference's examples), scikit-learn's default shrinking=True, and four conditions (one-vs-one).  Wall clock, host numpy epochs ->
sorted (voxel, accuracy) list; second run of each (the first pays allocations)."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sklearn import svm
from brainiak_b200.fcma.voxelselector import VoxelSelector
V, T, E, eps = 50000, 200, 32, 8
dev = torch.device("cuda:0")
ep = bench.device_epochs(V, T, E, dev, bench.SEED).cpu()          # generated on the device: seconds instead of minutes
raw = [ep[e].numpy() for e in range(E)]
for name, labels, clf in (("two conditions, shrinking=False", [e % 2 for e in range(E)], svm.SVC(kernel="precomputed", shrinking=False, C=1)),
                          ("two conditions, shrinking=True (scikit-learn default)", [e % 2 for e in range(E)], svm.SVC(kernel="precomputed", C=1)),
                          ("four conditions (one-vs-one), shrinking=False", [e % 4 for e in range(E)], svm.SVC(kernel="precomputed", shrinking=False, C=1))):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = VoxelSelector(labels, eps, E // eps, raw, voxel_unit=64, process_num=0).run(clf)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    acc = np.array([a for _, a in res])
    print(f"{name:56s}: {dt:.3f} s   accuracies min/mean/max {acc.min():.3f} / {acc.mean():.3f} / {acc.max():.3f}", flush=True)
