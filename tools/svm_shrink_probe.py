"""GPU SMO solver with libsvm's shrinking heuristic vs scikit-learn, problem by problem: iteration counts (SVC.n_iter_) and
held-out decisions.  Usage: python tools/svm_shrink_probe.py [nv]"""
import sys
import numpy as np
import torch
from numpy.random import RandomState
from sklearn import svm, model_selection
sys.path.insert(0, ".")
from brainiak_b200.fcma import engine
from brainiak_b200.fcma.voxelselector import shrink_kernels_

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
rng = RandomState(5)
for (E, folds, C, tol, T, sig, lab) in (
        (32, 4, 1.0, 1e-3, 200, 0.35, None), (32, 8, 0.05, 1e-3, 200, 0.35, None), (32, 2, 10.0, 1e-4, 200, 0.35, None),
        (64, 4, 1.0, 1e-3, 60, 0.2, None), (64, 2, 100.0, 1e-3, 30, 0.1, None), (48, 3, 1.0, 1e-3, 20, 0.3, None),
        (36, 3, 1.0, 1e-3, 120, 0.3, [e % 3 for e in range(36)])):
    Z = rng.randn(nv, E, T).astype(np.float32)
    lab = lab if lab is not None else [e % 2 for e in range(E)]
    code = np.asarray(lab)
    for c in np.unique(code):
        Z[:, code == c, 5 * c:5 * c + 5] += sig
    K = np.einsum('vej,vfj->vef', Z, Z).astype(np.float32)
    shrink_kernels_(K)
    Kd = torch.from_numpy(K).to(dev)
    for shrinking in (False, True):
        got, iters = engine.svm_cv_precomputed(Kd, lab, folds, C=C, tol=tol, return_iters=True, shrinking=shrinking,
                                               max_iter=500000)
        fd = engine.make_svm_folds(lab, folds)
        npairs = len(fd.pairs)
        skf = model_selection.StratifiedKFold(n_splits=folds, shuffle=False)
        y = np.asarray(lab)
        bad_it = bad_acc = 0
        total = 0
        ref_acc = np.zeros(nv)
        for v in range(nv):
            accs = []
            for f, (tr, te) in enumerate(skf.split(np.zeros((E, 1)), y)):
                clf = svm.SVC(kernel="precomputed", C=C, tol=tol, shrinking=shrinking)
                clf.fit(K[v][np.ix_(tr, tr)].astype(np.float64), y[tr])
                ref_it = np.asarray(clf.n_iter_).ravel()
                total += len(ref_it)
                bad_it += int(np.sum(ref_it != iters[v, f * npairs:(f + 1) * npairs]))
                accs.append(np.mean(clf.predict(K[v][np.ix_(te, tr)].astype(np.float64)) == y[te]))
            ref_acc[v] = np.mean(accs)
        bad_acc = int(np.sum(ref_acc != got))
        print(f"E={E} folds={folds} C={C} tol={tol} classes={len(np.unique(y))} shrinking={shrinking}: "
              f"iteration counts differ in {bad_it}/{total} problems, accuracies differ for {bad_acc}/{nv} voxels; "
              f"iters max {iters.max()} mean {iters.mean():.1f}", flush=True)
