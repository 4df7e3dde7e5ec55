#!/usr/bin/env python
"""Small runs of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py
Covers the plain tiled pipeline (one and two masks, E > 32), the symmetric pipeline (symmetric GEMM with its TMA-store
staging, row pass, column pass fed by TMA bricks and by cp.async, transposed-copy variant, fp16 block), the
classifier kernel, the decimal shrink and the batched SVM cross-validation."""
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
    print("plain ok", V, V2, E, float(K.abs().max()), flush=True)
# symmetric pipeline: 3 passes of 256 rows + ragged tail, every variant of the column voxels' sums
V, T, E, eps = 800, 24, 32, 8
raw, labels = synthetic.make_epochs(V, T, E, seed=3)
ep, T_e = engine.stack_epochs(raw, dev)
op = engine.pack_epochs(ep, T_e, "fp32")
ref = None
for name, fl in (("cols", 0), ("cols umma", _lib.FLAG_COLS_UMMA), ("cols v2", _lib.FLAG_COLS_V2), ("tma", _lib.FLAG_COLS_TMA), ("transposed", _lib.FLAG_SYM_TRANSPOSED),
                 ("f16", _lib.FLAG_F16_INTERMEDIATE), ("f16 tma", _lib.FLAG_F16_INTERMEDIATE | _lib.FLAG_COLS_TMA)):
    K = torch.zeros((V, E, E), device=dev)
    work = engine.SymWorkspace(E, V, 256, dev)
    engine.voxel_kernels_sym(op, 0, V, eps, flags=fl | _lib.FLAG_MASK_SELF, work=work, out=K)
    torch.cuda.synchronize()
    ref = K if ref is None else ref
    print("sym ok", name, float((K - ref).abs().max() / ref.abs().max()), flush=True)
# E <= 16: 16-epoch row and column kernels
raw16, _ = synthetic.make_epochs(V, T, 12, seed=4)
ep16, T16 = engine.stack_epochs(raw16, dev)
op16 = engine.pack_epochs(ep16, T16, "fp32")
K16 = torch.zeros((V, 12, 12), device=dev)
engine.voxel_kernels_sym(op16, 0, V, 4, flags=_lib.FLAG_MASK_SELF, work=engine.SymWorkspace(12, V, 256, dev), out=K16)
torch.cuda.synchronize()
print("sym ok E=12 (16-epoch kernels)", float(K16.abs().max()), flush=True)
# 32 < E <= 64: the 64-epoch column kernel (8-column strips, ragged last strip: 812 % 8 = 4)
raw40, _ = synthetic.make_epochs(812, 24, 40, seed=5)
ep40, T40 = engine.stack_epochs(raw40, dev)
op40 = engine.pack_epochs(ep40, T40, "fp32")
K40 = torch.zeros((812, 40, 40), device=dev)
engine.voxel_kernels_sym(op40, 0, 812, 8, flags=_lib.FLAG_MASK_SELF | _lib.FLAG_COLS_WIDE, work=engine.Workspace(40, 812, 256, dev), out=K40)
torch.cuda.synchronize()
print("sym ok E=40 (64-epoch column kernel)", float(K40.abs().max()), flush=True)
Kc = engine.classifier_kernel(op, op, 0, V, eps)
engine.shrink_kernels_(ref)
acc = engine.svm_cv_precomputed(ref[:64], labels, E // eps)
torch.cuda.synchronize()
print("classifier / shrink / svm cv ok", float(Kc.abs().max()), float(np.mean(acc)), flush=True)
# the shrinking solver on problems long enough to shrink, swap and reconstruct; and the one-vs-one decisions
rs = np.random.RandomState(3)
Zs = rs.randn(24, 48, 20).astype(np.float32)
Ks = torch.from_numpy(np.einsum('vej,vfj->vef', Zs, Zs).astype(np.float32)).to(dev)
engine.shrink_kernels_(Ks)
acc_s, it_s = engine.svm_cv_precomputed(Ks, [e % 2 for e in range(48)], 3, shrinking=True, return_iters=True)
acc_m = engine.svm_cv_precomputed(Ks, [e % 3 for e in range(48)], 3, shrinking=True)
torch.cuda.synchronize()
print("svm cv with shrinking / three classes ok", float(np.mean(acc_s)), int(it_s.max()), float(np.mean(acc_m)), flush=True)
