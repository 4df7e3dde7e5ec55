#!/usr/bin/env python
"""Where does the 8e-5 kernel difference at BASELINE configs[3] (V=100 000, T=500, E=64) come from?  16 rows through
 - the unmodified reference (fp32 ssyrk),  - the oracle's float64 kernel matrices of the reference-exact z values,
 - the GPU plain pipeline,  - the GPU symmetric pipeline.      python tools/r2_e64_probe.py [V T E]"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine
from oracle import fcma_oracle as orc
V, T, E = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (100000, 500, 64)
eps = 8
dev = torch.device("cuda:0")
ep = bench.device_epochs(V, T, E, dev, seed=bench.SEED + 17 * E + T)
op = engine.pack_epochs(ep, None, "fp16x3")
s0, n0 = (V // 2 // 256) * 256 + 32, 16
Kplain = engine.voxel_kernels(op, op, s0, n0, eps).cpu().numpy()
Ksym = torch.zeros((V, E, E), device=dev)
cols = bool(_lib.load().fcma_sym_uses_column_pass(_lib.PREC["fp16x3"], E, eps, 0))
work = engine.SymWorkspace(E, V, 1024, dev, transposed_copy=not cols)
engine.voxel_kernels_sym(op, 0, V, eps, work=work, out=Ksym)
Ksym = Ksym[s0:s0 + n0].cpu().numpy()
host = ep.cpu()
raw = [host[e].numpy() for e in range(E)]
bench.host_threads()
m, rvs, clf, labels = bench.reference_selector(raw, eps)
corr = rvs._correlation_computation((s0, n0))
m.fcma_extension.normalization(corr, eps)
z = corr.copy()
K64 = orc.kernel_matrices(z, f64=True)              # float64 Gram matrices of the reference's own z values
Kref = np.zeros((n0, E, E), np.float32)
# the reference's ssyrk, WITHOUT the decimal shrink
from oracle import reference
blas = reference.load().cython_blas if hasattr(reference.load(), "cython_blas") else None
Kref_s = rvs._prepare_for_cross_validation(corr, clf)     # shrunk
nd = np.array([len(str(int(K64[i, 0, 0]))) for i in range(n0)])
scale = np.array([10.0 ** (2 - d) if d > 2 else 1.0 for d in nd], np.float64)[:, None, None]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


print("shape V=%d T=%d E=%d rows [%d, %d)" % (V, T, E, s0, s0 + n0))
print("reference fp32 ssyrk (shrunk) vs float64 Gram of its own z :", rel(Kref_s.astype(np.float64), K64 * scale))
print("GPU plain pipeline            vs float64 Gram of ref z     :", rel(Kplain.astype(np.float64), K64))
print("GPU symmetric pipeline        vs float64 Gram of ref z     :", rel(Ksym.astype(np.float64), K64))
print("GPU symmetric vs GPU plain                                 :", rel(Ksym, Kplain))
print("GPU symmetric (shrunk) vs reference (shrunk)               :", rel(Ksym.astype(np.float64) * scale, Kref_s.astype(np.float64)))
