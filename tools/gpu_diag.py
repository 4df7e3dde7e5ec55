#!/usr/bin/env python
"""GPU diagnostic: every kernel against the CPU oracle on small seeded inputs, error magnitudes
printed, plus quick timings.  Run on the B200 box:  timeout 600 python tools/gpu_diag.py [--big]"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brainiak_b200 import _lib  # noqa: E402
from brainiak_b200.fcma import engine, synthetic  # noqa: E402
from oracle import fcma_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
FAILS = []


def section(name):
    print("\n=== " + name, flush=True)


def report(name, got, ref, tol, rel_to=None):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    err = np.max(np.abs(got - ref))
    scale = np.max(np.abs(ref)) if rel_to is None else rel_to
    ok = np.isfinite(err) and err <= tol * max(scale, 1e-30)
    print("  %-46s max|d|=%.3e  scale=%.3e  tol=%.1e  %s" %
          (name, err, scale, tol, "ok" if ok else "FAIL"), flush=True)
    if not ok:
        FAILS.append(name)
    return ok


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:
        FAILS.append(name)
        print("  EXCEPTION in %s:\n%s" % (name, traceback.format_exc()), flush=True)


def small_case(V=300, V2=None, T=50, E=8, eps=4, start=33, nb=70, seed=1):
    two = V2 is not None
    raw, labels = synthetic.make_epochs(V, T, E, seed=1234 + seed)
    raw2 = synthetic.make_epochs(V2, T, E, seed=777 + seed)[0] if two else None
    return raw, raw2, dict(V=V, V2=V2 if two else V, T=T, E=E, eps=eps, start=start, nb=nb)


def check_corr(raw, raw2, c):
    ep, T_e = engine.stack_epochs(raw, dev)
    ep2 = engine.stack_epochs(raw2, dev)[0] if raw2 is not None else ep
    ref64 = orc.corr_block(raw, raw2, c["start"], c["nb"], f64=True)
    ref32 = orc.corr_block(raw, raw2, c["start"], c["nb"])
    report("oracle f32 vs f64", ref32, ref64, 2e-6, 1.0)

    def simt():
        out = engine.corr_block_f32(ep, ep2, c["start"], c["nb"])
        report("corr_block_f32 (SIMT)", out.cpu().numpy(), ref64, 2e-6, 1.0)
    run("corr simt", simt)
    tols = {"bf16": 8e-3, "tf32": 1e-3, "bf16x3": 4e-5, "tf32x3": 1e-6, "fp16x3": 1e-6}
    for prec in ("bf16", "tf32", "bf16x3", "tf32x3", "fp16x3"):
        def umma(prec=prec):
            rows = engine.pack_epochs(ep, T_e, prec)
            cols = engine.pack_epochs(ep2, T_e, prec) if raw2 is not None else rows
            for layout in (0, 1):
                out = engine.corr_block(rows, cols, c["start"], c["nb"], layout=layout)
                ref = ref64 if layout == 0 else np.transpose(ref64, (1, 0, 2))
                report("corr_block umma %-7s layout %d" % (prec, layout), out.cpu().numpy(), ref,
                       tols[prec], 1.0)
        run("corr umma " + prec, umma)


def check_norm_and_kernels(raw, raw2, c):
    r, z, K = orc.voxel_block(raw, raw2, c["start"], c["nb"], c["eps"], shrink=False)

    def norm():
        t = torch.from_numpy(r.copy()).to(dev)
        engine.within_subject_norm_(t, c["eps"])
        got = t.cpu().numpy()
        nbits = np.mean(got == z)
        print("  within_subject_norm: bit-identical fraction %.6f" % nbits)
        selfmask = np.ones_like(z, bool)
        if raw2 is None:
            for i in range(c["nb"]):
                selfmask[i, :, c["start"] + i] = False
        report("within_subject_norm (off-self)", got[selfmask], z[selfmask], 5e-4, 1.0)
    run("norm", norm)

    def syrk():
        zt = torch.from_numpy(z).to(dev)
        got = engine.kernel_matrices(zt).cpu().numpy()
        K64 = orc.kernel_matrices(z, f64=True)
        report("kernel_matrices (tf32 mma) vs f64", got, K64, 2e-4)
        report("oracle K f32 vs f64", K, K64, 1e-5)
        Ks = engine.kernel_matrices(zt, sum_over_rows=True).cpu().numpy()
        report("kernel_matrices sum_over_rows", Ks, K64.sum(0), 2e-4)
    run("syrk", syrk)

    def fused():
        rt = torch.from_numpy(r).to(dev)
        if not engine.fused_supported(c["E"], c["eps"]):
            print("  (fused path not applicable for eps=%d)" % c["eps"])
            return
        got = engine.norm_kernel_matrices(rt, c["eps"]).cpu().numpy()
        # compare against a self-column-free reference: zero that column on both sides
        zz = z.copy()
        if raw2 is None:
            for i in range(c["nb"]):
                zz[i, :, c["start"] + i] = 0
            gotm = engine.norm_kernel_matrices(rt, c["eps"], self_col0=c["start"]).cpu().numpy()
            report("norm_kernel_matrices mask_self vs f64", gotm, orc.kernel_matrices(zz, f64=True), 3e-4)
        else:
            report("norm_kernel_matrices vs f64", got, orc.kernel_matrices(z, f64=True), 3e-4)
    run("fused norm+syrk", fused)


def check_pipeline(raw, raw2, c):
    ep, T_e = engine.stack_epochs(raw, dev)
    ep2 = engine.stack_epochs(raw2, dev)[0] if raw2 is not None else ep
    r, z, K = orc.voxel_block(raw, raw2, c["start"], c["nb"], c["eps"], shrink=False)
    zz = z.copy()
    if raw2 is None:
        for i in range(c["nb"]):
            zz[i, :, c["start"] + i] = 0
    Kref = orc.kernel_matrices(zz, f64=True)
    for prec, tol in (("tf32x3", 3e-4), ("fp16x3", 3e-4), ("bf16x3", 3e-4), ("bf16", 3e-2)):
        for flags in (0, _lib.FLAG_FISHER_IN_PASS2):
            def pipe(prec=prec, flags=flags, tol=tol):
                rows = engine.pack_epochs(ep, T_e, prec)
                cols = engine.pack_epochs(ep2, T_e, prec) if raw2 is not None else rows
                fl = flags | (_lib.FLAG_MASK_SELF if raw2 is None and engine.fused_supported(c["E"], c["eps"]) else 0)
                got = engine.voxel_kernels(rows, cols, c["start"], c["nb"], c["eps"], flags=fl)
                ref = Kref if (fl & _lib.FLAG_MASK_SELF) or raw2 is not None else None
                if ref is None:
                    print("  (unfused path keeps the self column: loose check)")
                    ref, t2 = orc.kernel_matrices(z, f64=True), 0.5
                else:
                    t2 = tol
                report("voxel_kernels %-7s flags=%d" % (prec, fl), got.cpu().numpy(), ref, t2)
            run("pipeline %s %d" % (prec, flags), pipe)

    def host():
        Kh = engine.host_voxel_kernels(raw, raw2, c["start"], c["nb"], c["eps"], "tf32x3",
                                       flags=_lib.FLAG_MASK_SELF if raw2 is None and engine.fused_supported(c["E"], c["eps"]) else 0)
        if raw2 is None and not engine.fused_supported(c["E"], c["eps"]):
            return
        report("host_voxel_kernels tf32x3", Kh, Kref, 3e-4)
    run("host pipeline", host)

    def clf():
        rows = engine.pack_epochs(ep, T_e, "tf32x3")
        cols = engine.pack_epochs(ep2, T_e, "tf32x3") if raw2 is not None else rows
        got = engine.classifier_kernel(rows, cols, 0, c["V"], c["eps"]).cpu().numpy()
        Kc, _ = orc.classifier_kernel(raw, raw2 if raw2 is not None else raw, c["eps"], 64, shrink=False)
        report("classifier_kernel tf32x3 (incl. self cols)", got, Kc, 0.2 if raw2 is None else 1e-4)
    run("classifier", clf)


def check_prologue():
    section("normalise prologue (a14)")
    V, T, E = 200, 37, 5
    rawu = [synthetic.raw_epoch(e, T, V) for e in range(E)]
    rawu[2][:, 11] = 4.0       # constant voxel -> 0
    ref = [orc.epoch_normalize(m) for m in rawu]

    def inplace():
        ep, T_e = engine.stack_epochs(rawu, dev)
        engine.epoch_normalize_(ep)
        report("epoch_normalize_ in place", ep.cpu().numpy(), np.stack(ref), 2e-6, 1.0)
    run("prologue in place", inplace)

    def fusedp():
        ep, T_e = engine.stack_epochs(rawu, dev)
        rows = engine.pack_epochs(ep, T_e, "tf32x3", normalize=True)
        out = engine.corr_block(rows, rows, 0, V).cpu().numpy()
        refc = orc.corr_block(ref, None, 0, V, f64=True)
        report("pack(normalize=1) -> corr", out, refc, 2e-6, 1.0)
    run("prologue fused", fusedp)


def timings(big):
    section("timings")
    V, T, E, eps = (50000, 200, 32, 8) if big else (8192, 200, 32, 8)
    nb = 2048 if big else 1024
    g = torch.Generator(device=dev).manual_seed(0)
    ep = torch.randn((E, T, V), device=dev, generator=g)
    engine.epoch_normalize_(ep)

    def timeit(fn, n=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / n

    work = engine.Workspace(E, V, nb, dev)
    ld = ((V + 31) // 32) * 32
    cbuf = work.buf.view(torch.float32)[: nb * E * ld].view(nb, E, ld)
    Kout = torch.empty((nb, E, E), device=dev)
    for prec in ("bf16", "bf16x3", "fp16x3", "tf32x3"):
        def one(prec=prec):
            rows = engine.pack_epochs(ep, None, prec)
            ms_g = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld))
            ms_gf = timeit(lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld, fisher_epochs=E))
            corr = nb * V * E
            flop = 2.0 * T * corr
            print("  corr gemm %-7s %8.3f ms  %7.1f TFLOP/s(alg)  %6.0f GB/s written | +fisher %8.3f ms" %
                  (prec, ms_g, flop / ms_g / 1e9, corr * 4 / ms_g / 1e6, ms_gf), flush=True)
            engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld)
            ms_s = timeit(lambda: engine.norm_kernel_matrices(cbuf[:, :, :V], eps, out=Kout))
            print("  norm+syrk (fisher in pass 2)      %8.3f ms  %6.0f GB/s read" % (ms_s, corr * 4 / ms_s / 1e6))
            ms_p = timeit(lambda: engine.voxel_kernels(rows, rows, 0, nb, eps, work=work, out=Kout))
            print("  voxel_kernels pipeline            %8.3f ms  -> %.3e corr/s" % (ms_p, corr / ms_p * 1e3), flush=True)
            ms_p2 = timeit(lambda: engine.voxel_kernels(rows, rows, 0, nb, eps, flags=_lib.FLAG_FISHER_IN_PASS2, work=work, out=Kout))
            print("  voxel_kernels (fisher in pass 2)   %8.3f ms  -> %.3e corr/s" % (ms_p2, corr / ms_p2 * 1e3), flush=True)
        run("timing " + prec, one)


def main():
    if "--timing-only" in sys.argv:
        timings("--big" in sys.argv)
        return 0
    print("lib version", _lib.load().fcma_version(), "devices", _lib.device_count(), torch.cuda.get_device_name(0))
    cases = [("self V=300 eps=4", small_case()),
             ("two masks V=260 V2=333 eps=8 E=16", small_case(V=260, V2=333, T=40, E=16, eps=8, start=5, nb=131, seed=2)),
             ("self E=10 eps=4 trailing, odd V", small_case(V=157, T=24, E=10, eps=4, start=40, nb=37, seed=3)),
             ("self E=12 eps=3 (generic eps)", small_case(V=128, T=24, E=12, eps=3, start=0, nb=128, seed=4)),
             ("two masks E=48 eps=16 (R=8)", small_case(V=96, V2=200, T=30, E=48, eps=16, start=0, nb=96, seed=5))]
    for name, (raw, raw2, c) in cases:
        section("corr: " + name)
        check_corr(raw, raw2, c)
        section("norm / kernels: " + name)
        check_norm_and_kernels(raw, raw2, c)
        section("pipeline: " + name)
        check_pipeline(raw, raw2, c)
    check_prologue()
    timings("--big" in sys.argv)
    print("\nFAILS:", FAILS)
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
