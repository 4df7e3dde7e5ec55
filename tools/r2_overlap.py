#!/usr/bin/env python
"""Round-2 A/B for VERDICT item 3: does overlapping the write-bound correlation GEMM with the read-bound
normalise+SYRK passes shorten the step?  Two halves of the job (equal-area shards, own scratch, own K arrays) run
  (a) one after the other on ONE stream (what the product does, per pass: GEMM -> row pass -> column pass), and
  (b) concurrently on TWO streams, so that the GEMM of one shard can overlap the SYRK passes of the other,
with the GEMM grid either taking every SM (hardware fills the tails) or capped at n CTA pairs (FCMA_GEMM_PAIRS, diagnostic
build) so that the other stream's kernels always find free SMs.  Prints step times, clocks and power."""
import os, sys, subprocess, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib, build as _build
_build.build(diag=True)
_lib.use_diag_build()
from brainiak_b200.fcma import engine
V, T, E, eps = 50000, 200, 32, 8
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
op = engine.pack_epochs(ep, None, "fp16x3")
parts = engine.sym_row_partition(V, 2, pack_frac=0.0)
works = [engine.SymWorkspace(E, V, 4096, dev, start=s, transposed_copy=False) for s, _ in parts]
Ks = [torch.zeros((V, E, E), device=dev) for _ in parts]
streams = [torch.cuda.Stream(device=dev) for _ in parts]


def sequential():
    for (s, n), w, K in zip(parts, works, Ks):
        K.zero_()
        engine.voxel_kernels_sym(op, s, n, eps, work=w, out=K)


def concurrent():
    cur = torch.cuda.current_stream()
    for (s, n), w, K, st in zip(parts, works, Ks, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            K.zero_()
            engine.voxel_kernels_sym(op, s, n, eps, work=w, out=K)
    for st in streams:
        cur.wait_stream(st)


def clocks():
    out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                         capture_output=True, text=True).stdout.strip()
    return out


def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    mid = None
    b.record()
    # sample clocks / power while the queue drains
    time.sleep(0.15)
    mid = clocks()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, mid


ref = None
for rep in range(2):
    for name, fn, pairs in (("sequential, one stream", sequential, None), ("two streams, GEMM on all SMs", concurrent, None),
                            ("two streams, GEMM capped at 64 pairs", concurrent, "64"),
                            ("two streams, GEMM capped at 56 pairs", concurrent, "56"),
                            ("two streams, GEMM capped at 37 pairs", concurrent, "37"),
                            ("sequential, GEMM capped at 64 pairs", sequential, "64")):
        os.environ.pop("FCMA_GEMM_PAIRS", None)
        if pairs:
            os.environ["FCMA_GEMM_PAIRS"] = pairs
        ms, ck = timeit(fn)
        tot = Ks[0] + Ks[1]
        if ref is None:
            ref = tot.clone()
        err = float((tot - ref).abs().max() / ref.abs().max())
        print("%-40s %.1f ms per step   [sm MHz, W: %s]   max|dK|/max|K| vs first run %.1e" % (name, ms, ck, err), flush=True)
os.environ.pop("FCMA_GEMM_PAIRS", None)
