#!/usr/bin/env python
"""SM clock / board power while ONE kernel of the pipeline runs back to back (NVML polled at ~500 Hz):
tells whether a kernel's time is tensor-pipe time at a power-capped clock or pipeline bubbles."""
import os, sys, time, threading, statistics, torch, pynvml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200 import _lib, build as _build
_build.build(diag=True)      # the FCMA_* knobs exist only in the diagnostic build (-DFCMA_DIAG)
_lib.use_diag_build()
from brainiak_b200.fcma import engine
V, T, E, nb = 50000, 200, 32, 4096
dev = torch.device("cuda:0")
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
g = torch.Generator(device=dev).manual_seed(0)
ep = torch.randn((E, T, V), device=dev, generator=g)
engine.epoch_normalize_(ep)
work = engine.Workspace(E, V, nb, dev)
ld = ((V + 31) // 32) * 32
cbuf = work.buf.view(torch.float32)[: nb * E * ld].view(nb, E, ld)
K = torch.empty((nb, E, E), device=dev)
def run(label, fn, secs=1.5):
    fn(); torch.cuda.synchronize()
    samples, stop = [], False
    def poll():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.002)
    th = threading.Thread(target=poll); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(5): fn()
        n += 5; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop = True; th.join()
    half = samples[len(samples) // 2:]
    mhz = statistics.median(s[0] for s in half); w = statistics.median(s[1] for s in half)
    print(f"{label:34s} {e0.elapsed_time(e1)/n:7.3f} ms/launch  SM {mhz:6.0f} MHz  {w:6.0f} W", flush=True)
for prec in ("fp16x3", "bf16", "tf32x3"):
    rows = engine.pack_epochs(ep, None, prec)
    for dbg in sys.argv[1:] or ("0", "4", "20"):
        os.environ["FCMA_GEMM_DEBUG"] = dbg
        run(f"gemm {prec} stream debug={dbg}", lambda: engine.corr_block(rows, rows, 0, nb, out=cbuf, ld=ld))
    os.environ["FCMA_GEMM_DEBUG"] = "0"
    del rows
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
run("torch.matmul bf16 8192^3", lambda: torch.matmul(a, b))
