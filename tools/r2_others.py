#!/usr/bin/env python
"""The `other_configs` part of bench.py alone (BASELINE configs[1], [3], [4]): python tools/r2_others.py"""
import json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from brainiak_b200 import _lib
from brainiak_b200.fcma import engine
args = types.SimpleNamespace(precision="fp16x3", no_cpu_baseline="--no-parity" in sys.argv)
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _lib.load()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
out = bench.other_configs(args, lib, engine, torch, dist, dev, 0, 1, 0, float(peaks["hbm_gbs"]), float(peaks["bf16_tflops_sustained"]))
print(json.dumps(out, indent=1))
