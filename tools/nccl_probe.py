#!/usr/bin/env python
"""Collective / copy timings behind the multi-GPU step (run under torchrun):
   reduce-scatter and reduce of the [V,E,E] kernel array, the epoch exchange (IPC copy engines vs NCCL all-gather),
   the H2D of a rank's share.  CUDA events, max over ranks."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brainiak_b200.fcma.exchange import EpochExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pg2 = dist.new_group(backend="nccl")
V, T, E = 50000, 200, 32
per = -(-V // world)


def timeit(fn, reps=5):
    fn(); fn()
    dist.barrier(device_ids=[local]); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0])


K = torch.zeros((world * per, E, E), device=dev)
Km = torch.empty((per, E, E), device=dev)
res = {}
res["reduce_scatter_%dMB" % (K.numel() * 4 >> 20)] = timeit(lambda: dist.reduce_scatter_tensor(Km, K))
res["reduce_to_rank0"] = timeit(lambda: dist.reduce(K, dst=0))
res["all_reduce"] = timeit(lambda: dist.all_reduce(K))
flag = torch.zeros(1, device=dev)
res["all_reduce_1elem"] = timeit(lambda: dist.all_reduce(flag, group=pg2), reps=20)
for ipc in (True, False):
    x = EpochExchange(E, T, V, dev, group=pg2, nbuf=1, use_ipc=ipc)
    e0, n = x.share_of()
    host = torch.randn((n, T, V)).pin_memory()
    res["exchange_%s_with_h2d" % x.mode] = timeit(lambda: x.gather(0, host))
    res["h2d_share_%dMB" % (host.numel() * 4 >> 20)] = timeit(lambda: x.buffers[0][e0:e0 + n].copy_(host, non_blocking=True))
    x.close()
    del x
if rank == 0:
    print(world, "ranks:", {k: round(v, 3) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
