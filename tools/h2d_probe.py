import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from brainiak_b200.fcma import engine
E, T, V = 32, 200, 50000
raw = [np.random.rand(T, V).astype(np.float32) for _ in range(E)]
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ep, _ = engine.stack_epochs(raw, dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d = torch.empty((E, T, V), dtype=torch.float32, device=dev)
    for e, m in enumerate(raw):
        d[e].copy_(torch.from_numpy(m))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("pinned staging %.3f s   direct pageable copies %.3f s" % (t1 - t0, t2 - t1), flush=True)
    del ep, d
