"""world_size-2 gloo test of the N>1 host logic of VoxelSelector.run: static row shards, per-rank
scoring, gather on the master, stable sort.  The GPU stage is replaced by a CPU scorer built from
the oracle (tests may use the oracle) so the plumbing can run without a device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from sklearn import svm

from brainiak_b200.fcma import synthetic
from brainiak_b200.fcma.voxelselector import VoxelSelector, shrink_kernels_


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_score_rows(self, start, n, clf):
    from oracle import fcma_oracle as orc
    _, _, K = orc.voxel_block(self.raw_data, self.raw_data2, start, n, self.epochs_per_subj, shrink=False)
    shrink_kernels_(K)
    return self._do_cross_validation(clf, K, (start, n))


def _worker(rank, world, port, V, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raw, labels = synthetic.make_epochs(V, 16, 8, informative=3, signal=1.5)
        VoxelSelector._score_rows = _cpu_score_rows
        vs = VoxelSelector(labels, 4, 2, raw, voxel_unit=5, process_num=0, master_rank=1)
        res = vs.run(svm.SVC(kernel="precomputed", shrinking=False, C=1))
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_run_world_size_2_gloo(monkeypatch):
    V, world = 21, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), V, out), nprocs=world, join=True)
    assert out[0] == []                       # non-master ranks return [] (voxelselector.py:172-173)
    res = out[1]
    assert sorted(v for v, _ in res) == list(range(V))
    accs = [a for _, a in res]
    assert accs == sorted(accs, reverse=True)
    # identical to the single-process result
    raw, labels = synthetic.make_epochs(V, 16, 8, informative=3, signal=1.5)
    monkeypatch.setattr(VoxelSelector, "_score_rows", _cpu_score_rows)
    vs = VoxelSelector(labels, 4, 2, raw, voxel_unit=5, process_num=0)
    serial = vs.run(svm.SVC(kernel="precomputed", shrinking=False, C=1))
    assert serial == res
    # ties keep voxel order (stable sort over rank-ordered shards)
    for (v0, a0), (v1, a1) in zip(res, res[1:]):
        if a0 == a1:
            assert v0 < v1


# ----------------------------------------------------------------------------------------------------------------
# Symmetric single-mask pipeline over several ranks: every rank contracts rows [s, s+n) with the
# columns [s, V) only, adds every block to its row voxels AND (transposed) to its column voxels, and the ranks' partial
# [V, E, E] arrays are summed with one all-reduce.  Here the device stage is the oracle's normalised correlation block, so
# the decomposition (engine.sym_row_partition, the pass structure, exactly-once coverage) and the collective run on CPU.
def _sym_partial(z, s, n, rows_per_pass):
    """What fcma_voxel_kernels_sym(start=s, nb=n) accumulates: z = [V, E, V] normalised correlations (symmetric in
    its first and last axis)."""
    V, E, _ = z.shape
    K = np.zeros((V, E, E), np.float64)
    for a in range(s, s + n, rows_per_pass):
        nn = min(rows_per_pass, s + n - a)
        blk = z[a:a + nn, :, a:]                                    # block A: rows of the pass x columns [a, V)
        K[a:a + nn] += np.einsum("iej,ifj->ief", blk, blk)          # row pass
        right = blk[:, :, nn:]                                      # columns right of the diagonal part
        K[a + nn:] += np.einsum("iej,ifj->jef", right, right)       # column pass: z(j, :, i) == z(i, :, j)
    return K


def _sym_worker(rank, world, port, V, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import fcma_oracle as orc
        from brainiak_b200.fcma import engine
        raw, _ = synthetic.make_epochs(V, 12, 8, seed=5)
        _, z, _ = orc.voxel_block(raw, None, 0, V, 4, shrink=False)     # [V, E, V]
        z = z.astype(np.float64)
        for i in range(V):
            z[i, :, i] = 0                                              # mask_self: the one non-symmetric column
        s, n = engine.sym_row_partition(V, world, align=4)[rank]
        part = torch.from_numpy(_sym_partial(z, s, n, rows_per_pass=8) if n > 0 else np.zeros((V, 8, 8)))
        dist.all_reduce(part)                                           # the path's one exchange step
        if rank == 0:
            out["K"] = part.numpy()
            out["full"] = np.einsum("iej,ifj->ief", z, z)
            out["sym_err"] = float(np.max(np.abs(z - np.transpose(z, (2, 1, 0)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_symmetric_decomposition_world_size_n_gloo(world):
    V = 45
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sym_worker, args=(world, _free_port(), V, out), nprocs=world, join=True)
    assert out["sym_err"] <= 2e-6                  # the oracle's normalised correlations are symmetric (fp32 rounding)
    scale = np.max(np.abs(out["full"]))
    assert np.max(np.abs(out["K"] - out["full"])) <= 1e-5 * scale
