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
