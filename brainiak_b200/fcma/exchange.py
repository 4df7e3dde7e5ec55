"""Epoch exchange between the GPUs of one box without SMs: every rank uploads ITS share of the epochs over its own
PCIe link and the shares are all-gathered over NVLink by the copy engines (CUDA IPC mappings of the peers' buffers +
``cudaMemcpyAsync``), so the transfer can run under the persistent correlation GEMM of the previous dataset without
taking SMs away from it.

This replaces the reference's distribution step -- rank 0 reads everything and ``comm.bcast``s it epoch by epoch
(``prepare_fcma_data``, reference preprocessing.py:211-223) -- for the case where the data already sits in (shared)
host memory or on a parallel file system every rank can read: one PCIe link moves 1/W of the bytes instead of all.
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _lib

__all__ = ["EpochExchange", "epoch_partition", "epoch_groups"]


def epoch_partition(E, world):
    """Contiguous shares of the E epochs: rank r uploads epochs [e0, e0 + n)."""
    per, extra = divmod(E, world)
    out, e0 = [], 0
    for r in range(world):
        n = per + (1 if r < extra else 0)
        out.append((e0, n))
        e0 += n
    return out


def epoch_groups(E, ngroups=4):
    """Contiguous epoch groups [(e0, count)] in which an interleaved exchange completes (``gather_groups``)."""
    ngroups = max(1, min(int(ngroups), E))
    out, e = [], 0
    for g in range(ngroups):
        n = (E - e) // (ngroups - g)
        out.append((e, n))
        e += n
    return out


class EpochExchange:
    """``nbuf`` replicated epoch buffers ``[E, T, V]`` per rank, mapped into every peer.

    ``gather(k, host_share, stream)`` on every rank: H2D of this rank's epochs into its own buffer ``k``, peer copies
    of that share into buffer ``k`` of every other rank (copy engines, NVLink), then a one-element NCCL all-reduce as
    the stream-ordered "all shares have landed everywhere" barrier.  Afterwards ``buffers[k]`` holds all E epochs on
    every rank.  ``use_ipc=False`` (or a failed IPC mapping) falls back to ``all_gather_into_tensor`` in place."""

    def __init__(self, E, T, V, device, group=None, nbuf=2, use_ipc=True):
        self.lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.E, self.T, self.V = E, T, V
        self.device = torch.device(device)
        self.shares = epoch_partition(E, self.world)
        self.buffers = [torch.empty((E, T, V), dtype=torch.float32, device=self.device) for _ in range(nbuf)]
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._peer = None          # [buffer][rank] -> device address of that rank's buffer in this process
        self._opened = []
        self._aux, self._landed = None, None
        self.mode = "single"
        if self.world > 1:
            self.mode = "nccl-allgather"
            if use_ipc:
                try:
                    self._map_peers()
                    self.mode = "ipc-copy-engine"
                except Exception as exc:     # pragma: no cover - depends on the box
                    self._peer = None
                    self.ipc_error = repr(exc)
                ok = torch.tensor([1.0 if self._peer is not None else 0.0], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks take the same path
                if float(ok[0]) == 0.0:
                    self._peer, self.mode = None, "nccl-allgather"

    def _map_peers(self):
        mine = []
        for b in self.buffers:
            h = (ctypes.c_ubyte * 64)()
            off = ctypes.c_size_t(0)
            _lib.check(self.lib.fcma_ipc_get_handle(ctypes.c_void_p(b.data_ptr()), h, ctypes.byref(off)))
            mine.append((bytes(h), int(off.value)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._peer = []
        cache = {}
        with torch.cuda.device(self.device):
            for k in range(len(self.buffers)):
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(self.buffers[k].data_ptr())
                        continue
                    hb, off = everyone[r][k]
                    if (r, hb) not in cache:
                        base = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
                        _lib.check(self.lib.fcma_ipc_open_handle(buf, ctypes.byref(base)))
                        cache[(r, hb)] = base.value
                        self._opened.append(base.value)
                    row.append(cache[(r, hb)] + off)
                self._peer.append(row)

    def close(self):
        for base in self._opened:
            self.lib.fcma_ipc_close_handle(ctypes.c_void_p(base))
        self._opened = []

    def share_of(self, rank=None):
        return self.shares[self.rank if rank is None else rank]

    def interleaved_share(self, rank=None):
        """Epochs ``rank, rank + W, rank + 2W, ...``: the share of ``gather_groups`` (every W consecutive epochs come from
        W different ranks, so contiguous epoch groups complete one after the other)."""
        r = self.rank if rank is None else rank
        return list(range(r, self.E, self.world))

    def gather_groups(self, k, host_share, stream=None, ngroups=4):
        """Like ``gather`` with INTERLEAVED shares (``host_share[j]`` = epoch ``interleaved_share()[j]``): returns
        ``(buffers[k], groups, events)`` where ``events[g]`` (on ``stream``) fires when the contiguous epoch group
        ``groups[g]`` is complete on THIS rank, i.e. has been uploaded / pushed by every rank.  A consumer
        (``engine.voxel_kernels_sym_grouped``) can start on group 0 while the later ones are still in flight."""
        stream = stream or torch.cuda.current_stream(self.device)
        buf = self.buffers[k]
        groups = epoch_groups(self.E, ngroups)
        mine = self.interleaved_share()
        events = []
        ebytes = self.T * self.V * 4
        with torch.cuda.stream(stream):
            use_ipc = self.world > 1 and self._peer is not None
            if use_ipc and self._aux is None:
                self._aux = torch.cuda.Stream(device=self.device)
                self._landed = [torch.cuda.Event() for _ in range(max(1, -(-self.E // self.world)))]
            if use_ipc:
                self._aux.wait_stream(stream)
            j = 0
            for (g0, gn) in groups:
                while j < len(mine) and mine[j] < g0 + gn:
                    e = mine[j]
                    buf[e].copy_(host_share[j], non_blocking=True)
                    if use_ipc:
                        self._landed[j].record(stream)
                        self._aux.wait_event(self._landed[j])
                        sp = ctypes.c_void_p(self._aux.cuda_stream)
                        off = e * ebytes
                        with torch.cuda.device(self.device):
                            for d in range(1, self.world):
                                r = (self.rank + d) % self.world
                                _lib.check(self.lib.fcma_peer_copy_async(ctypes.c_void_p(self._peer[k][r] + off),
                                                                         ctypes.c_void_p(buf.data_ptr() + off), ebytes, sp))
                    j += 1
                if self.world > 1:
                    if use_ipc:
                        stream.wait_stream(self._aux)
                        dist.all_reduce(self._flag, group=self.group)      # every rank's epochs of this group have landed
                    else:
                        # NCCL fallback: broadcast every epoch of the group from its owner
                        for e in range(g0, g0 + gn):
                            dist.broadcast(buf[e], src=dist.get_global_rank(self.group, e % self.world)
                                           if self.group is not None else e % self.world, group=self.group)
                ev = torch.cuda.Event()
                ev.record(stream)
                events.append(ev)
        return buf, groups, events

    def gather(self, k, host_share, stream=None):
        """host_share: this rank's epochs, a (pinned) host float32 tensor ``[n, T, V]``.  Everything is enqueued on
        ``stream`` (default: the current stream); returns ``buffers[k]``.  The share goes up epoch by epoch and every
        epoch is forwarded to the peers (on an auxiliary stream) as soon as it has landed, so the PCIe upload and the
        NVLink copies overlap."""
        stream = stream or torch.cuda.current_stream(self.device)
        e0, n = self.shares[self.rank]
        buf = self.buffers[k]
        with torch.cuda.stream(stream):
            if self.world == 1 or self._peer is None:
                if n:
                    buf[e0:e0 + n].copy_(host_share, non_blocking=True)
                if self.world == 1:
                    return buf
                if len(set(m for _, m in self.shares)) == 1:
                    dist.all_gather_into_tensor(buf, buf[e0:e0 + n], group=self.group)
                else:
                    parts = [buf[a:a + m] for a, m in self.shares]
                    dist.all_gather(parts, buf[e0:e0 + n].clone(), group=self.group)
                return buf
            if self._aux is None:
                self._aux = torch.cuda.Stream(device=self.device)
                self._landed = [torch.cuda.Event() for _ in range(max(m for _, m in self.shares))]
            aux = self._aux
            aux.wait_stream(stream)              # the destination buffers are free as of this point of `stream`
            ebytes = self.T * self.V * 4
            sp = ctypes.c_void_p(aux.cuda_stream)
            with torch.cuda.device(self.device):
                for j in range(n):
                    buf[e0 + j].copy_(host_share[j], non_blocking=True)
                    self._landed[j].record(stream)
                    aux.wait_event(self._landed[j])
                    off = (e0 + j) * ebytes
                    for d in range(1, self.world):          # staggered so that no peer is everybody's first target
                        r = (self.rank + d) % self.world
                        _lib.check(self.lib.fcma_peer_copy_async(ctypes.c_void_p(self._peer[k][r] + off),
                                                                 ctypes.c_void_p(buf.data_ptr() + off), ebytes, sp))
            stream.wait_stream(aux)
            dist.all_reduce(self._flag, group=self.group)    # stream-ordered barrier: every rank's copies are done
        return buf
