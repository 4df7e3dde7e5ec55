"""Epoch separation + the per-epoch normalisation (the "normalise prologue") on the GPU.

Mirrors the arithmetic of ``brainiak.fcma.preprocessing._separate_epochs`` (reference
preprocessing.py:41-92).  NIfTI loading / masking (preprocessing.py:156-232 via brainiak.image) is
out of scope: the functions here take the already masked ``[nVoxels, nTRs]`` activity arrays.
"""
import numpy as np

from .. import _lib
from . import engine

__all__ = ["separate_epochs", "separate_epochs_device", "broadcast_epochs"]


def _epoch_slices(activity_data, epoch_list):
    """Enumerate (subject, condition, TR-mask) in the reference's order (preprocessing.py:71-77)."""
    out = []
    for sid in range(len(epoch_list)):
        epoch = epoch_list[sid]
        for cond in range(epoch.shape[0]):
            sub_epoch = epoch[cond, :, :]
            for eid in range(epoch.shape[1]):
                r = np.sum(sub_epoch[eid, :])
                if r > 0:
                    out.append((sid, cond, sub_epoch[eid, :] == 1))
    return out


def separate_epochs_device(activity_data, epoch_list, device=None):
    """Returns ``(epochs [E, Tmax, V] float32 CUDA (normalised, zero rows beyond T_e), T_e, labels)``."""
    import torch
    _lib.load()
    _lib.require_device()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    slices = _epoch_slices(activity_data, epoch_list)
    raw, labels = [], []
    for sid, cond, m in slices:
        mat = np.ascontiguousarray(np.asarray(activity_data[sid], dtype=np.float32)[:, m].T)
        raw.append(mat)
        labels.append(cond)
    ep, T_e = engine.stack_epochs(raw, dev)
    engine.epoch_normalize_(ep, T_e)
    return ep, T_e, labels


def separate_epochs(activity_data, epoch_list, device=None):
    """Drop-in for ``_separate_epochs``: ``(raw_data: list of float32 [T_e, V], labels)``."""
    ep, T_e, labels = separate_epochs_device(activity_data, epoch_list, device)
    host = ep.cpu().numpy()
    return [np.ascontiguousarray(host[e, :T_e[e], :]) for e in range(len(T_e))], labels


_separate_epochs = separate_epochs


def broadcast_epochs(epochs, src=0):
    """NCCL broadcast of the normalised epoch tensor — the GPU replacement of the per-epoch
    ``comm.bcast`` loop of ``prepare_fcma_data`` (reference preprocessing.py:211-223).
    ``epochs`` must be allocated with the right shape on every rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(epochs, src=src)
    return epochs
