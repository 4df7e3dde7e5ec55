"""Host-buffer face of the normalisation kernel with the signature of the reference's pybind11
module ``brainiak.fcma.fcma_extension`` (reference src/fcma_extension.cc:29-90)."""
import ctypes

import numpy as np

from .. import _lib

__all__ = ["normalization"]


def normalization(py_data, epochsPerSubj):
    """In-place Fisher-z + within-subject z-score of a float32 C-contiguous ``[n, E, V]`` array.

    Like the reference's ``py::array::forcecast`` argument (fcma_extension.cc:29-30) a non-float32 or
    non-contiguous input is converted to a temporary and the caller's array is left unchanged.
    Raises ``RuntimeError`` for non-3D input (fcma_extension.cc:47-48)."""
    lib = _lib.load()
    arr = np.asarray(py_data)
    if arr.ndim != 3:
        raise RuntimeError("The multi-subject correlation data structure must be 3D")
    work = arr if (arr.dtype == np.float32 and arr.flags.c_contiguous) else \
        np.ascontiguousarray(arr, dtype=np.float32)
    n0, E, n2 = work.shape
    if work.size == 0:
        return None
    _lib.check(lib.fcma_host_within_subject_norm(work.ctypes.data_as(ctypes.c_void_p), n0, E, n2,
                                                 int(epochsPerSubj), _current_device()))
    return None


def _current_device():
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    except Exception:  # pragma: no cover
        return 0
