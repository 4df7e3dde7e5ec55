"""``compute_correlation`` on the GPU — drop-in for ``brainiak.fcma.util`` (reference util.py:32-134)."""
import numpy as np

from .. import _lib
from . import engine

__all__ = ["compute_correlation"]


def compute_correlation(matrix1, matrix2, return_nans=False):
    """Pearson correlation between the rows of ``matrix1`` [r1, c] and ``matrix2`` [r2, c].

    Same contract as reference util.py:63-134: inputs are cast to float32, rows are z-scored
    (ddof=0) and divided by sqrt(c) (util.py:32-60; NaN -> 0 unless ``return_nans``), the result is
    a C-contiguous float32 ``[r1, r2]`` array.  Raises ``ValueError('Dimension discrepancy')``.
    """
    import torch
    # self-correlation is decided on the caller's objects, BEFORE the float32 copies below (which never alias)
    a1, a2 = np.asarray(matrix1), np.asarray(matrix2)
    same = matrix2 is matrix1 or (a1.shape == a2.shape and a1.dtype == a2.dtype and a1.strides == a2.strides
                                  and a1.size > 0 and a1.__array_interface__['data'][0] == a2.__array_interface__['data'][0])
    matrix1 = a1.astype(np.float32)
    matrix2 = matrix1 if same else a2.astype(np.float32)
    [r1, d1] = matrix1.shape
    [r2, d2] = matrix2.shape
    if d1 != d2:
        raise ValueError('Dimension discrepancy')
    _lib.load()
    _lib.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    m1 = torch.from_numpy(np.ascontiguousarray(matrix1)).to(dev)
    engine.row_normalize_(m1, nan_to_zero=not return_nans)
    if same:
        m2 = m1
    else:
        m2 = torch.from_numpy(np.ascontiguousarray(matrix2)).to(dev)
        engine.row_normalize_(m2, nan_to_zero=not return_nans)
    # The normalised rows are one "epoch" of d time points x r voxels: the same tensor-core contraction
    # as the FCMA correlation block (E = 1), fp32-faithful 3-product split; unit-norm rows keep |x| <= 1.
    op1 = engine.pack_epochs(m1.t().contiguous().unsqueeze(0), None, "fp32")
    op2 = op1 if same else engine.pack_epochs(m2.t().contiguous().unsqueeze(0), None, op1.precision)
    corr = engine.corr_block(op1, op2, 0, r1, layout=0)          # [r1, 1, r2]
    return np.ascontiguousarray(corr[:, 0, :].cpu().numpy())
