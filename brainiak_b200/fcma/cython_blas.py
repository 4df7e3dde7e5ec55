"""Host-buffer BLAS-shaped entry points with the signatures of the reference's Cython module
``brainiak.fcma.cython_blas`` (reference cython_blas.pyx:20, 118, 388, 480), computed on the GPU.

They exist so that code written against the reference's native-module boundary keeps working;
the classes in this package use the fused device pipelines instead.  As in the reference all
results are written in place into caller-allocated float32 C-contiguous numpy buffers, and a
wrong dtype / layout raises ``ValueError`` (Cython typed-memoryview behaviour)."""
import numpy as np

from .. import _lib
from . import engine

__all__ = ["compute_self_corr_for_voxel_sel", "compute_corr_vectors", "compute_kernel_matrix",
           "compute_single_matrix_multiplication"]


def _f32c(a, ndim, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.ndim != ndim \
            or not a.flags.c_contiguous:
        raise ValueError("%s: Buffer dtype mismatch or not C-contiguous float32 with %d dims"
                         % (name, ndim))
    return a


def _dev():
    import torch
    _lib.load()
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _epoch_gemm(py_a, py_b, start_voxel, n):
    """[n, V2] = B[:, s:s+n].T @ A for one epoch (fp32 FFMA path, reference-order numerics)."""
    import torch
    dev = _dev()
    a = torch.from_numpy(py_a).to(dev).unsqueeze(0)     # cols operand [1, T, V2]
    b = a if py_b is py_a else torch.from_numpy(py_b).to(dev).unsqueeze(0)   # rows operand [1, T, V]
    out = engine.corr_block_f32(b, a, start_voxel, n, layout=0)      # [n, 1, V2]
    return out[:, 0, :]


def compute_self_corr_for_voxel_sel(py_trans_a, py_trans_b, py_m, py_n, py_k, py_alpha, py_a,
                                    py_lda, py_start_voxel, py_b, py_ldb, py_beta, py_c, py_ldc,
                                    py_start_epoch):
    """cython_blas.pyx:20-116: ``C[:, start_epoch, :] = B[:, s:s+n].T @ A`` (C is [n, E, V2])."""
    a, b, c = _f32c(py_a, 2, "py_a"), _f32c(py_b, 2, "py_b"), _f32c(py_c, 3, "py_c")
    if py_trans_a != 'N' or py_trans_b != 'T' or py_alpha != 1.0 or py_beta != 0.0:
        raise ValueError("only the reference's call pattern ('N','T',alpha=1,beta=0) is supported")
    c[:py_n, py_start_epoch, :py_m] = _epoch_gemm(a, b, int(py_start_voxel), int(py_n)).cpu().numpy()


def compute_corr_vectors(py_trans_a, py_trans_b, py_m, py_n, py_k, py_alpha, py_a, py_lda, py_b,
                         py_ldb, py_beta, py_c, py_ldc, py_start_voxel, py_start_sample):
    """cython_blas.pyx:388-478: ``C[start_sample, :, :] = B[:, s:s+n].T @ A`` (C is [E, n, V2])."""
    a, b, c = _f32c(py_a, 2, "py_a"), _f32c(py_b, 2, "py_b"), _f32c(py_c, 3, "py_c")
    if py_trans_a != 'N' or py_trans_b != 'T' or py_alpha != 1.0 or py_beta != 0.0:
        raise ValueError("only the reference's call pattern ('N','T',alpha=1,beta=0) is supported")
    c[py_start_sample, :py_n, :py_m] = _epoch_gemm(a, b, int(py_start_voxel), int(py_n)).cpu().numpy()


def compute_kernel_matrix(py_uplo, py_trans, py_n, py_k, py_alpha, py_a, py_start_voxel, py_lda,
                          py_beta, py_c, py_ldc):
    """cython_blas.pyx:118-207: ``C = beta*C + A[start_voxel] A[start_voxel]^T`` (full, mirrored)."""
    import torch
    a, c = _f32c(py_a, 3, "py_a"), _f32c(py_c, 2, "py_c")
    if py_alpha != 1.0:
        raise ValueError("alpha must be 1.0")
    dev = _dev()
    z = torch.from_numpy(a[py_start_voxel, :py_n, :py_k]).to(dev).unsqueeze(0).contiguous()
    K = torch.from_numpy(c[:py_n, :py_n].copy()).to(dev).unsqueeze(0).contiguous()
    engine.kernel_matrices(z, beta=float(py_beta), out=K)
    c[:py_n, :py_n] = K[0].cpu().numpy()


def compute_single_matrix_multiplication(py_trans_a, py_trans_b, py_m, py_n, py_k, py_alpha, py_a,
                                         py_lda, py_b, py_ldb, py_beta, py_c, py_ldc):
    """cython_blas.pyx:480-560 for the reference's only call pattern ('T','N'): row-major
    ``C[py_n, py_m] = B[py_n, py_k] @ A[py_m, py_k].T`` (classifier.py:253-264, util.py:126-133)."""
    import torch
    a, b, c = _f32c(py_a, 2, "py_a"), _f32c(py_b, 2, "py_b"), _f32c(py_c, 2, "py_c")
    if py_trans_a != 'T' or py_trans_b != 'N' or py_alpha != 1.0 or py_beta != 0.0:
        raise ValueError("only the reference's call pattern ('T','N',alpha=1,beta=0) is supported")
    dev = _dev()
    A = torch.from_numpy(a).to(dev)
    B = torch.from_numpy(b).to(dev)
    c[:py_n, :py_m] = engine.gemm_nt(B[:py_n, :py_k], A[:py_m, :py_k]).cpu().numpy()
