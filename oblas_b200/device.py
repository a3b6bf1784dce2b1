"""Pass torch CUDA tensors and streams to liboblas_b200.so.

torch is used only for device memory, streams and events; every function below
ends in exactly one library call (a kernel launch on the current torch stream).
"""
import ctypes as C

import torch

from . import capi
from .capi import FIELD_BITS, GF2, GF4, GF16, GF256  # noqa: F401


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def bind(device=None):
    """Initialise the library on the current torch device and make it launch on the
    current torch stream (so torch.cuda.Event timing sees the kernels)."""
    if device is not None:
        torch.cuda.set_device(device)
    capi.init(torch.cuda.current_device())
    capi.lib().oblas_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))


def sync():
    """Wait for the library's stream and collect the tensor-core kernels' status words (raises if a
    launch since the last call gave up on a barrier wait -- its result would be invalid).  The checked
    way to finish after solve_batch / apply, which only enqueue."""
    capi.check(capi.lib().oblas_b200_sync(), "sync")


def fill_random(t, seed, byte_offset=0):
    assert t.is_cuda and t.is_contiguous()
    capi.check(capi.lib().oblas_b200_fill_random(_p(t), t.numel() * t.element_size(), seed, byte_offset),
               "fill_random")


def rowops(field, op, a, b, dst_rows, src_rows, u, row_bytes=None, variant=0, mul=capi.MUL_AUTO):
    """a, b: uint8 CUDA tensors [rows, stride]; dst_rows/src_rows: int32 CUDA tensors;
    u: uint8 CUDA tensor.  One kernel launch."""
    assert a.dtype == torch.uint8 and a.dim() == 2 and a.stride(1) == 1
    n = dst_rows.numel()
    rb = row_bytes if row_bytes is not None else a.shape[1]
    bb = b if b is not None else a
    capi.check(capi.lib().oblas_b200_rowops(field, variant, op, mul, _p(a), a.stride(0), _p(bb), bb.stride(0),
                                            rb, _p(dst_rows), _p(src_rows), _p(u), n), "rowops")


def solve_batch(field, A, cols, B=None, want_pivcols=False):
    """A: uint8 CUDA tensor [nsys, rows, a_bytes] (in place), B: [nsys, rows, b_bytes] or None.
    Returns (rank int32 [nsys], pivcols int32 [nsys, rows] or None)."""
    assert A.dtype == torch.uint8 and A.dim() == 3 and A.stride(2) == 1
    nsys, rows, a_bytes = A.shape
    rank = torch.empty(nsys, dtype=torch.int32, device=A.device)
    piv = torch.full((nsys, rows), -1, dtype=torch.int32, device=A.device) if want_pivcols else None
    capi.check(capi.lib().oblas_b200_solve_batch(
        field, _p(A), A.stride(1), A.stride(0), rows, cols, a_bytes,
        _p(B), B.stride(1) if B is not None else 0, B.stride(0) if B is not None else 0,
        B.shape[2] if B is not None else 0, _p(rank), _p(piv), nsys), "solve_batch")
    return rank, piv


def apply(Cm, D, Y=None, accumulate=False):
    """Y = C * D over GF(256).  C: [M, Kc(+pad)] uint8, D: [Kc, nbytes] uint8."""
    M, Kc = Cm.shape[0], D.shape[0]
    nbytes = D.shape[1]
    if Y is None:
        Y = torch.empty((M, nbytes), dtype=torch.uint8, device=D.device)
    capi.check(capi.lib().oblas_b200_gf256_apply(_p(Cm), Cm.stride(0), M, Kc, _p(D), D.stride(0), _p(Y),
                                                 Y.stride(0), nbytes, 1 if accumulate else 0), "apply")
    return Y
