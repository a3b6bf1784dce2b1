"""GPU-side test plumbing (torch tensors <-> liboblas_b200.so).  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from helpers import GF256, oracle_table
from oblas_b200 import capi, device


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


_MUL = None


def gf256_mul_table():
    """MUL[u][x] from the oracle's split-nibble tables (oblas_ref.c:3-7 byte map)"""
    global _MUL
    if _MUL is None:
        lo, hi = oracle_table(GF256, 3), oracle_table(GF256, 4)
        x = np.arange(256)
        _MUL = np.stack([hi[u * 16 + (x >> 4)] ^ lo[u * 16 + (x & 15)] for u in range(256)]).astype(np.uint8)
    return _MUL


def gf256_vecmat(coeff_row, X):
    """XOR_j coeff_row[j] * X[j, :] on the CPU (numpy), the CPU side of the
    encode -> decode round-trip checks at sizes the scalar oracle cannot reach"""
    MUL = gf256_mul_table()
    acc = np.zeros(X.shape[1], dtype=np.uint8)
    for j in np.nonzero(coeff_row)[0]:
        acc ^= MUL[coeff_row[j]][X[j]]
    return acc
