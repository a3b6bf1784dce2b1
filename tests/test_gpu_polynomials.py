"""Row f-4: elimination and apply over OTHER field polynomials than the reference's (tablegen.c:7-12).
The truth is an independent numpy model of the field (carry-less multiplication mod the polynomial) driving
the caller-side elimination loop of SURVEY 3.7 -- pure Python, small systems; all three elimination paths
(table kernel, outer-panel + tcgen05 update) and both apply kernels must reproduce it byte for byte."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_helpers import dev, host
from helpers import GF4, GF16, GF256, FIELD_BITS, splitmix_bytes
from oblas_b200 import capi, device

pytestmark = pytest.mark.gpu


def mul_table(bits, poly):
    n = 1 << bits
    T = np.zeros((n, n), dtype=np.uint8)
    for a in range(n):
        for b in range(n):
            r, x = 0, a
            for i in range(bits):
                if (b >> i) & 1:
                    r ^= x
                x <<= 1
                if x & n:
                    x ^= poly
            T[a, b] = r
    return T


def unpack(M, bits, cols):
    per = 8 // bits
    out = np.zeros((M.shape[0], cols), dtype=np.uint8)
    for j in range(cols):
        out[:, j] = (M[:, j // per] >> ((j % per) * bits)) & ((1 << bits) - 1)
    return out


def pack(E, bits, nbytes):
    per = 8 // bits
    out = np.zeros((E.shape[0], nbytes), dtype=np.uint8)
    for j in range(E.shape[1]):
        out[:, j // per] |= (E[:, j] << ((j % per) * bits)).astype(np.uint8)
    return out


def model_solve(T, A, B):
    """the loop of oracle/ref_driver.c on element matrices A (rows x cols) and B (rows x n), any field"""
    A, B = A.copy(), B.copy()
    inv = np.zeros(T.shape[0], dtype=np.uint8)
    for a in range(1, T.shape[0]):
        inv[a] = int(np.nonzero(T[a] == 1)[0][0])
    rows, cols = A.shape
    t, piv = 0, []
    for c in range(cols):
        if t >= rows:
            break
        nz = np.nonzero(A[t:, c])[0]
        if len(nz) == 0:
            continue
        p = t + int(nz[0])
        A[[t, p]] = A[[p, t]]
        B[[t, p]] = B[[p, t]]
        s = inv[A[t, c]]
        A[t] = T[s][A[t]]
        B[t] = T[s][B[t]]
        for r in range(rows):
            f = A[r, c]
            if r != t and f:
                A[r] ^= T[f][A[t]]
                B[r] ^= T[f][B[t]]
        piv.append(c)
        t += 1
    return t, piv, A, B


@pytest.fixture()
def lib(cuda_lib):
    device.bind(0)
    yield cuda_lib
    for f in (GF4, GF16, GF256):
        cuda_lib.oblas_b200_set_field_tables(f, 0)
    cuda_lib.oblas_b200_set_tuning(-1, -1)


def register(lib, bits, poly):
    n = 1 << bits
    lo, hi, inv = (C.c_uint8 * (16 * n))(), (C.c_uint8 * (16 * n))(), (C.c_uint8 * n)()
    assert lib.oblas_b200_make_tables(bits, poly, lo, hi, inv) == 0
    h = lib.oblas_b200_register_tables(bits, lo, hi, inv)
    assert h >= 16
    return h


@pytest.mark.parametrize("field,poly,geo", [
    (GF256, 0x11B, 1), (GF256, 0x11B, 3), (GF256, 0x12B, 4), (GF256, 0x187, 3),
    (GF16, 0x19, -1), (GF16, 0x1F, 0), (GF4, 0x7, -1)])
def test_elimination_over_another_polynomial(lib, field, poly, geo):
    bits = FIELD_BITS[field]
    K, Lel, nsys = 96, 64 * (8 // bits), 3          # B: 64 bytes of elements
    a_bytes = (K * bits // 8 + 15) // 16 * 16
    T = mul_table(bits, poly)
    h = register(lib, bits, poly)
    capi.check(lib.oblas_b200_set_field_tables(field, h), "set_field_tables")
    lib.oblas_b200_set_tuning(geo, -1)
    A = np.zeros((nsys, K, a_bytes), dtype=np.uint8)
    used = K * bits // 8
    A[:, :, :used] = splitmix_bytes(500 + poly, nsys * K * used).reshape(nsys, K, used)
    A[1, 40] = A[1, 9]                                 # a rank-deficient system
    B = splitmix_bytes(600 + poly, nsys * K * 64).reshape(nsys, K, 64).copy()
    dA, dB = dev(A), dev(B)
    rank, piv = device.solve_batch(field, dA, K, dB, want_pivcols=True)
    device.sync()
    gA, gB, rank, piv = host(dA), host(dB), host(rank), host(piv)
    for s in range(nsys):
        rk, pv, wa, wb = model_solve(T, unpack(A[s], bits, K), unpack(B[s], bits, Lel))
        assert rank[s] == rk and piv[s][:rk].tolist() == pv, s
        assert np.array_equal(gA[s], pack(wa, bits, a_bytes)), s
        assert np.array_equal(gB[s], pack(wb, bits, 64)), s
    assert rank[1] == K - 1
    # back on the reference's polynomial the same input gives other bytes (the switch really switches)
    capi.check(lib.oblas_b200_set_field_tables(field, 0))
    if poly not in (285, 19, 7):
        dA2, dB2 = dev(A), dev(B)
        device.solve_batch(field, dA2, K, dB2)
        device.sync()
        assert not np.array_equal(host(dB2), gB)


@pytest.mark.parametrize("geo", [1, 3], ids=["table_kernel", "tcgen05"])
def test_apply_over_another_polynomial(lib, geo):
    poly = 0x11B
    T = mul_table(8, poly)
    M, Kc, N = 40, 24, 496
    Cm = splitmix_bytes(701, M * Kc).reshape(M, Kc).copy()
    D = splitmix_bytes(702, Kc * N).reshape(Kc, N).copy()
    want = np.zeros((M, N), dtype=np.uint8)
    for i in range(M):
        for j in range(Kc):
            want[i] ^= T[Cm[i, j]][D[j]]
    capi.check(lib.oblas_b200_set_field_tables(GF256, register(lib, 8, poly)))
    lib.oblas_b200_set_tuning(-1, geo)
    Y = device.apply(dev(Cm), dev(D))
    device.sync()
    assert np.array_equal(host(Y), want)


def test_tables_that_are_not_a_field_are_refused(lib):
    """the shipped GF(4) pair (HI table defect) registers for the row kernels but cannot drive an elimination"""
    lo, hi, inv = capi.table("GF2_2_SHUF_LO"), capi.table("GF2_2_SHUF_HI"), capi.table("GF2_2_INV")
    h = lib.oblas_b200_register_tables(2, lo, hi, inv)
    assert h >= 16
    assert lib.oblas_b200_set_field_tables(GF4, h) != 0 and b"not those of a field" in lib.oblas_b200_last_error()
    h8 = register(lib, 8, 0x11B)
    assert lib.oblas_b200_set_field_tables(GF16, h8) != 0      # wrong element width
