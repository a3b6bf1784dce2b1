"""Column operations and element accessors on device-resident packed matrices
(oblas_b200_swapcols_batch / get_elements / set_elements) against the
element semantics of gfmat_swapcol / gfmat_get / gfmat_set (gfmat.c:83-98, :215-231), restated
with numpy on the unpacked elements, and against the compiled reference where it is present."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_helpers import dev, host
from helpers import GF2, GF4, GF16, GF256, FIELD_BITS, mat_view, reference, splitmix_bytes
from oblas_b200 import capi, device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _bind(cuda_lib):
    device.bind(0)


def _packed(field, rows, cols, seed, nsys=1):
    bits = FIELD_BITS[field]
    st = ((cols * bits + 31) // 32 * 4 + 15) // 16 * 16
    M = splitmix_bytes(seed, nsys * rows * st).reshape(nsys, rows, st).copy()
    return M, st


def _unpack(M, field, cols):
    bits = FIELD_BITS[field]
    b = np.unpackbits(M, axis=-1, bitorder="little")[..., :cols * bits]
    b = b.reshape(M.shape[:-1] + (cols, bits))
    return (b * (1 << np.arange(bits, dtype=np.uint8))).sum(axis=-1).astype(np.uint8)


def _vp(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("field", [GF2, GF4, GF16, GF256])
def test_swapcols_batch(field, cuda_lib):
    rows, cols, nsys = 37, 203, 5
    M, st = _packed(field, rows, cols, 40 + field, nsys)
    ci = np.array([0, 7, 202, 33, 64], dtype=np.uint32)
    cj = np.array([1, 7, 0, 34, 95], dtype=np.uint32)     # same word, i == j, far apart, ...
    want = _unpack(M, field, cols)
    for s in range(nsys):
        want[s][:, [ci[s], cj[s]]] = want[s][:, [cj[s], ci[s]]]
    dM = dev(M)
    # index arrays in plain host memory (staged by the library) ...
    capi.check(cuda_lib.oblas_b200_swapcols_batch(field, _vp(dM), st, rows * st, rows,
                                                  ci.ctypes.data_as(C.c_void_p), cj.ctypes.data_as(C.c_void_p), nsys))
    got = host(dM)
    assert np.array_equal(_unpack(got, field, cols), want)
    # padding bits beyond `cols` are untouched
    bits = FIELD_BITS[field]
    used = (cols * bits + 7) // 8
    assert np.array_equal(got[..., used:], M[..., used:])
    # ... and on the device; swapping twice restores the matrix
    dci, dcj = dev(ci.view(np.int32)), dev(cj.view(np.int32))
    capi.check(cuda_lib.oblas_b200_swapcols_batch(field, _vp(dM), st, rows * st, rows, _vp(dci), _vp(dcj), nsys))
    assert np.array_equal(host(dM), M)


@pytest.mark.parametrize("field", [GF4, GF16, GF256])
def test_swapcol_matches_compiled_reference(field, cuda_lib):
    ref = reference("ref")
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rows, cols = 20, 77
    g = ref.gfmat_new(field, rows, cols)
    v = mat_view(g)
    st = v.shape[1]
    src = splitmix_bytes(60 + field, rows * st)
    v[:] = src.reshape(rows, st)
    ref.gfmat_swapcol(g, 3, 70)
    want = v.copy()
    ref.gfmat_free(g)
    dM = dev(src.reshape(rows, st).copy())
    i, j = np.array([3], dtype=np.uint32), np.array([70], dtype=np.uint32)
    capi.check(cuda_lib.oblas_b200_swapcols_batch(field, _vp(dM), st, 0, rows, i.ctypes.data_as(C.c_void_p),
                                                  j.ctypes.data_as(C.c_void_p), 1))
    assert np.array_equal(host(dM), want)


@pytest.mark.parametrize("field", [GF2, GF4, GF16, GF256])
def test_get_set_elements(field, cuda_lib):
    rows, cols = 29, 150
    M, st = _packed(field, rows, cols, 80 + field)
    M = M[0]
    bits = FIELD_BITS[field]
    rng = np.random.default_rng(field)
    flat = rng.choice(rows * cols, 400, replace=False)
    ri, cidx = (flat // cols).astype(np.uint32), (flat % cols).astype(np.uint32)
    elems = _unpack(M, field, cols)
    dM = dev(M)
    out = torch.zeros(len(flat), dtype=torch.uint8, device="cuda")
    capi.check(cuda_lib.oblas_b200_get_elements(field, _vp(dM), st, ri.ctypes.data_as(C.c_void_p),
                                                cidx.ctypes.data_as(C.c_void_p), len(flat), _vp(out)))
    assert np.array_equal(host(out), elems[ri, cidx])
    vals = rng.integers(0, 1 << bits, len(flat)).astype(np.uint8)
    capi.check(cuda_lib.oblas_b200_set_elements(field, _vp(dM), st, ri.ctypes.data_as(C.c_void_p),
                                                cidx.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p), len(flat)))
    elems[ri, cidx] = vals
    got = host(dM)
    assert np.array_equal(_unpack(got, field, cols), elems)
    used = (cols * bits + 7) // 8
    assert np.array_equal(got[:, used:], M[:, used:])
