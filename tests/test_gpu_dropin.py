"""The drop-in symbols of liboblas_b200.so run the SAME case functions that
generated tests/golden/golden.json from the unmodified reference (tests/cases.py):
identical calls, identical bytes.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest

import cases
from helpers import GF4, GF16, GF256, fnv1a, mat_view, oracle, ptr, splitmix_bytes, u32p
from oblas_b200 import capi

pytestmark = pytest.mark.gpu


def test_octmat_rowops_digest(cuda_lib, golden):
    got = cases.octmat_rowop_digest(cuda_lib)
    assert got == golden["rowops"]["octmat"]


def test_octmat_odd_stride(cuda_lib, golden):
    assert cases.octmat_odd_stride_digest(cuda_lib) == golden["rowops"]["octmat_odd_stride"]


@pytest.mark.parametrize("field", [GF4, GF16, GF256])
def test_gfmat_rowops_digest(cuda_lib, golden, field):
    assert cases.gfmat_rowop_digest(cuda_lib, field) == golden["rowops"]["gfmat"][str(field)]


def test_binmat_ops(cuda_lib, golden):
    assert cases.binmat_ops(cuda_lib) == golden["binmat"]


def test_b32(cuda_lib, golden):
    """oblas_ref.c semantics (one mask word per 32 bytes), not the AVX-512 build's"""
    got = cases.b32_digest(cuda_lib)
    assert got == golden["rowops"]["b32"]
    assert got != golden["rowops"]["b32_avx512_divergent"]


@pytest.mark.parametrize("field", [0, GF4, GF16, GF256])
def test_gfmat_aux(cuda_lib, golden, field):
    assert cases.gfmat_aux(cuda_lib, field) == golden["gfmat_aux"][str(field)]


def test_l1_any_length_any_pointer(cuda_lib, golden):
    got = cases.l1_ops(cuda_lib, lambda bits, which: np.frombuffer(
        capi.table("GF2_%d_%s" % (bits, {3: "SHUF_LO", 4: "SHUF_HI"}[which])), dtype=np.uint8))
    assert got == golden["l1"]


def test_layout_all_alignments(cuda_lib, golden):
    for al, ent in golden["layout"]["octmat_align"].items():
        for cols, stride in ent["octmat"]:
            m = cuda_lib.octmat_new_aligned(2, cols, int(al))
            assert m.contents.stride == stride and not mat_view(m).any()
            cuda_lib.octmat_free(m)
        for field, cols, stride, align in ent["gfmat"]:
            m = cuda_lib.gfmat_new_aligned(field, 2, cols, int(al))
            assert (m.contents.stride, m.contents.align) == (stride, align)
            cuda_lib.gfmat_free(m)
        for cols, stride in ent["binmat"]:
            m = cuda_lib.binmat_new(2, cols)
            assert m.contents.stride == stride
            cuda_lib.binmat_free(m)


def test_solve_cases(cuda_lib, golden):
    fns = cases.dropin_solve_fns(cuda_lib)
    for g in golden["solve"]:
        got = cases.solve_case(cuda_lib, tuple(g["case"]), fns)
        for k in ("rank", "pivcols", "A", "B", "a_bytes", "b_row_bytes"):
            assert got[k] == g[k], (g["case"], k)


def test_apply_cases(cuda_lib, golden):
    for g in golden["apply"]:
        assert cases.apply_case(cuda_lib, tuple(g["case"]), cuda_lib.octmat_apply)["Y"] == g["Y"], g["case"]


def test_batched_struct_api(cuda_lib):
    """oaxpy_batch & co. == the same operations one by one on the oracle"""
    o = oracle()
    rows, cols = 64, 1280
    a, b = cuda_lib.octmat_new(rows, cols), cuda_lib.octmat_new(rows, cols)
    va, vb = mat_view(a), mat_view(b)
    va[:] = splitmix_bytes(81, rows * cols).reshape(rows, cols)
    vb[:] = splitmix_bytes(82, rows * cols).reshape(rows, cols)
    A0, B0 = va.copy(), vb.copy()
    rng = np.random.default_rng(3)
    i = rng.permutation(rows).astype(np.uint32)
    j = rng.integers(0, rows, rows).astype(np.uint32)
    u = rng.integers(0, 256, rows).astype(np.uint8)
    cuda_lib.oaxpy_batch(a, b, ptr(i, u32p), ptr(j, u32p), ptr(u), rows)
    cuda_lib.oscal_batch(a, ptr(i, u32p), ptr(u), rows)
    cuda_lib.oaddrow_batch(a, b, ptr(i, u32p), ptr(j, u32p), rows)
    pi, pj = i[:20].copy(), i[20:40].copy()
    cuda_lib.oswaprow_batch(a, ptr(pi, u32p), ptr(pj, u32p), 20)
    for k in range(rows):
        o.orc_row_axpy(GF256, 0, ptr(A0), cols, ptr(B0), cols, i[k], j[k], u[k])
    for k in range(rows):
        o.orc_row_scal(GF256, 0, ptr(A0), cols, i[k], u[k])
    for k in range(rows):
        o.orc_row_add(ptr(A0), cols, ptr(B0), cols, i[k], j[k])
    for k in range(20):
        o.orc_row_swap(ptr(A0), cols, pi[k], pj[k])
    assert np.array_equal(mat_view(a), A0)
    cuda_lib.octmat_free(a)
    cuda_lib.octmat_free(b)


def test_solve_batch_struct_api(cuda_lib, golden):
    PO = C.POINTER(capi.Octmat)
    n, K, L = 5, 64, 272
    As = [cuda_lib.octmat_new(K, K) for _ in range(n)]
    Bs = [cuda_lib.octmat_new(K, L) for _ in range(n)]
    refA, refB, want = [], [], []
    o = oracle()
    for s in range(n):
        mat_view(As[s])[:] = splitmix_bytes(900 + s, K * K).reshape(K, K)
        mat_view(Bs[s])[:] = splitmix_bytes(950 + s, K * L).reshape(K, L)
        a0, b0 = mat_view(As[s]).copy(), mat_view(Bs[s]).copy()
        want.append(o.orc_solve(GF256, 1, ptr(a0), K, K, K, ptr(b0), L, None))
        refA.append(a0)
        refB.append(b0)
    rank = np.zeros(n, dtype=np.uint32)
    cuda_lib.octmat_solve_batch((PO * n)(*As), (PO * n)(*Bs), ptr(rank, u32p), n)
    assert rank.tolist() == want
    for s in range(n):
        assert np.array_equal(mat_view(As[s]), refA[s]) and np.array_equal(mat_view(Bs[s]), refB[s])
        cuda_lib.octmat_free(As[s])
        cuda_lib.octmat_free(Bs[s])


def test_launch_counter_moves(cuda_lib):
    before = cuda_lib.oblas_b200_launch_count()
    a = cuda_lib.octmat_new(2, 64)
    cuda_lib.oaddrow(a, a, 0, 1)
    assert cuda_lib.oblas_b200_launch_count() == before + 1
    cuda_lib.octmat_free(a)
