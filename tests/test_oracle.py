"""The oracle (oracle/oblas_oracle.c) against the golden fixtures generated from the
unmodified reference, and against the compiled reference itself when it is present
(build container).  CPU only."""
import hashlib

import numpy as np
import pytest

import cases
from helpers import (GF2, GF4, GF16, GF256, FIELD_BITS, field_row_bytes, fnv1a, oracle, oracle_solve,
                     oracle_table, ptr, random_matrix, ref_table, reference, splitmix_bytes, u32p)

HEX = cases.HEX
NAMES = {0: "LOG", 1: "EXP", 2: "INV", 3: "SHUF_LO", 4: "SHUF_HI"}


def test_tables_match_golden(golden):
    for bits, field in ((2, GF4), (4, GF16), (8, GF256)):
        for which, nm in NAMES.items():
            t = oracle_table(field, which, 0)
            g = golden["tables"]["GF2_%d_%s" % (bits, nm)]
            assert len(t) == g["len"]
            assert hashlib.sha256(t.tobytes()).hexdigest() == g["sha256"], (bits, nm)
    fix = oracle_table(GF4, 4, 1)
    assert hashlib.sha256(fix.tobytes()).hexdigest() == golden["tables"]["GF2_2_SHUF_HI_fixed"]["sha256"]


def test_table_selfcheck_counts(golden):
    """GF256 0 bad, GF16 0 bad, GF4 144 bad (shipped HI table)."""
    o = oracle()
    for bits, field in ((2, GF4), (4, GF16), (8, GF256)):
        lo, hi = oracle_table(field, 3, 0), oracle_table(field, 4, 0)
        n, bad = 1 << bits, 0
        for u in range(n):
            for x in range(256):
                got = hi[u * 16 + (x >> 4)] ^ lo[u * 16 + (x & 15)]
                want = 0
                for k in range(8 // bits):
                    want |= o.orc_gf_mul(field, u, (x >> (k * bits)) & (n - 1)) << (k * bits)
                bad += int(got != want)
        assert bad == golden["table_selfcheck_bad"][str(bits)]
    assert golden["table_selfcheck_bad"] == {"2": 144, "4": 0, "8": 0}


def test_layout_matches_golden(golden):
    o = oracle()
    assert golden["layout"]["sizeof"] == [24, 32, 24]
    assert golden["layout"]["bits_offset"] == [16, 24, 16]
    for al, ent in golden["layout"]["octmat_align"].items():
        al = int(al)
        for cols, stride in ent["octmat"]:
            assert o.orc_octmat_stride(cols, al) == stride
        for field, cols, stride, align in ent["gfmat"]:
            assert o.orc_gfmat_stride(field, cols, al) == stride
        for cols, stride in ent["binmat"]:
            assert o.orc_binmat_stride(cols) == stride
    e16 = golden["layout"]["octmat_align"]["16"]
    assert [1281, 1296] in e16["octmat"]


def _oracle_octmat_digest():
    o = oracle()
    st = 1280
    a = splitmix_bytes(11, 4 * 1280).reshape(4, st).copy()
    b = splitmix_bytes(12, 4 * 1280).reshape(4, st).copy()
    for u in range(256):
        o.orc_row_axpy(GF256, 0, ptr(a), st, ptr(b), st, u & 3, (u >> 2) & 3, u)
        o.orc_row_scal(GF256, 0, ptr(a), st, u & 3, u)
    o.orc_row_add(ptr(a), st, ptr(b), st, 1, 2)
    o.orc_row_swap(ptr(a), st, 0, 3)
    o.orc_row_swap(ptr(a), st, 2, 2)
    o.orc_row_axpy(GF256, 0, ptr(a), st, ptr(a), st, 1, 1, 7)
    return HEX(fnv1a(a))


def test_octmat_rowops_digest(golden):
    assert _oracle_octmat_digest() == golden["rowops"]["octmat"]["digest"]


def test_gfmat_rowops_digest(golden):
    o = oracle()
    for field in (GF4, GF16, GF256):
        cols = 2048
        st = field_row_bytes(field, cols)
        assert st == 4 * golden["rowops"]["gfmat"][str(field)]["stride_words"]
        used = cols * FIELD_BITS[field] // 8
        a = np.zeros((4, st), dtype=np.uint8)
        b = np.zeros((4, st), dtype=np.uint8)
        a[:, :used] = splitmix_bytes(21 + field, 4 * used).reshape(4, used)
        b[:, :used] = splitmix_bytes(31 + field, 4 * used).reshape(4, used)
        for u in range(1 << FIELD_BITS[field]):
            o.orc_row_axpy(field, 0, ptr(a), st, ptr(b), st, u & 3, (u >> 2) & 3, u)
            o.orc_row_scal(field, 0, ptr(a), st, u & 3, u)
        o.orc_row_add(ptr(a), st, ptr(b), st, 1, 2)
        o.orc_row_swap(ptr(a), st, 0, 3)
        assert HEX(fnv1a(a)) == golden["rowops"]["gfmat"][str(field)]["digest"], field
        o.orc_row_zero(ptr(a), st, 2)
        assert HEX(fnv1a(a)) == golden["rowops"]["gfmat"][str(field)]["digest_after_zero"]


def test_b32_digest(golden):
    o = oracle()
    a = splitmix_bytes(51, 3 * 1280).reshape(3, 1280).copy()
    mask = splitmix_bytes(52, 1280 // 8).view(np.uint32).copy()
    o.orc_row_axpy_b32(ptr(a), 1280, ptr(mask, u32p), 1, 0x53)
    o.orc_row_axpy_b32(ptr(a), 1280, ptr(mask, u32p), 2, 0x01)
    assert HEX(fnv1a(a)) == golden["rowops"]["b32"]["digest"]
    # documented divergence of the AVX-512 build
    if golden["rowops"]["b32_avx512_divergent"]:
        assert golden["rowops"]["b32_avx512_divergent"]["digest"] != golden["rowops"]["b32"]["digest"]


def test_aux_quirks(golden):
    """get/set/nnz/fill/swapcol incl. the reference's slips"""
    o = oracle()
    for field in (0, GF4, GF16, GF256):
        g = golden["gfmat_aux"][str(field)]
        cols, rows = 100, 5
        stw = g["stride_words"]
        bits = FIELD_BITS[field]
        m = np.zeros((rows, stw), dtype=np.uint32)
        used = (cols * bits + 7) // 8
        mb = m.view(np.uint8).reshape(rows, stw * 4)
        mb[:, :used] = splitmix_bytes(61 + field, rows * used).reshape(rows, used)
        tail = cols * bits - (used - 1) * 8
        if tail < 8:
            mb[:, used - 1] &= (1 << tail) - 1
        n = 1 << bits
        P = ptr(m, u32p)
        o.orc_set(field, P, stw, rows, cols, 1, 5, 3 % n)
        o.orc_set(field, P, stw, rows, cols, 1, 99, n - 1)
        o.orc_set(field, P, stw, rows, cols, 1, 100, 1)
        o.orc_set(field, P, stw, rows, cols, 2, 0, n + 1)
        gets = [int(o.orc_get(field, P, stw, rows, cols, r, c)) for r in (0, 1, 2) for c in (0, 5, 31, 32, 99, 100)]
        assert gets == g["gets"], field
        nnz = [int(o.orc_nnz(field, P, stw, rows, cols, r, s, e)) for r in (0, 3)
               for (s, e) in ((0, 100), (0, 33), (3, 20), (40, 100), (16, 32), (0, 101))]
        assert nnz == g["nnz"], field
        vpw = 32 // bits
        fill = np.zeros(stw * vpw, dtype=np.uint8)
        o.orc_fill(field, P, stw, 0, ptr(fill))
        assert HEX(fnv1a(fill)) == g["fill_digest"]
        o.orc_swapcol(field, P, stw, rows, 2, 77)
        o.orc_swapcol(field, P, stw, rows, 9, 9)
        assert HEX(fnv1a(mb)) == g["digest"], field


def test_binmat_quirks(golden):
    o = oracle()
    g = golden["binmat"]
    stw = g["stride_words"]
    st = stw * 4
    a = np.zeros((6, st), dtype=np.uint8)
    b = np.zeros((6, st), dtype=np.uint8)
    a[:, :25] = splitmix_bytes(41, 150).reshape(6, 25)
    b[:, :25] = splitmix_bytes(42, 150).reshape(6, 25)
    o.orc_row_add(ptr(a), st, ptr(b), st, 0, 1)
    o.orc_row_add(ptr(a), st, ptr(a), st, 2, 3)
    o.orc_row_swap(ptr(a), st, 4, 5)
    o.orc_row_zero(ptr(a), st, 3)
    A32 = a.view(np.uint32)
    P = ptr(A32, u32p)
    o.orc_binmat_set(P, stw, 6, 200, 3, 7, 0)
    o.orc_binmat_set(P, stw, 6, 200, 3, 199, 1)
    o.orc_binmat_set(P, stw, 6, 200, 3, 200, 1)
    gets = [int(o.orc_get(GF2, P, stw, 6, 200, r, c)) for r in (0, 3, 5) for c in (0, 7, 63, 64, 199, 200)]
    assert gets == g["gets"]
    nnz = [int(o.orc_nnz(GF2, P, stw, 6, 200, r, s, e)) for r in (0, 1, 4)
           for (s, e) in ((0, 200), (0, 64), (3, 20), (40, 100), (32, 64), (0, 201), (5, 4))]
    assert nnz == g["nnz"]
    fill = np.zeros(stw * 32, dtype=np.uint8)
    o.orc_fill(GF2, P, stw, 0, ptr(fill))
    assert HEX(fnv1a(fill)) == g["fill_digest"] and int((fill != 0).sum()) == g["fill_marks"]
    mask = splitmix_bytes(43, st // 8).view(np.uint32).copy()
    o.orc_row_axpy_b32(ptr(a), st, ptr(mask, u32p), 5, 0x5A)
    assert HEX(fnv1a(a)) == g["digest"]


def test_l1_odd_lengths(golden):
    o = oracle()
    lo8, hi8 = oracle_table(GF256, 3), oracle_table(GF256, 4)
    for k, want in golden["l1"].items():
        k = int(k)
        a, b = splitmix_bytes(71, k).copy(), splitmix_bytes(72, k).copy()
        lo, hi = lo8[0x8E * 16:0x8E * 16 + 16].copy(), hi8[0x8E * 16:0x8E * 16 + 16].copy()
        o.orc_axpy(ptr(a), ptr(b), k, ptr(lo), ptr(hi))
        o.orc_scal(ptr(a), k, ptr(lo), ptr(hi))
        o.orc_xor(ptr(a), ptr(b), k)
        o.orc_swap(ptr(a), ptr(b), k)
        assert [HEX(fnv1a(a)), HEX(fnv1a(b))] == want


def oracle_solve_case(case):
    field, rows, cols, b_bytes, seed, deficient = case
    A = random_matrix(field, rows, cols, seed, rank_deficient=deficient)
    B = None
    if b_bytes:
        if field == GF256:
            nb = field_row_bytes(GF256, b_bytes)
            B = np.zeros((rows, nb), dtype=np.uint8)
            B[:, :b_bytes] = splitmix_bytes(seed + 500, rows * b_bytes).reshape(rows, b_bytes)
        else:
            B = random_matrix(field, rows, b_bytes * 8 // FIELD_BITS[field], seed + 500)
    rank, piv = oracle_solve(field, A, B, cols)
    return rank, piv, A, B


def test_solve_matches_golden(golden):
    for g in golden["solve"]:
        case = tuple(g["case"])
        rank, piv, A, B = oracle_solve_case(case)
        assert rank == g["rank"], case
        assert [int(x) for x in piv[:rank]] == g["pivcols"], case
        assert A.shape[1] == g["a_bytes"]
        assert HEX(fnv1a(A)) == g["A"], case
        if B is not None:
            assert HEX(fnv1a(B)) == g["B"], case


def test_solve_full_rank_gives_identity():
    rank, piv, A, B = oracle_solve_case((GF256, 64, 64, 272, 1002, 0))
    assert rank == 64 and np.array_equal(A[:, :64], np.eye(64, dtype=np.uint8))


def test_apply_matches_golden(golden):
    o = oracle()
    for g in golden["apply"]:
        M, Kc, nbytes, seed = g["case"]
        cs, ds = field_row_bytes(GF256, Kc), field_row_bytes(GF256, nbytes)
        Cm = np.zeros((M, cs), dtype=np.uint8)
        Cm[:, :Kc] = splitmix_bytes(seed, M * Kc).reshape(M, Kc)
        D = np.zeros((Kc, ds), dtype=np.uint8)
        D[:, :nbytes] = splitmix_bytes(seed + 1, Kc * nbytes).reshape(Kc, nbytes)
        Y = np.zeros((M, ds), dtype=np.uint8)
        o.orc_gf256_apply(ptr(Cm), cs, M, Kc, ptr(D), ds, ptr(Y), ds, ds)
        assert HEX(fnv1a(Y)) == g["Y"]


def test_splitmix_generators_agree():
    o = oracle()
    for off, n in ((0, 100), (5, 64), (8, 8), (1234567, 333)):
        a = np.zeros(n, dtype=np.uint8)
        o.orc_fill_random(ptr(a), n, 99, off)
        assert np.array_equal(a, splitmix_bytes(99, n, off))
    assert o.orc_fnv1a(ptr(a), n) == fnv1a(a)


# ---- direct comparison with the compiled reference (build container only) ----------
@pytest.mark.parametrize("variant", ["ref", "avx2", "avx512"])
def test_reference_reproduces_golden(golden, variant):
    lib = reference(variant)
    if lib is None:
        pytest.skip("oracle/_ref/liboblas_%s.so not available on this host" % variant)
    assert cases.octmat_rowop_digest(lib)["digest"] == golden["rowops"]["octmat"]["digest"]
    for f in (GF4, GF16, GF256):
        assert cases.gfmat_rowop_digest(lib, f)["digest"] == golden["rowops"]["gfmat"][str(f)]["digest"]
    got = cases.b32_digest(lib)["digest"]
    if variant == "avx512":
        assert got == golden["rowops"]["b32_avx512_divergent"]["digest"]
    else:
        assert got == golden["rowops"]["b32"]["digest"]
    for bits in (2, 4, 8):
        for which, nm in NAMES.items():
            t = ref_table(lib, bits, which)
            assert hashlib.sha256(t.tobytes()).hexdigest() == golden["tables"]["GF2_%d_%s" % (bits, nm)]["sha256"]
    if variant == "ref":
        for g in golden["solve"]:
            got = cases.solve_case(lib, tuple(g["case"]), cases.ref_solve_fns(lib))
            assert {k: got[k] for k in ("rank", "pivcols", "A", "B")} == {k: g[k] for k in ("rank", "pivcols", "A", "B")}
