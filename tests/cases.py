"""Parity cases, written ONCE against the reference's struct API.

The same functions are run on
  * oracle/_ref/liboblas_<variant>.so  -> tests/golden/golden.json (make_golden.py)
  * oblas_b200/liboblas_b200.so        -> tests/test_gpu_dropin.py (needs a B200)
because the two libraries export the same ABI; that is what "drop-in" means.
`tests/test_oracle.py` re-derives the same numbers with the raw-buffer oracle.
TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from helpers import (GF2, GF4, GF16, GF256, FIELD_BITS, Binmat, Gfmat, Octmat, fnv1a, mat_view, ptr,
                     splitmix_bytes, u8p, u32p)

HEX = lambda v: "%016x" % v  # noqa: E731

# (field, rows, cols, b_bytes, seed, deficient)   deficient: 0 none, 1 dup row, 2 zero col
SOLVE_CASES = [
    (GF256, 24, 24, 48, 1001, 0),
    (GF256, 64, 64, 272, 1002, 0),
    (GF256, 64, 64, 0, 1003, 0),
    (GF256, 48, 48, 64, 1004, 1),
    (GF256, 48, 48, 64, 1005, 2),
    (GF256, 33, 50, 32, 1006, 0),
    (GF256, 50, 33, 32, 1007, 0),
    (GF256, 200, 200, 528, 1008, 0),
    (GF16, 70, 70, 48, 1011, 0),
    (GF16, 70, 70, 48, 1012, 2),
    (GF16, 150, 150, 0, 1013, 0),
    (GF4, 70, 70, 32, 1021, 0),
    (GF4, 150, 150, 16, 1022, 1),
    (GF2, 70, 70, 32, 1031, 0),
    (GF2, 300, 300, 32, 1032, 0),
    (GF2, 100, 300, 16, 1033, 0),
    (GF2, 300, 100, 16, 1034, 0),
]
APPLY_CASES = [(10, 10, 32, 2001), (40, 37, 272, 2002), (70, 100, 512, 2003), (33, 16, 256, 2004)]


def _fill_oct(lib, m, seed):
    v = mat_view(m)
    rows, stride = v.shape
    cols = m.contents.cols
    v[:, :cols] = splitmix_bytes(seed, rows * cols).reshape(rows, cols)


def _fill_packed(m, seed, field):
    """uniform field elements in the first `cols` positions, zero padding"""
    v = mat_view(m)
    rows = v.shape[0]
    cols = m.contents.cols
    bits = FIELD_BITS[field]
    used = (cols * bits + 7) // 8
    v[:, :used] = splitmix_bytes(seed, rows * used).reshape(rows, used)
    tail = cols * bits - (used - 1) * 8
    if tail < 8:
        v[:, used - 1] &= (1 << tail) - 1


def _make_deficient(v, field, cols, deficient):
    rows = v.shape[0]
    bits = FIELD_BITS[field]
    if deficient >= 1 and rows > 3:
        v[rows // 2] = v[1]
    if deficient >= 2 and cols > 5:
        e = 3
        byte, sh = (e * bits) // 8, (e * bits) % 8
        v[:, byte] &= np.uint8(~(((1 << bits) - 1) << sh) & 0xFF)


def octmat_rowop_digest(lib, sync=None):
    """SURVEY.md section 10 'cross-build digest test' on two 4x1280 octmats."""
    a, b = lib.octmat_new(4, 1280), lib.octmat_new(4, 1280)
    _fill_oct(lib, a, 11)
    _fill_oct(lib, b, 12)
    for u in range(256):
        lib.oaxpy(a, b, u & 3, (u >> 2) & 3, u)
        lib.oscal(a, u & 3, u)
    lib.oaddrow(a, b, 1, 2)
    lib.oswaprow(a, 0, 3)
    lib.oswaprow(a, 2, 2)
    lib.oaxpy(a, a, 1, 1, 7)  # same matrix, same row
    if sync:
        sync()
    out = {"stride": a.contents.stride, "digest": HEX(fnv1a(mat_view(a)))}
    lib.octmat_free(a)
    lib.octmat_free(b)
    return out


def octmat_odd_stride_digest(lib, sync=None):
    """cols = 1281 -> stride 1296 at ALIGN 16 (padding must stay consistent)"""
    a, b = lib.octmat_new(5, 1281), lib.octmat_new(5, 1281)
    _fill_oct(lib, a, 13)
    _fill_oct(lib, b, 14)
    for u in (0, 1, 2, 3, 29, 76, 128, 200, 255):
        lib.oaxpy(a, b, u % 5, (u + 2) % 5, u)
        lib.oscal(a, (u + 1) % 5, u)
    if sync:
        sync()
    out = {"stride": a.contents.stride, "digest": HEX(fnv1a(mat_view(a)))}
    lib.octmat_free(a)
    lib.octmat_free(b)
    return out


def gfmat_rowop_digest(lib, field, sync=None):
    cols = 2048
    a, b = lib.gfmat_new(field, 4, cols), lib.gfmat_new(field, 4, cols)
    _fill_packed(a, 21 + field, field)
    _fill_packed(b, 31 + field, field)
    n = 1 << FIELD_BITS[field]
    for u in range(n):
        lib.gfmat_axpy(a, b, u & 3, (u >> 2) & 3, u)
        lib.gfmat_scal(a, u & 3, u)
    lib.gfmat_add(a, b, 1, 2)
    lib.gfmat_swaprow(a, 0, 3)
    if sync:
        sync()
    d1 = HEX(fnv1a(mat_view(a)))
    lib.gfmat_zero(a, 2)
    if sync:
        sync()
    out = {"stride_words": a.contents.stride, "digest": d1, "digest_after_zero": HEX(fnv1a(mat_view(a)))}
    lib.gfmat_free(a)
    lib.gfmat_free(b)
    return out


def binmat_ops(lib, sync=None):
    a, b = lib.binmat_new(6, 200), lib.binmat_new(6, 200)
    va, vb = mat_view(a), mat_view(b)
    va[:, :25] = splitmix_bytes(41, 6 * 25).reshape(6, 25)
    vb[:, :25] = splitmix_bytes(42, 6 * 25).reshape(6, 25)
    lib.binmat_add(a, b, 0, 1)
    lib.binmat_add(a, a, 2, 3)
    lib.binmat_swaprow(a, 4, 5)
    lib.binmat_swaprow(a, 1, 1)
    lib.binmat_zero(a, 3)
    if sync:
        sync()
    lib.binmat_set(a, 3, 7, 0)   # quirk: sets the bit although b == 0 (binmat.c:41)
    lib.binmat_set(a, 3, 199, 1)
    lib.binmat_set(a, 3, 200, 1)  # out of range: ignored
    gets = [int(lib.binmat_get(a, r, c)) for r in (0, 3, 5) for c in (0, 7, 63, 64, 199, 200)]
    nnz = [int(lib.binmat_nnz(a, r, s, e)) for r in (0, 1, 4) for (s, e) in ((0, 200), (0, 64), (3, 20), (40, 100), (32, 64), (0, 201), (5, 4))]
    fill = np.zeros(a.contents.stride * 32, dtype=np.uint8)
    lib.binmat_fill(a, 0, ptr(fill))
    mask = splitmix_bytes(43, a.contents.stride * 4 // 8).view(np.uint32).copy()  # stride*4/32 words
    lib.binmat_expand(a, ptr(mask, u32p), 5, 0x5A)
    if sync:
        sync()
    out = {"stride_words": a.contents.stride, "digest": HEX(fnv1a(mat_view(a))), "gets": gets,
           "nnz": nnz, "fill_digest": HEX(fnv1a(fill)), "fill_marks": int((fill != 0).sum())}
    lib.binmat_free(a)
    lib.binmat_free(b)
    return out


def b32_digest(lib, sync=None):
    a = lib.octmat_new(3, 1280)
    _fill_oct(lib, a, 51)
    mask = splitmix_bytes(52, 1280 // 8).view(np.uint32).copy()
    lib.oaxpy_b32(a, ptr(mask, u32p), 1, 0x53)
    lib.oaxpy_b32(a, ptr(mask, u32p), 2, 0x01)
    if sync:
        sync()
    out = {"digest": HEX(fnv1a(mat_view(a)))}
    lib.octmat_free(a)
    return out


def gfmat_aux(lib, field, sync=None):
    cols = 100
    m = lib.gfmat_new(field, 5, cols)
    _fill_packed(m, 61 + field, field)
    n = 1 << FIELD_BITS[field]
    lib.gfmat_set(m, 1, 5, 3 % n)
    lib.gfmat_set(m, 1, 99, (n - 1))
    lib.gfmat_set(m, 1, 100, 1)        # out of range
    lib.gfmat_set(m, 2, 0, n + 1)      # reduced mod field size (gfmat.c:97)
    gets = [int(lib.gfmat_get(m, r, c)) for r in (0, 1, 2) for c in (0, 5, 31, 32, 99, 100)]
    nnz = [int(lib.gfmat_nnz(m, r, s, e)) for r in (0, 3) for (s, e) in ((0, 100), (0, 33), (3, 20), (40, 100), (16, 32), (0, 101))]
    vpw = 32 // FIELD_BITS[field]
    fill = np.zeros(m.contents.stride * vpw, dtype=np.uint8)
    lib.gfmat_fill(m, 0, ptr(fill))
    lib.gfmat_swapcol(m, 2, 77)
    lib.gfmat_swapcol(m, 9, 9)
    out = {"stride_words": m.contents.stride, "align": m.contents.align, "gets": gets, "nnz": nnz,
           "fill_digest": HEX(fnv1a(fill)), "digest": HEX(fnv1a(mat_view(m)))}
    lib.gfmat_free(m)
    return out


def l1_ops(lib, table_fn, sync=None):
    """oblas_xor/axpy/scal/swap on raw buffers incl. lengths that are not multiples of 16"""
    res = {}
    lo8, hi8 = table_fn(8, 3), table_fn(8, 4)
    for k in (1, 15, 16, 48, 100, 1280, 4099):
        a = splitmix_bytes(71, k).copy()
        b = splitmix_bytes(72, k).copy()
        u = 0x8E
        lo = lo8[u * 16:u * 16 + 16].copy()
        hi = hi8[u * 16:u * 16 + 16].copy()
        lib.oblas_axpy(ptr(a), ptr(b), k, ptr(lo), ptr(hi))
        lib.oblas_scal(ptr(a), k, ptr(lo), ptr(hi))
        lib.oblas_xor(ptr(a), ptr(b), k)
        lib.oblas_swap(ptr(a), ptr(b), k)
        if sync:
            sync()
        res[str(k)] = [HEX(fnv1a(a)), HEX(fnv1a(b))]
    return res


def solve_case(lib, case, solve_fns):
    """solve_fns: dict with 'oct', 'gf', 'bin' callables (A, B, pivptr) -> rank"""
    field, rows, cols, b_bytes, seed, deficient = case
    piv = np.full(rows, 0xFFFFFFFF, dtype=np.uint32)
    if field == GF256:
        A = lib.octmat_new(rows, cols)
        B = lib.octmat_new(rows, b_bytes) if b_bytes else None
        _fill_oct(lib, A, seed)
        _make_deficient(mat_view(A), field, cols, deficient)
        if B:
            _fill_oct(lib, B, seed + 500)
        rank = solve_fns["oct"](A, B, ptr(piv, u32p))
        free = lib.octmat_free
    elif field == GF2:
        A = lib.binmat_new(rows, cols)
        B = lib.binmat_new(rows, b_bytes * 8) if b_bytes else None
        _fill_packed(A, seed, field)
        _make_deficient(mat_view(A), field, cols, deficient)
        if B:
            _fill_packed(B, seed + 500, field)
        rank = solve_fns["bin"](A, B, ptr(piv, u32p))
        free = lib.binmat_free
    else:
        bits = FIELD_BITS[field]
        A = lib.gfmat_new(field, rows, cols)
        B = lib.gfmat_new(field, rows, b_bytes * 8 // bits) if b_bytes else None
        _fill_packed(A, seed, field)
        _make_deficient(mat_view(A), field, cols, deficient)
        if B:
            _fill_packed(B, seed + 500, field)
        rank = solve_fns["gf"](A, B, ptr(piv, u32p))
        free = lib.gfmat_free
    out = {"rank": int(rank), "pivcols": [int(x) for x in piv[:rank]],
           "A": HEX(fnv1a(mat_view(A))), "B": HEX(fnv1a(mat_view(B))) if B else None,
           "a_bytes": int(mat_view(A).shape[1]), "b_row_bytes": int(mat_view(B).shape[1]) if B else 0}
    free(A)
    if B:
        free(B)
    return out


def apply_case(lib, case, apply_fn):
    M, Kc, nbytes, seed = case
    Cm, D, Y = lib.octmat_new(M, Kc), lib.octmat_new(Kc, nbytes), lib.octmat_new(M, nbytes)
    _fill_oct(lib, Cm, seed)
    _fill_oct(lib, D, seed + 1)
    apply_fn(Cm, D, Y)
    out = {"Y": HEX(fnv1a(mat_view(Y)))}
    for m in (Cm, D, Y):
        lib.octmat_free(m)
    return out


def ref_solve_fns(lib):
    return {"oct": lambda A, B, p: lib.refdrv_octmat_solve(A, B, p),
            "gf": lambda A, B, p: lib.refdrv_gfmat_solve(A, B, p, 1),
            "bin": lambda A, B, p: lib.refdrv_binmat_rref(A, B, p)}


def dropin_solve_fns(lib):
    return {"oct": lambda A, B, p: lib.octmat_solve(A, B, p),
            "gf": lambda A, B, p: lib.gfmat_solve(A, B, p),
            "bin": lambda A, B, p: lib.binmat_rref(A, B, p)}
