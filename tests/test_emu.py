"""Kernel LOGIC of the product's CUDA sources, executed on host threads
(tests/emu): row ops, blocked elimination, apply -- bit-exact against the oracle.
This is not the product path (the GPU tests are); it lets the GPU-less build
container catch indexing / pivoting / table-build mistakes."""
import ctypes as C

import numpy as np
import pytest

from helpers import (GF2, GF4, GF16, GF256, FIELD_BITS, MUL_LIN, MUL_LUT, OP_ADD, OP_AXPY, OP_AXPY_B32,
                     OP_SCAL, OP_SWAP, OP_ZERO, TS_GF4_FIXED, TS_GF4_SHIPPED, TS_GF16, TS_GF256, emu,
                     i32p, oracle, oracle_solve, ptr, random_matrix, splitmix_bytes, u32p)

SMEM = 227 * 1024


def test_fastdiv_exact():
    e = emu()
    for d in (1, 2, 3, 5, 16, 80, 81, 640, 1000, 4096, 65536, 12345, 0x7FFFFFF):
        assert e.emu_fastdiv_check(d, 2_000_000, 7) == 1, d


def test_update_tile_geometry():
    """tile_geo.h: the column tiles of the tensor-core update launches.  Both geometries cover the range exactly once
    with widths that are multiples of 16 and at most 240; the even one (the default) has no narrow rest tile: widths
    differ by at most 16 columns.  All ranges of K = 1024 / L = 1280 and a sweep."""
    e = emu()
    cap = 512
    starts, widths = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
    for vend in [2176, 2048, 1920, 1792, 1664, 1536, 1408, 1280] + list(range(16, 4000, 16)) + [65536, 100000 // 16 * 16]:
        for even in (0, 1):
            n = e.emu_tile_geo(vend, even, starts, widths, cap)
            assert n == (vend + 239) // 240
            pos = 0
            for t in range(n):
                assert starts[t] == pos and widths[t] % 16 == 0 and 16 <= widths[t] <= 240, (vend, even, t)
                pos += widths[t]
            assert pos == vend
            if even:
                assert max(widths[:n]) - min(widths[:n]) <= 16 and list(widths[:n]) == sorted(widths[:n], reverse=True)
            else:
                assert all(w == 240 for w in widths[:n - 1])
    n = e.emu_tile_geo(2176, 1, starts, widths, cap)
    assert list(widths[:n]) == [224] * 6 + [208] * 4


def test_table_linearity_flags():
    e = emu()
    assert [e.emu_table_linear(t) for t in range(4)] == [0, 1, 1, 1]


@pytest.mark.parametrize("field,ts,variant", [(GF256, TS_GF256, 0), (GF16, TS_GF16, 0),
                                              (GF4, TS_GF4_FIXED, 1), (GF4, TS_GF4_SHIPPED, 0)])
@pytest.mark.parametrize("mode", [MUL_LIN, MUL_LUT])
def test_rowops(field, ts, variant, mode):
    if ts == TS_GF4_SHIPPED and mode == MUL_LIN:
        pytest.skip("shipped GF(4) HI table is not linear")
    o, e = oracle(), emu()
    rng = np.random.default_rng(field * 10 + mode)
    rows = 12
    for op in (OP_AXPY, OP_ADD, OP_SCAL, OP_ZERO):
        for stride in (16, 48, 1280):
            nops = 9
            A = splitmix_bytes(10 + op, rows * stride).reshape(rows, stride).copy()
            Bm = splitmix_bytes(20 + op, rows * stride).reshape(rows, stride).copy()
            dst = rng.permutation(rows)[:nops].astype(np.uint32)
            src = rng.integers(0, rows, nops).astype(np.uint32)
            u = rng.integers(0, 1 << FIELD_BITS[field], nops).astype(np.uint8)
            u[0], u[1] = 0, 1
            A0, A1 = A.copy(), A.copy()
            for k in range(nops):
                if op == OP_AXPY:
                    o.orc_row_axpy(field, variant, ptr(A0), stride, ptr(Bm), stride, dst[k], src[k], u[k])
                elif op == OP_ADD:
                    o.orc_row_add(ptr(A0), stride, ptr(Bm), stride, dst[k], src[k])
                elif op == OP_SCAL:
                    o.orc_row_scal(field, variant, ptr(A0), stride, dst[k], u[k])
                else:
                    o.orc_row_zero(ptr(A0), stride, dst[k])
            e.emu_rowops(op, mode, ts, ptr(A1), stride, ptr(Bm), stride, stride, ptr(dst, u32p),
                         ptr(src, u32p), ptr(u), 0, nops)
            assert np.array_equal(A0, A1), (op, stride)


@pytest.mark.parametrize("mode", [MUL_LIN, MUL_LUT])
def test_axpy_every_scalar_every_byte(mode):
    o, e = oracle(), emu()
    stride = rows = 256
    A = splitmix_bytes(5, rows * stride).reshape(rows, stride).copy()
    Bm = np.tile(np.arange(256, dtype=np.uint8), (rows, 1))
    idx = np.arange(rows, dtype=np.uint32)
    u = np.arange(256, dtype=np.uint8)
    A0, A1 = A.copy(), A.copy()
    for k in range(rows):
        o.orc_row_axpy(GF256, 0, ptr(A0), stride, ptr(Bm), stride, k, k, k)
    e.emu_rowops(OP_AXPY, mode, TS_GF256, ptr(A1), stride, ptr(Bm), stride, stride, ptr(idx, u32p),
                 ptr(idx, u32p), ptr(u), 0, rows)
    assert np.array_equal(A0, A1)


def test_swap_and_b32():
    o, e = oracle(), emu()
    A = splitmix_bytes(7, 10 * 64).reshape(10, 64).copy()
    A0, A1 = A.copy(), A.copy()
    i = np.array([0, 2, 5, 7], dtype=np.uint32)
    j = np.array([1, 2, 9, 3], dtype=np.uint32)
    for k in range(4):
        o.orc_row_swap(ptr(A0), 64, i[k], j[k])
    e.emu_rowops(OP_SWAP, 0, -1, ptr(A1), 64, ptr(A1), 64, 64, ptr(i, u32p), ptr(j, u32p), None, 0, 4)
    assert np.array_equal(A0, A1)
    A = splitmix_bytes(8, 6 * 96).reshape(6, 96).copy()
    A0, A1 = A.copy(), A.copy()
    masks = splitmix_bytes(9, 6 * 12).view(np.uint32).reshape(6, 3).copy()
    i = np.array([0, 3, 4], dtype=np.uint32)
    j = np.array([5, 1, 1], dtype=np.uint32)
    u = np.array([0x53, 1, 0xFF], dtype=np.uint8)
    for k in range(3):
        o.orc_row_axpy_b32(ptr(A0), 96, ptr(masks[j[k]], u32p), i[k], u[k])
    e.emu_rowops(OP_AXPY_B32, 0, -1, ptr(A1), 96, ptr(masks.view(np.uint8)), 12, 96, ptr(i, u32p),
                 ptr(j, u32p), ptr(u), 0, 3)
    assert np.array_equal(A0, A1)


def _solve(field, rows, cols, b_bytes, seed, nsys=1, deficient=0, threads=64, grid=1, nbw=0, geo=-1, no_fast=0):
    e = emu()
    A = np.stack([random_matrix(field, rows, cols, seed + 2 * s, rank_deficient=deficient) for s in range(nsys)])
    B = np.stack([splitmix_bytes(seed + 2 * s + 1, rows * b_bytes).reshape(rows, b_bytes)
                  for s in range(nsys)]).copy() if b_bytes else None
    A0, B0 = A.copy(), (B.copy() if B is not None else None)
    want = [oracle_solve(field, A0[s], B0[s] if B0 is not None else None, cols) for s in range(nsys)]
    rank = np.zeros(nsys, dtype=np.int32)
    piv = np.full((nsys, rows), 0xFFFFFFFF, dtype=np.uint32)
    rc = e.emu_solve(field, ptr(A), A.strides[1], A.strides[0], rows, cols, A.shape[2],
                     ptr(B) if B is not None else None, B.strides[1] if B is not None else 0,
                     B.strides[0] if B is not None else 0, B.shape[2] if B is not None else 0,
                     ptr(rank, i32p), ptr(piv, u32p), nsys, grid, threads, SMEM, nbw, geo, no_fast)
    assert rc == 0
    assert [int(r) for r in rank] == [w[0] for w in want]
    for s in range(nsys):
        assert np.array_equal(piv[s][:want[s][0]], want[s][1][:want[s][0]])
    assert np.array_equal(A, A0)
    if B is not None:
        assert np.array_equal(B, B0)
    return [w[0] for w in want]


@pytest.mark.parametrize("args", [
    dict(field=GF256, rows=20, cols=20, b_bytes=48, seed=100),
    dict(field=GF256, rows=40, cols=40, b_bytes=272, seed=101),
    dict(field=GF256, rows=40, cols=40, b_bytes=0, seed=102),
    dict(field=GF256, rows=40, cols=40, b_bytes=64, seed=103, deficient=1),
    dict(field=GF256, rows=40, cols=40, b_bytes=64, seed=104, deficient=2),
    dict(field=GF256, rows=33, cols=50, b_bytes=32, seed=105),
    dict(field=GF256, rows=50, cols=33, b_bytes=32, seed=106),
    dict(field=GF256, rows=40, cols=40, b_bytes=64, seed=107, nsys=3, grid=2),
    dict(field=GF256, rows=40, cols=40, b_bytes=64, seed=108, nbw=2),
    dict(field=GF256, rows=150, cols=150, b_bytes=528, seed=110, threads=128),
    dict(field=GF16, rows=70, cols=70, b_bytes=48, seed=111),
    dict(field=GF16, rows=70, cols=70, b_bytes=48, seed=112, deficient=2),
    dict(field=GF4, rows=70, cols=70, b_bytes=32, seed=114),
    dict(field=GF4, rows=150, cols=150, b_bytes=16, seed=115, deficient=1),
    dict(field=GF4, rows=150, cols=150, b_bytes=16, seed=116, nbw=2),
    dict(field=GF2, rows=70, cols=70, b_bytes=32, seed=117),
    dict(field=GF2, rows=200, cols=200, b_bytes=32, seed=118, threads=128),
    dict(field=GF2, rows=200, cols=200, b_bytes=0, seed=119, nbw=2, threads=128),
    dict(field=GF2, rows=100, cols=300, b_bytes=16, seed=120),
    dict(field=GF2, rows=300, cols=100, b_bytes=16, seed=121),
    dict(field=GF256, rows=1, cols=1, b_bytes=16, seed=122),
    dict(field=GF256, rows=5, cols=700, b_bytes=16, seed=123),
])
@pytest.mark.parametrize("geo", [0, 1, 2, 4])
def test_solve(args, geo):
    _solve(geo=geo, **args)


@pytest.mark.parametrize("args", [
    dict(field=GF256, rows=40, cols=40, b_bytes=64, seed=130),
    dict(field=GF256, rows=150, cols=150, b_bytes=528, seed=131, threads=128, deficient=1),
    dict(field=GF16, rows=70, cols=70, b_bytes=48, seed=132, deficient=2),
    dict(field=GF4, rows=150, cols=150, b_bytes=16, seed=133, nsys=3, grid=2),
    dict(field=GF2, rows=200, cols=200, b_bytes=32, seed=134, threads=128),
    dict(field=GF2, rows=300, cols=100, b_bytes=0, seed=135, deficient=1),
])
@pytest.mark.parametrize("geo", [1, 4])
@pytest.mark.parametrize("no_fast", [0, 1], ids=["fast_panel", "general_panel"])
def test_solve_tall_variant_rows_in_global_memory(args, geo, no_fast):
    """solve_kernel<.., NBW=4, GROWS>: the per-row state (panel chunk, coefficient vector, row map, candidate flags)
    in a per-CTA block of a workspace outside shared memory; every field, 6- and 5-bit tables, both panel paths"""
    _solve(nbw=40, geo=geo, no_fast=no_fast, **args)


@pytest.mark.parametrize("args", [
    dict(field=GF256, rows=100, cols=100, b_bytes=272, seed=201, threads=64),
    dict(field=GF256, rows=100, cols=100, b_bytes=64, seed=202, threads=64, deficient=1),
    dict(field=GF256, rows=100, cols=100, b_bytes=64, seed=203, threads=64, deficient=2),
    dict(field=GF256, rows=70, cols=120, b_bytes=32, seed=204, threads=64),
    dict(field=GF16, rows=150, cols=150, b_bytes=48, seed=205, threads=64),
    dict(field=GF16, rows=150, cols=150, b_bytes=48, seed=206, threads=64, nbw=2),
    dict(field=GF4, rows=200, cols=200, b_bytes=32, seed=207, threads=96),
    dict(field=GF4, rows=200, cols=200, b_bytes=32, seed=208, threads=64, nbw=2),
    dict(field=GF2, rows=400, cols=400, b_bytes=32, seed=209, threads=160),
    dict(field=GF2, rows=300, cols=300, b_bytes=0, seed=210, threads=96, nbw=2),
    dict(field=GF2, rows=300, cols=500, b_bytes=16, seed=211, threads=160),
])
@pytest.mark.parametrize("no_fast", [0, 1])
def test_solve_fast_and_general_panel_paths(args, no_fast):
    """block sizes large enough for the register-resident fast panel path (NPIV+32 threads),
    and the same systems forced through the general path"""
    _solve(no_fast=no_fast, **args)


@pytest.mark.parametrize("args", [
    # blocks with streaming threads beside the candidate group (NBT = 64 / 64 / 96 / 160 threads for
    # GF(256) / GF(16) / GF(4) / GF(2)): the next panel's candidate elimination runs as a resumable chain
    # beside the tile passes of the trailing update (several A tiles and B tiles per panel)
    dict(field=GF256, rows=150, cols=150, b_bytes=528, seed=501, threads=128),
    dict(field=GF256, rows=150, cols=150, b_bytes=0, seed=502, threads=128),
    dict(field=GF256, rows=120, cols=200, b_bytes=64, seed=503, threads=128, deficient=2),
    dict(field=GF256, rows=200, cols=150, b_bytes=272, seed=504, threads=192, nsys=2),
    dict(field=GF16, rows=200, cols=200, b_bytes=48, seed=505, threads=128),
    dict(field=GF16, rows=200, cols=200, b_bytes=48, seed=506, threads=128, deficient=1),
    dict(field=GF4, rows=260, cols=260, b_bytes=32, seed=507, threads=160),
    dict(field=GF4, rows=260, cols=260, b_bytes=0, seed=508, threads=160, deficient=2),
    dict(field=GF2, rows=400, cols=400, b_bytes=32, seed=509, threads=224),
    dict(field=GF2, rows=300, cols=500, b_bytes=16, seed=510, threads=224),
    dict(field=GF16, rows=200, cols=200, b_bytes=48, seed=511, threads=128, nbw=2),
])
@pytest.mark.parametrize("no_fast", [0, 2])
def test_solve_look_ahead_chain(args, no_fast):
    """table kernel with the look-ahead chain on (0) and off (2): same bytes as the oracle"""
    _solve(no_fast=no_fast, **args)


def test_solve_every_row_moves():
    """zero above the anti-diagonal: column c has its first non-zero entry in row n-1-c, so every row
    changes position -- the final row permutation stages all rows (several column blocks) instead of
    the usual handful; table kernel and outer-panel path"""
    e = emu()
    rows = cols = 96
    b_bytes = 48
    A = random_matrix(GF256, rows, cols, 301)[None].copy()
    for i in range(rows):
        A[0, i, :cols - 1 - i] = 0
    A[0, np.arange(rows), cols - 1 - np.arange(rows)] |= 1
    B = splitmix_bytes(302, rows * b_bytes).reshape(1, rows, b_bytes).copy()
    A0, B0 = A.copy(), B.copy()
    rk, pv = oracle_solve(GF256, A0[0], B0[0], cols)
    assert rk == rows
    for outer in (0, 1):
        a, b = A.copy(), B.copy()
        rank = np.zeros(1, dtype=np.int32)
        piv = np.full((1, rows), 0xFFFFFFFF, dtype=np.uint32)
        if outer:
            rc = e.emu_solve_outer(ptr(a), a.strides[1], a.strides[0], rows, cols, a.shape[2], ptr(b), b.strides[1],
                                   b.strides[0], b.shape[2], ptr(rank, i32p), ptr(piv, u32p), 1, 1, 128, SMEM, 0)
        else:
            rc = e.emu_solve(GF256, ptr(a), a.strides[1], a.strides[0], rows, cols, a.shape[2], ptr(b), b.strides[1],
                             b.strides[0], b.shape[2], ptr(rank, i32p), ptr(piv, u32p), 1, 1, 64, SMEM, 4, -1, 0)
        assert rc == 0 and rank[0] == rk
        assert np.array_equal(piv[0][:rk], pv[:rk])
        assert np.array_equal(a, A0) and np.array_equal(b, B0)


def test_solve_zero_and_identity():
    e = emu()
    for field in (GF256, GF2):
        rows = cols = 40
        A = random_matrix(field, rows, cols, 1) * 0
        rank = np.zeros(1, dtype=np.int32)
        rc = e.emu_solve(field, ptr(A), A.strides[0], A.nbytes, rows, cols, A.shape[1], None, 0, 0, 0,
                         ptr(rank, i32p), None, 1, 1, 64, SMEM, 0, -1, 0)
        assert rc == 0 and rank[0] == 0 and not A.any()


def test_apply():
    o, e = oracle(), emu()
    for (M, Kc, nbytes, seed, threads, acc) in [(10, 10, 32, 1, 64, 0), (40, 37, 272, 2, 64, 0),
                                               (70, 100, 512, 3, 128, 0), (5, 3, 16, 5, 64, 0),
                                               (40, 37, 272, 6, 64, 1), (100, 50, 48, 7, 32, 0)]:
        cs = (Kc + 15) // 16 * 16 if seed != 5 else 7
        Cm = np.zeros((M, cs), dtype=np.uint8)
        Cm[:, :Kc] = splitmix_bytes(seed, M * Kc).reshape(M, Kc)
        D = splitmix_bytes(seed + 1, Kc * nbytes).reshape(Kc, nbytes).copy()
        Y0 = splitmix_bytes(seed + 2, M * nbytes).reshape(M, nbytes).copy()
        Y1, prev = Y0.copy(), Y0.copy()
        o.orc_gf256_apply(ptr(Cm), cs, M, Kc, ptr(D), nbytes, ptr(Y0), nbytes, nbytes)
        if acc:
            Y0 ^= prev
        for geo in (0, 1, 2):
            Y2 = Y1.copy()
            assert e.emu_apply(ptr(Cm), cs, M, Kc, ptr(D), nbytes, ptr(Y2), nbytes, nbytes, acc, threads, geo) == 0
            assert np.array_equal(Y0, Y2), geo


@pytest.mark.parametrize("args", [
    dict(rows=150, cols=150, b_bytes=48, seed=400),                      # two outer panels
    dict(rows=150, cols=150, b_bytes=48, seed=401, no_fast=1),           # general panel path
    dict(rows=140, cols=270, b_bytes=32, seed=402),                      # wide: three outer panels, rank = rows
    dict(rows=270, cols=140, b_bytes=0, seed=403),                       # tall, no right-hand side
    dict(rows=160, cols=160, b_bytes=64, seed=404, deficient=2),         # free columns inside an outer panel
    dict(rows=130, cols=130, b_bytes=16, seed=405, nsys=3, grid=2),      # persistent loop over systems
    # 128 threads: the candidate group (64) runs the NEXT step's candidate elimination while the other
    # 64 stream the tile pass (look-ahead); 64-thread runs above have no streaming threads left for it
    dict(rows=150, cols=150, b_bytes=48, seed=406, threads=128),
    dict(rows=150, cols=150, b_bytes=48, seed=406, threads=128, no_fast=2),   # same with the look-ahead off
    dict(rows=270, cols=140, b_bytes=0, seed=407, threads=128),          # tall: candidate window never short
    dict(rows=140, cols=270, b_bytes=32, seed=408, threads=128),         # wide: rows run out inside an outer panel
    dict(rows=160, cols=160, b_bytes=64, seed=409, deficient=2, threads=128),  # look-ahead fails (free column) -> general path
    dict(rows=200, cols=200, b_bytes=16, seed=410, nsys=3, grid=2, threads=128),
    # a whole 16-column step without pivots inside an outer panel: the next step's own columns are then
    # window chunk 1 while E still has no columns (split 0) -- the directly written chunk is q, not split
    dict(rows=150, cols=150, b_bytes=32, seed=411, threads=128, zero_cols=(0, 16)),
    dict(rows=150, cols=150, b_bytes=32, seed=412, threads=128, zero_cols=(32, 48)),
])
def test_outer_panel_elimination(args):
    """the algorithm of the tensor-core elimination (outer panels of 128 columns, E tracked as a
    right-hand side, rank-128 update) with panel_kernel on host threads and the tcgen05 product
    replaced by a CPU loop, against the oracle: ranks, pivot columns and every byte of A and B"""
    rows, cols, b_bytes, seed = args["rows"], args["cols"], args["b_bytes"], args["seed"]
    nsys, deficient = args.get("nsys", 1), args.get("deficient", 0)
    e = emu()
    A = np.stack([random_matrix(GF256, rows, cols, seed + 2 * s, rank_deficient=deficient) for s in range(nsys)])
    if "zero_cols" in args:
        A[:, :, args["zero_cols"][0]:args["zero_cols"][1]] = 0
    B = np.stack([splitmix_bytes(seed + 2 * s + 1, rows * b_bytes).reshape(rows, b_bytes)
                  for s in range(nsys)]).copy() if b_bytes else None
    A0, B0 = A.copy(), (B.copy() if B is not None else None)
    want = [oracle_solve(GF256, A0[s], B0[s] if B0 is not None else None, cols) for s in range(nsys)]
    rank = np.zeros(nsys, dtype=np.int32)
    piv = np.full((nsys, rows), 0xFFFFFFFF, dtype=np.uint32)
    rc = e.emu_solve_outer(ptr(A), A.strides[1], A.strides[0], rows, cols, A.shape[2],
                           ptr(B) if B is not None else None, B.strides[1] if B is not None else 0,
                           B.strides[0] if B is not None else 0, B.shape[2] if B is not None else 0,
                           ptr(rank, i32p), ptr(piv, u32p), nsys, args.get("grid", 1), args.get("threads", 64), SMEM,
                           args.get("no_fast", 0))
    assert rc == 0
    assert [int(r) for r in rank] == [w[0] for w in want]
    for s in range(nsys):
        assert np.array_equal(piv[s][:want[s][0]], want[s][1][:want[s][0]])
    assert np.array_equal(A, A0)
    if B is not None:
        assert np.array_equal(B, B0)
    if deficient:
        assert want[0][0] < min(rows, cols)


@pytest.mark.parametrize("args", [
    # (flagged = outer panels the compact pass gave up and the full kernel redid)
    dict(rows=300, cols=256, b_bytes=32, seed=500, cmax=144, flagged=0),                 # two full panels, rows beyond the compact set
    dict(rows=300, cols=256, b_bytes=32, seed=500, cmax=128, flagged=0),                 # smallest compact set
    dict(rows=300, cols=300, b_bytes=32, seed=501, cmax=160, flagged=1),                 # narrow last panel -> full kernel
    dict(rows=256, cols=256, b_bytes=48, seed=502, cmax=192, flagged=0),                 # compact set reaches the end of the system
    dict(rows=320, cols=256, b_bytes=0, seed=503, cmax=144, flagged=0),                  # no right-hand side
    dict(rows=300, cols=256, b_bytes=32, seed=504, cmax=144, deficient=2, flagged=1),    # free column in panel 0 (redone by the full kernel), a zero row in panel 1
    dict(rows=300, cols=384, b_bytes=16, seed=505, cmax=144, flagged=1),                 # wide: the third panel has 44 rows left
    dict(rows=300, cols=256, b_bytes=32, seed=506, cmax=144, sparse=0.25, flagged=None), # many zero entries: pivots swap rows inside the compact set
    dict(rows=300, cols=256, b_bytes=32, seed=507, cmax=144, nsys=3, grid=2, flagged=0), # persistent loop, unit lists of several systems
    dict(rows=300, cols=256, b_bytes=32, seed=508, cmax=144, antidiag=True, flagged=1),  # panel 0: pivots far below the compact set (redone); panel 1: right at its top
    dict(rows=300, cols=256, b_bytes=32, seed=509, cmax=144, swap_far=True, flagged=None),
])
@pytest.mark.parametrize("kb", [4, 6], ids=["compact_pass_4bit_tables", "compact_pass_6bit_tables"])
def test_lazy_outer_panel_elimination(args, kb):
    """the LAZY tensor-core elimination: compact outer-panel factorisation (windows of the first cmax unpivoted
    rows only), full kernel for the panels it gives up, update launch A (compact rows, old pivot rows) and
    launch B (all other rows with their own window as coefficients, NEW pivot rows) walked through the unit
    lists on the CPU -- ranks, pivot columns and every byte of A and B against the oracle"""
    rows, cols, b_bytes, seed, cmax = args["rows"], args["cols"], args["b_bytes"], args["seed"], args["cmax"]
    nsys, deficient = args.get("nsys", 1), args.get("deficient", 0)
    e = emu()
    A = np.stack([random_matrix(GF256, rows, cols, seed + 2 * s, rank_deficient=deficient) for s in range(nsys)])
    if "sparse" in args:
        keep = np.random.default_rng(seed).random(A.shape) < args["sparse"]
        A *= keep.astype(np.uint8)
    if args.get("antidiag"):
        for i in range(rows):
            A[0, i, :max(cols - 1 - i, 0)] = 0
    if args.get("swap_far"):
        # the first 130 rows are zero in column 5: its pivot is row 130 (inside a compact set of 144), which puts an
        # already used row in the middle of the positions the next outer panel takes its compact set from
        A[0, :130, 5] = 0
        A[0, 130, 5] = 7
    B = np.stack([splitmix_bytes(seed + 2 * s + 1, rows * b_bytes).reshape(rows, b_bytes)
                  for s in range(nsys)]).copy() if b_bytes else None
    A0, B0 = A.copy(), (B.copy() if B is not None else None)
    want = [oracle_solve(GF256, A0[s], B0[s] if B0 is not None else None, cols) for s in range(nsys)]
    rank = np.zeros(nsys, dtype=np.int32)
    piv = np.full((nsys, rows), 0xFFFFFFFF, dtype=np.uint32)
    stats = np.zeros(3, dtype=np.uint32)
    e.emu_set_lazy_kb(kb)   # the product launches the compact pass with the 4-bit tables (two CTAs per SM)
    rc = e.emu_solve_outer_lazy(ptr(A), A.strides[1], A.strides[0], rows, cols, A.shape[2],
                                ptr(B) if B is not None else None, B.strides[1] if B is not None else 0,
                                B.strides[0] if B is not None else 0, B.shape[2] if B is not None else 0,
                                ptr(rank, i32p), ptr(piv, u32p), nsys, args.get("grid", 1), args.get("threads", 128), SMEM,
                                args.get("no_fast", 0), cmax, ptr(stats, u32p))
    assert rc == 0
    assert [int(r) for r in rank] == [w[0] for w in want]
    for s in range(nsys):
        assert np.array_equal(piv[s][:want[s][0]], want[s][1][:want[s][0]])
    assert np.array_equal(A, A0)
    if B is not None:
        assert np.array_equal(B, B0)
    if args["flagged"] is not None:
        assert int(stats[0]) == args["flagged"], stats
