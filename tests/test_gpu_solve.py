"""Batched elimination + RHS update on the GPU (oblas_b200_solve_batch and the host
streaming variant) against the oracle / the compiled reference, bit for bit, and at
BASELINE.json's full size through properties that do not need the CPU to redo the
work (A -> I, A_orig * X == B_orig on sampled rows)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import cases
from gpu_helpers import dev, gf256_vecmat, host
from helpers import (GF2, GF4, GF16, GF256, FIELD_BITS, field_row_bytes, mat_view, oracle, oracle_solve, ptr,
                     random_matrix, reference, splitmix_bytes, u32p)
from oblas_b200 import capi, device

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True,
                params=[(-1, 0, -1), (0, 0, -1), (1, 0, -1), (2, 0, -1), (4, 0, -1), (-1, 1, -1), (3, 0, -1), (3, 1, -1),
                        (3, 0, 0), (3, 0, 192), (3, 0, 128), (3, 1, 144)],
                ids=["auto", "geo4x256", "geo6x128", "geo7x64", "geo5x128", "general_panel_path", "tcgen05", "tcgen05_general_panel",
                     "tcgen05_full_panels", "tcgen05_lazy192", "tcgen05_lazy128", "tcgen05_lazy144_general_panel"])
def _bind(cuda_lib, request):
    """every test of this module runs once per Four-Russians table geometry, once with the
    register-resident fast panel factorisation disabled, and (GF(256)) on the tensor-core path
    (outer panels of 128 columns + tcgen05 rank-128 updates) with both panel factorisations, with the full outer
    panels (what batches of this size get by default, and forced) and with the lazy outer panels (compact sets of 192,
    128 and 144 rows; the default for batches of at least 2 x (SM count) systems)"""
    device.bind(0)
    geo, general, lazy = request.param
    cuda_lib.oblas_b200_set_tuning(geo, geo)
    cuda_lib.oblas_b200_set_panel_path(general)
    cuda_lib.oblas_b200_set_lazy_panels(lazy)
    yield
    cuda_lib.oblas_b200_set_tuning(-1, -1)
    cuda_lib.oblas_b200_set_panel_path(0)
    cuda_lib.oblas_b200_set_lazy_panels(-1)


def _batch(field, rows, cols, b_bytes, seed, nsys, deficient_every=0):
    A = np.stack([random_matrix(field, rows, cols, seed + 2 * s,
                                rank_deficient=(1 + s % 2) if deficient_every and s % deficient_every == 0 else 0)
                  for s in range(nsys)])
    B = None
    if b_bytes:
        B = np.stack([splitmix_bytes(seed + 2 * s + 1, rows * b_bytes).reshape(rows, b_bytes) for s in range(nsys)]).copy()
    return A, B


def _check_vs_oracle(field, A, B, cols, rank, piv):
    nsys = A.shape[0]
    A0, B0 = A.copy(), (B.copy() if B is not None else None)
    for s in range(nsys):
        rk, pv = oracle_solve(field, A0[s], B0[s] if B0 is not None else None, cols)
        assert rank[s] == rk, s
        assert np.array_equal(piv[s][:rk].astype(np.uint32), pv[:rk]), s
    return A0, B0


@pytest.mark.parametrize("field,rows,cols,b_bytes,nsys,defe", [
    (GF256, 64, 64, 272, 7, 3),
    (GF256, 100, 100, 1280, 3, 0),
    (GF256, 256, 256, 512, 4, 2),
    (GF256, 33, 50, 32, 2, 0),
    (GF256, 50, 33, 32, 2, 0),
    (GF16, 128, 128, 64, 5, 2),
    (GF16, 300, 300, 0, 2, 0),
    (GF4, 128, 128, 32, 5, 2),
    (GF4, 300, 300, 16, 2, 0),
    (GF2, 128, 128, 32, 5, 0),
    (GF2, 500, 500, 64, 3, 0),
    (GF2, 100, 300, 16, 2, 0),
])
def test_solve_batch_vs_oracle(field, rows, cols, b_bytes, nsys, defe):
    A, B = _batch(field, rows, cols, b_bytes, 3000 + rows + field, nsys, defe)
    dA, dB = dev(A), (dev(B) if B is not None else None)
    rank, piv = device.solve_batch(field, dA, cols, dB, want_pivcols=True)
    rank, piv = host(rank), host(piv)
    A0, B0 = _check_vs_oracle(field, A, B, cols, rank, piv)
    assert np.array_equal(host(dA), A0)
    if B is not None:
        assert np.array_equal(host(dB), B0)


def test_solve_more_systems_than_sms():
    """persistent CTAs loop over the batch: 300 systems on 148 SMs"""
    A, B = _batch(GF256, 48, 48, 64, 4000, 300, 7)
    dA, dB = dev(A), dev(B)
    rank, piv = device.solve_batch(GF256, dA, 48, dB, want_pivcols=True)
    A0, B0 = _check_vs_oracle(GF256, A, B, 48, host(rank), host(piv))
    assert np.array_equal(host(dA), A0) and np.array_equal(host(dB), B0)


def test_zero_and_identity_systems():
    K = 64
    A = np.zeros((2, K, K), dtype=np.uint8)
    A[1] = np.eye(K, dtype=np.uint8) * 7
    B = splitmix_bytes(1, 2 * K * 32).reshape(2, K, 32).copy()
    dA, dB = dev(A), dev(B)
    rank, _ = device.solve_batch(GF256, dA, K, dB)
    assert host(rank).tolist() == [0, K]
    A0, B0 = A.copy(), B.copy()
    for s in range(2):
        oracle_solve(GF256, A0[s], B0[s], K)
    assert np.array_equal(host(dA), A0) and np.array_equal(host(dB), B0)


def test_pivotless_steps_inside_an_outer_panel():
    """whole 16-column steps without pivots (zero columns) at the start and in the middle of an outer
    panel of the tensor-core path: window chunk of the next full step != number of E chunks"""
    K, L = 640, 96
    A = np.stack([random_matrix(GF256, K, K, 7800 + s) for s in range(2)])
    A[:, :, 0:16] = 0
    A[:, :, 160:192] = 0
    A[1, :, 300] = 0
    B = splitmix_bytes(7802, 2 * K * L).reshape(2, K, L).copy()
    dA, dB = dev(A), dev(B)
    rank, piv = device.solve_batch(GF256, dA, K, dB, want_pivcols=True)
    A0, B0 = A.copy(), B.copy()
    for s in range(2):
        rk, pv = oracle_solve(GF256, A0[s], B0[s], K)
        assert int(host(rank)[s]) == rk
        assert np.array_equal(host(piv)[s][:rk].astype(np.uint32), pv[:rk])
    assert np.array_equal(host(dA), A0) and np.array_equal(host(dB), B0)


def test_every_row_moves():
    """row-reversed systems: the pivot of column c is the LAST row with a non-zero entry, so every row
    changes position -- the row permutation at the end stages all rows (several column blocks), the
    other extreme of the usual handful of swaps.  Both elimination paths; system 1 is rank deficient."""
    K, L = 640, 272
    A = np.stack([random_matrix(GF256, K, K, 7700 + s) for s in range(2)])
    for s in range(2):   # zero above the anti-diagonal: column c has its first non-zero entry in row K-1-c
        for i in range(K):
            A[s, i, :K - 1 - i] = 0
        A[s, np.arange(K), K - 1 - np.arange(K)] |= 1
    A[1, :, 5] = 0   # a free column
    B = splitmix_bytes(7702, 2 * K * L).reshape(2, K, L).copy()
    dA, dB = dev(A), dev(B)
    rank, piv = device.solve_batch(GF256, dA, K, dB, want_pivcols=True)
    A0, B0 = A.copy(), B.copy()
    for s in range(2):
        rk, pv = oracle_solve(GF256, A0[s], B0[s], K)
        assert int(host(rank)[s]) == rk
        assert np.array_equal(host(piv)[s][:rk].astype(np.uint32), pv[:rk])
    assert np.array_equal(host(dA), A0) and np.array_equal(host(dB), B0)


def test_config2_shape_vs_reference():
    """K=1024, 1280-byte RHS symbols (BASELINE.json configs[1]) -- whole buffers against
    the unmodified reference's AVX build driven by oracle/ref_driver.c (falls back to a
    K=384 system on the scalar oracle when the host has no AVX2)."""
    lib = reference("avx512") or reference("avx2")
    K, L, nsys = (1024, 1280, 3) if lib is not None else (384, 1280, 2)
    A, B = _batch(GF256, K, K, L, 5000, nsys, 0)
    A[1, K // 2] = A[1, 3]  # one singular system in the batch
    dA, dB = dev(A), dev(B)
    rank, piv = device.solve_batch(GF256, dA, K, dB, want_pivcols=True)
    rank, piv, gA, gB = host(rank), host(piv), host(dA), host(dB)
    for s in range(nsys):
        if lib is not None:
            ma, mb = lib.octmat_new(K, K), lib.octmat_new(K, L)
            assert ma.contents.stride == K and mb.contents.stride == L
            mat_view(ma)[:] = A[s]
            mat_view(mb)[:] = B[s]
            pv = np.zeros(K, dtype=np.uint32)
            rk = lib.refdrv_octmat_solve(ma, mb, ptr(pv, u32p))
            wa, wb = mat_view(ma).copy(), mat_view(mb).copy()
            lib.octmat_free(ma)
            lib.octmat_free(mb)
        else:
            wa, wb = A[s].copy(), B[s].copy()
            rk, pv = oracle_solve(GF256, wa, wb, K)
        assert rank[s] == rk and np.array_equal(piv[s][:rk].astype(np.uint32), pv[:rk])
        assert np.array_equal(gA[s], wa), s
        assert np.array_equal(gB[s], wb), s
    assert rank[1] == K - 1 and rank[0] == K


def test_full_size_roundtrip_properties():
    """A batch at full K=1024 / L=1280 generated on the device: every full-rank system
    ends with A == I, and A_orig * X == B_orig on sampled rows (checked on the CPU)."""
    K, L, nsys = 1024, 1280, 160
    dA = torch.empty((nsys, K, K), dtype=torch.uint8, device="cuda")
    dB = torch.empty((nsys, K, L), dtype=torch.uint8, device="cuda")
    device.fill_random(dA, 6000)
    device.fill_random(dB, 6001)
    A_orig, B_orig = dA.clone(), dB.clone()
    rank, _ = device.solve_batch(GF256, dA, K, dB)
    rank = host(rank)
    eye = torch.eye(K, dtype=torch.uint8, device="cuda")
    full = np.nonzero(rank == K)[0]
    assert len(full) >= nsys - 8  # P(singular) ~ 0.4 %
    for s in full:
        assert torch.equal(dA[s], eye), s
    rng = np.random.default_rng(0)
    for s in rng.choice(full, 6, replace=False):
        X = host(dB[s])
        Ao, Bo = host(A_orig[s]), host(B_orig[s])
        for r in rng.choice(K, 4, replace=False):
            assert np.array_equal(gf256_vecmat(Ao[r], X), Bo[r]), (s, r)


def test_solve_batch_host_streaming(cuda_lib):
    """host buffers in, host buffers out (the e2e entry point), several chunks"""
    K, L, nsys = 96, 272, 37
    A, B = _batch(GF256, K, K, L, 7000, nsys, 5)
    A0, B0 = A.copy(), B.copy()
    want = [oracle_solve(GF256, A0[s], B0[s], K) for s in range(nsys)]
    rank = np.zeros(nsys, dtype=np.int32)
    piv = np.zeros((nsys, K), dtype=np.uint32)
    capi.check(cuda_lib.oblas_b200_solve_batch_host(
        GF256, A.ctypes.data, A.strides[1], A.strides[0], K, K, K, B.ctypes.data, B.strides[1], B.strides[0], L,
        rank.ctypes.data, piv.ctypes.data, nsys, 8))
    assert rank.tolist() == [w[0] for w in want]
    assert np.array_equal(A, A0) and np.array_equal(B, B0)
    for s in range(nsys):
        assert np.array_equal(piv[s][:want[s][0]], want[s][1][:want[s][0]])


def test_rows_beyond_shared_memory_fail_loudly():
    K = 40000  # u16 row map / shared-memory panel cannot hold it; there is no fallback path
    dA = torch.zeros((1, 1, 16), dtype=torch.uint8, device="cuda")
    rank = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = capi.lib().oblas_b200_solve_batch(GF256, C.c_void_p(dA.data_ptr()), 16, 16, K, 16, 16, None, 0, 0, 0,
                                           C.c_void_p(rank.data_ptr()), None, 1)
    assert rc != 0 and b"no fallback" in capi.lib().oblas_b200_last_error()


@pytest.mark.parametrize("rows,cols,b_bytes,nsys,deficient", [
    (640, 640, 272, 3, 0),        # auto-dispatch size (rows >= 512): tensor-core path by default
    (700, 650, 400, 2, 2),        # tall, cols not a multiple of 128, rank deficient
    (520, 900, 0, 2, 0),          # wide, no right-hand side
    (1024, 1024, 1280, 2, 1),     # cfg2 shape, one deficient system
    (640, 640, 16, 2, 0),         # column tiles straddle A|B, last tile 48 / 16 columns wide (narrow MMA)
    (768, 768, 224, 2, 1),        # A|B boundary inside a tile in every outer panel; 3 tiles in the last panels
    (512, 512, 2400, 2, 0),       # B much wider than A: 10+ tiles per unit, exact multiple of the tile width
    (530, 530, 240, 1, 0),        # rows not a multiple of 16 (partial last row tile), single-tile last panel
])
def test_tensor_core_path_equals_table_kernel_at_mid_sizes(rows, cols, b_bytes, nsys, deficient, cuda_lib):
    """the two elimination paths must leave identical bytes, ranks and pivot columns (the table
    kernel itself is pinned against the oracle / the compiled reference above)"""
    a_bytes = (cols + 15) // 16 * 16
    A = torch.empty((nsys, rows, a_bytes), dtype=torch.uint8, device="cuda")
    device.fill_random(A, 900 + rows)
    if a_bytes > cols:
        A[:, :, cols:] = 0
    if deficient:   # make some rows / columns dependent in system 0
        A[0, 5] = A[0, 3] ^ A[0, 4]
        A[0, :, 17] = A[0, :, 2]
        if deficient > 1:
            A[0, :, 300] = 0
    B = None
    if b_bytes:
        B = torch.empty((nsys, rows, b_bytes), dtype=torch.uint8, device="cuda")
        device.fill_random(B, 901 + rows)
    out = {}
    for name, geo in (("table", 1), ("tc", 3)):
        cuda_lib.oblas_b200_set_tuning(geo, -1)
        a, b = A.clone(), (B.clone() if B is not None else None)
        rank, piv = device.solve_batch(GF256, a, cols, b, want_pivcols=True)
        capi.check(cuda_lib.oblas_b200_sync(), "sync")
        out[name] = (a, b, rank, piv)
    cuda_lib.oblas_b200_set_tuning(-1, -1)
    ta, tb, trank, tpiv = out["table"]
    ca, cb, crank, cpiv = out["tc"]
    assert torch.equal(trank, crank)
    for s in range(nsys):
        r = int(trank[s])
        assert torch.equal(tpiv[s, :r], cpiv[s, :r])
    assert torch.equal(ta, ca)
    if B is not None:
        assert torch.equal(tb, cb)
    if deficient:
        assert int(trank[0]) < min(rows, cols)


_pair_checked = False


def test_cta_pair_update_kernel_in_a_subprocess():
    """OBLAS_B200_TC_PAIR=1 switches the rank-128 update to CTA pairs (tcgen05 cta_group::2, multicast
    commits, remote barrier arrives); the switch is read once per process, so a child process runs a
    tensor-core elimination on it and compares every byte with the table kernel"""
    import subprocess
    import sys
    global _pair_checked
    if _pair_checked:
        pytest.skip("once per module is enough (the child process sets its own tuning)")
    _pair_checked = True
    code = r"""
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import torch
from oblas_b200 import capi, device
device.bind(0)
lib = capi.lib()
for rows, cols, L, nsys in ((640, 640, 272, 3), (530, 700, 96, 2)):
    ab = (cols + 15) // 16 * 16
    A = torch.empty((nsys, rows, ab), dtype=torch.uint8, device="cuda"); device.fill_random(A, 11 + rows)
    if ab > cols: A[:, :, cols:] = 0
    A[0, 5] = A[0, 3] ^ A[0, 4]
    B = torch.empty((nsys, rows, L), dtype=torch.uint8, device="cuda"); device.fill_random(B, 12 + rows)
    out = []
    for geo in (1, 3):
        lib.oblas_b200_set_tuning(geo, -1)
        a, b = A.clone(), B.clone()
        r, pv = device.solve_batch(capi.GF256, a, cols, b, want_pivcols=True)
        capi.check(lib.oblas_b200_sync(), "sync")
        out.append((a, b, r))
    assert torch.equal(out[0][2], out[1][2]) and torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
print("pair ok")
""" % (ROOT_DIR, ROOT_DIR)
    env = dict(os.environ, OBLAS_B200_TC_PAIR="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "pair ok" in r.stdout, r.stdout[-2000:]
    # the same child with OBLAS_B200_TC_EVEN=0: the single-CTA update kernel on the FIXED tile geometry (240-column
    # tiles + a rest tile, processed second) instead of the evenly dealt tiles every other test runs on
    env = dict(os.environ, OBLAS_B200_TC_EVEN="0")
    env.pop("OBLAS_B200_TC_PAIR", None)
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "pair ok" in r.stdout, r.stdout[-2000:]
