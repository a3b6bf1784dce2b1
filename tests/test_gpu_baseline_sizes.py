"""GPU parity at BASELINE.json's OWN sizes for configs[2], [3] and [4], whole buffers against the
unmodified reference (oracle/_ref: AVX-512 / AVX2 / scalar build + oracle/ref_driver.c, falling
back to the C restatement), plus the paths only large problems reach:

  * GF(2) K=8192 runs solve_kernel<1, 4, 6, 128, GROWS> (128-bit panels, per-row state in a global workspace:
    the default for systems too tall for shared memory) and solve_kernel<1, NBW=2, 4, 256> (64-bit panels in
    shared memory, oblas_b200_set_tall_mode(0)) -- binmat.c:26-33,44-57 semantics;
  * GF(16)/GF(4) K=2048 -- gfmat.c:83-89,124-156 (GF(4): corrected table);
  * apply with Kc=10000 coefficient columns (octmat.c:32-46, one oaxpy per coefficient);
  * several chunks of systems in solve_tc_run and several column passes in apply_tc_run, forced
    with oblas_b200_set_workspace_budget;
  * the give-up path of the tensor-core kernels (bounded barrier waits) through fault injection.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_helpers import dev, gf256_vecmat, host
from helpers import (GF2, GF4, GF16, GF256, FIELD_BITS, field_row_bytes, mat_view, oracle, oracle_solve, ptr,
                     random_matrix, reference, splitmix_bytes, u32p)
from oblas_b200 import capi, device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _bind(cuda_lib):
    device.bind(0)
    cuda_lib.oblas_b200_set_tuning(-1, -1)
    cuda_lib.oblas_b200_set_panel_path(0)
    yield
    cuda_lib.oblas_b200_set_workspace_budget(0)
    cuda_lib.oblas_b200_set_fault_injection(0)
    cuda_lib.oblas_b200_set_wait_limit_ms(0)


def best_reference():
    for v in ("avx512", "avx2", "ref"):
        lib = reference(v)
        if lib is not None:
            return lib
    return None


def ref_solve(field, A, B, cols):
    """rank, pivcols, A', B' of the caller-side loop run by the compiled reference (oracle port if absent)"""
    lib = best_reference()
    rows = A.shape[0]
    pv = np.full(rows, 0xFFFFFFFF, dtype=np.uint32)
    if lib is None or (field == GF256 and lib.refdrv_octmat_align() != 16 and (A.shape[1] % 64 or (B is not None and B.shape[1] % 64))):
        a, b = A.copy(), (B.copy() if B is not None else None)
        rk, pv = oracle_solve(field, a, b, cols)
        return rk, pv, a, b
    if field == GF2:
        new, free = lib.binmat_new, lib.binmat_free
        ma = new(rows, cols)
        mb = new(rows, B.shape[1] * 8) if B is not None else None
    elif field == GF256:
        new, free = lib.octmat_new, lib.octmat_free
        ma = new(rows, cols)
        mb = new(rows, B.shape[1]) if B is not None else None
    else:
        free = lib.gfmat_free
        ma = lib.gfmat_new(field, rows, cols)
        mb = lib.gfmat_new(field, rows, B.shape[1] * 8 // FIELD_BITS[field]) if B is not None else None
    assert mat_view(ma).shape == A.shape, (mat_view(ma).shape, A.shape)
    mat_view(ma)[:] = A
    if mb is not None:
        assert mat_view(mb).shape == B.shape, (mat_view(mb).shape, B.shape)
        mat_view(mb)[:] = B
    if field == GF2:
        rk = lib.refdrv_binmat_rref(ma, mb, ptr(pv, u32p))
    elif field == GF256:
        rk = lib.refdrv_octmat_solve(ma, mb, ptr(pv, u32p))
    else:
        rk = lib.refdrv_gfmat_solve(ma, mb, ptr(pv, u32p), 1)
    a, b = mat_view(ma).copy(), (mat_view(mb).copy() if mb is not None else None)
    free(ma)
    if mb is not None:
        free(mb)
    return rk, pv, a, b


def check_batch(field, A, B, cols):
    dA, dB = dev(A), (dev(B) if B is not None else None)
    rank, piv = device.solve_batch(field, dA, cols, dB, want_pivcols=True)
    device.sync()
    rank, piv, gA = host(rank), host(piv), host(dA)
    gB = host(dB) if B is not None else None
    ranks = []
    for s in range(A.shape[0]):
        rk, pv, wa, wb = ref_solve(field, A[s], B[s] if B is not None else None, cols)
        assert rank[s] == rk, (s, rank[s], rk)
        assert np.array_equal(piv[s][:rk].astype(np.uint32), pv[:rk]), s
        assert np.array_equal(gA[s], wa), s
        if B is not None:
            assert np.array_equal(gB[s], wb), s
        ranks.append(rk)
    return ranks


@pytest.fixture(params=[1, 0], ids=["rows_in_global_128bit_panels", "64bit_panels_in_shared_memory"])
def tall_mode(request):
    """the two table-kernel variants a system too tall for 128-bit panels in shared memory can take:
    solve_kernel<.., NBW=4, 6, 128, GROWS> (per-row state in a global workspace, the default) and
    solve_kernel<.., NBW=2, ..> (64-bit panels, everything in shared memory)"""
    capi.lib().oblas_b200_set_tall_mode(request.param)
    yield request.param
    capi.lib().oblas_b200_set_tall_mode(1)


def test_config3_gf2_k8192_whole_buffers_vs_reference(tall_mode):
    """BASELINE configs[2]: GF(2) binmat elimination K=8192, bit-packed (1024-byte rows).  Three systems:
    a random one (rank K-d, d = 0..2 with probability 0.29/0.58/0.13), one with a forced row
    dependency (rank <= K-1) and one with a zero column and a duplicated row.  Ranks, pivot columns
    and every byte of A against refdrv_binmat_rref (binmat_get / binmat_add / binmat_swaprow)."""
    K = 8192
    A = np.stack([random_matrix(GF2, K, K, 8100 + s) for s in range(3)])
    assert A.shape[2] == 1024
    A[1, 4000] = A[1, 17] ^ A[1, 4999]
    A[2, :, 3] &= 0xFE          # zero column 24
    A[2, 6000] = A[2, 1]
    ranks = check_batch(GF2, A, None, K)
    assert ranks[1] < K and ranks[2] < K - 1 and min(ranks) >= K - 8
    plan, chunk, geo = C.c_int(), C.c_size_t(), (C.c_int * 3)()
    capi.check(capi.lib().oblas_b200_solve_plan(GF2, K, K, 1024, 0, 3, C.byref(plan), C.byref(chunk), geo))
    assert plan.value == 0 and list(geo) == ([4, 6, 128] if tall_mode else [2, 4, 256])


def test_config3_gf2_k8192_with_rhs(tall_mode):
    """same size with a 32-byte right-hand side (binmat_add on B beside A)"""
    K = 8192
    A = random_matrix(GF2, K, K, 8200)[None]
    B = splitmix_bytes(8201, K * 32).reshape(1, K, 32).copy()
    check_batch(GF2, A, B, K)


@pytest.mark.parametrize("field", [GF16, GF4], ids=["gf16", "gf4"])
def test_config4_packed_fields_k2048_whole_buffers_vs_reference(field):
    """BASELINE configs[3]: GF(16) / GF(4) elimination K=2048 (1024- / 512-byte rows), three systems, one
    of them rank deficient, one with a right-hand side check in the second call; against
    refdrv_gfmat_solve (gfmat_get/swaprow/scal/axpy; GF(4) with the corrected table)."""
    K = 2048
    A = np.stack([random_matrix(field, K, K, 8300 + 10 * field + s) for s in range(3)])
    assert A.shape[2] == K * FIELD_BITS[field] // 8
    A[1, 1500] = A[1, 2]                 # duplicate row
    A[1, :, 0] &= 0xF0 if field == GF16 else 0xFC   # zero column 0
    ranks = check_batch(field, A, None, K)
    assert ranks[1] < K
    B = splitmix_bytes(8400 + field, K * 256).reshape(1, K, 256).copy()   # 256 bytes: a legal stride for OCTMAT_ALIGN 16/32/64
    check_batch(field, A[:1].copy(), B, K)


def ref_apply(Cm, Kc, D, rows):
    """rows `rows` of Y = C*D by the reference's oaxpy loop (refdrv_octmat_apply), oracle port if absent"""
    lib = best_reference()
    n = D.shape[1]
    Cs = np.ascontiguousarray(Cm[rows, :Kc])
    if lib is None or (lib.refdrv_octmat_align() != 16 and n % 64):
        Y = np.zeros((len(rows), n), dtype=np.uint8)
        oracle().orc_gf256_apply(ptr(Cs), Cs.strides[0], len(rows), Kc, ptr(D), D.strides[0], ptr(Y), n, n)
        return Y
    c_, d_, y_ = lib.octmat_new(len(rows), Kc), lib.octmat_new(Kc, n), lib.octmat_new(len(rows), n)
    mat_view(c_)[:, :Kc] = Cs
    mat_view(d_)[:, :n] = D
    lib.refdrv_octmat_apply(c_, d_, y_)
    Y = mat_view(y_)[:, :n].copy()
    for m in (c_, d_, y_):
        lib.octmat_free(m)
    return Y


def test_config5_apply_k10000_column_window_vs_reference(cuda_lib):
    """BASELINE configs[4] at its own inner dimension: C 10000 x 10000, a 960-byte symbol-column window
    of D (two full 480-byte tiles of the tcgen05 kernel; 1250 K-steps per tile).  Every 16th row plus
    the last 16 rows of Y against the reference's K^2 oaxpy loop, all rows by linearity (D1 ^ D2), and
    the same result again in three column passes (workspace budget below one tile's operand blocks)."""
    K, N = 10000, 960 + 64
    cs = 10048   # multiple of 64: the same C serves the AVX-512 build of the reference
    Cm = torch.zeros((K, cs), dtype=torch.uint8, device="cuda")
    tmp = torch.empty((K, cs), dtype=torch.uint8, device="cuda")
    device.fill_random(tmp, 9100)
    Cm[:, :K] = tmp[:, :K]
    del tmp
    D1 = torch.empty((K, N), dtype=torch.uint8, device="cuda")
    D2 = torch.empty((K, N), dtype=torch.uint8, device="cuda")
    device.fill_random(D1, 9101)
    device.fill_random(D2, 9102)
    Y1 = device.apply(Cm, D1)
    Y2 = device.apply(Cm, D2)
    Y12 = device.apply(Cm, D1 ^ D2)
    device.sync()
    assert torch.equal(Y1 ^ Y2, Y12)
    rows = np.concatenate([np.arange(0, K, 16), np.arange(K - 16, K)])
    want = ref_apply(host(Cm), K, host(D1), rows)
    assert np.array_equal(host(Y1)[rows], want)
    # several passes over the column range: one 480-byte tile of 1250 K-steps needs 19.2 MB of fp4 blocks
    old = cuda_lib.oblas_b200_set_workspace_budget(20 << 20)
    try:
        launches0 = cuda_lib.oblas_b200_launch_count()
        Y1p = device.apply(Cm, D1)
        device.sync()
        assert cuda_lib.oblas_b200_launch_count() - launches0 == 6   # 3 passes x (expansion + GEMM)
        assert torch.equal(Y1p, Y1)
        Y1p ^= Y1
        Yacc = device.apply(Cm, D1, Y=Y1p, accumulate=True)       # 0 ^= C*D in three passes
        device.sync()
        assert torch.equal(Yacc, Y1)
    finally:
        cuda_lib.oblas_b200_set_workspace_budget(old)


def test_elimination_in_several_chunks_of_systems(cuda_lib):
    """solve_tc_run with a workspace budget that holds two systems: 7 systems = 4 chunks (the bench's
    4096 systems run as 2 chunks with the default 3 GiB).  Whole buffers against the table kernel and,
    for two systems of each chunk, against the reference."""
    K, L, nsys = 640, 320, 7   # strides that are multiples of 64: the AVX-512 build of the reference (OCTMAT_ALIGN=64) takes them
    A = np.stack([random_matrix(GF256, K, K, 9200 + s) for s in range(nsys)])
    A[3, 77] = A[3, 5]
    A[6, :, 130] = 0
    B = np.stack([splitmix_bytes(9300 + s, K * L).reshape(K, L) for s in range(nsys)]).copy()
    per_sys = 4 * 16 * 7680 + K * 128 + K * 2 + 272
    per_sys += 2 * 192 * 128 + 192 * 2 + 16 + (192 // 16 + K // 16) * 4    # lazy outer panels, compact sets of 192 rows
    cuda_lib.oblas_b200_set_lazy_panels(192)   # (the lazy flow has the bigger workspace and more launch classes)
    old = cuda_lib.oblas_b200_set_workspace_budget(2 * per_sys + 4096)
    try:
        cuda_lib.oblas_b200_set_tuning(3, -1)
        cuda_lib.oblas_b200_solve_tc_times(1, None, None)
        dA, dB = dev(A), dev(B)
        rank, piv = device.solve_batch(GF256, dA, K, dB, want_pivcols=True)
        device.sync()
        ms, n = (C.c_double * 5)(), (C.c_uint64 * 5)()
        cuda_lib.oblas_b200_solve_tc_times_ex(0, ms, n, 5)
        # 4 chunks x (5 outer panels: compact pass + the full kernel's pass over what it gave up, + 1 permutation)
        assert n[3] == 4 and n[0] == 4 * 5 and n[4] == 4 * 5, list(n)
    finally:
        cuda_lib.oblas_b200_set_workspace_budget(old)
        cuda_lib.oblas_b200_set_tuning(-1, -1)
        cuda_lib.oblas_b200_set_lazy_panels(-1)
    # the same batch with full outer panels, also in chunks of two systems
    per_sys0 = 4 * 16 * 7680 + K * 128 + K * 2 + 272
    old = cuda_lib.oblas_b200_set_workspace_budget(2 * per_sys0 + 4096)
    try:
        cuda_lib.oblas_b200_set_lazy_panels(0)
        cuda_lib.oblas_b200_set_tuning(3, -1)
        cuda_lib.oblas_b200_solve_tc_times(1, None, None)
        fA, fB = dev(A), dev(B)
        frank, fpiv = device.solve_batch(GF256, fA, K, fB, want_pivcols=True)
        device.sync()
        ms, n = (C.c_double * 5)(), (C.c_uint64 * 5)()
        cuda_lib.oblas_b200_solve_tc_times_ex(0, ms, n, 5)
        assert n[3] == 4 and n[0] == 4 * 5 and n[4] == 0, list(n)
        assert torch.equal(rank, frank) and torch.equal(piv, fpiv) and torch.equal(dA, fA) and torch.equal(dB, fB)
    finally:
        cuda_lib.oblas_b200_set_workspace_budget(old)
        cuda_lib.oblas_b200_set_tuning(-1, -1)
        cuda_lib.oblas_b200_set_lazy_panels(-1)
    cuda_lib.oblas_b200_set_tuning(1, -1)
    tA, tB = dev(A), dev(B)
    trank, tpiv = device.solve_batch(GF256, tA, K, tB, want_pivcols=True)
    device.sync()
    cuda_lib.oblas_b200_set_tuning(-1, -1)
    assert torch.equal(rank, trank) and torch.equal(dA, tA) and torch.equal(dB, tB)
    gA, gB, rank = host(dA), host(dB), host(rank)
    for s in (0, 3, 4, 6):
        rk, pv, wa, wb = ref_solve(GF256, A[s], B[s], K)
        assert rank[s] == rk and np.array_equal(gA[s], wa) and np.array_equal(gB[s], wb), s
    assert rank[3] == K - 1 and rank[6] == K - 1


def test_host_streaming_on_the_tensor_core_path(cuda_lib):
    """oblas_b200_solve_batch_host with systems big enough for the tcgen05 path: 3 lanes, each with its
    own stream, workspace slot and status word; 7 systems in chunks of 2"""
    K, L, nsys = 640, 320, 7   # strides that are multiples of 64: the AVX-512 build of the reference (OCTMAT_ALIGN=64) takes them
    A = np.stack([random_matrix(GF256, K, K, 9400 + s) for s in range(nsys)])
    A[2, 100] = A[2, 101]
    B = np.stack([splitmix_bytes(9500 + s, K * L).reshape(K, L) for s in range(nsys)]).copy()
    A0, B0 = A.copy(), B.copy()
    rank = np.zeros(nsys, dtype=np.int32)
    piv = np.zeros((nsys, K), dtype=np.uint32)
    capi.check(cuda_lib.oblas_b200_solve_batch_host(
        GF256, A.ctypes.data, A.strides[1], A.strides[0], K, K, K, B.ctypes.data, B.strides[1], B.strides[0], L,
        rank.ctypes.data, piv.ctypes.data, nsys, 2))
    want = [ref_solve(GF256, A0[s], B0[s], K) for s in range(nsys)]
    for s in range(nsys):
        rk, pv, wa, wb = want[s]
        assert rank[s] == rk and np.array_equal(piv[s][:rk], pv[:rk]), s
        assert np.array_equal(A[s], wa) and np.array_equal(B[s], wb), s
    assert rank[2] == K - 1
    # OBLAS_B200_SOLVE_SKIP_IDENTITY_A: A of the full-rank systems (= I, known in advance) is not copied back
    A2, B2 = A0.copy(), B0.copy()
    rank2 = np.zeros(nsys, dtype=np.int32)
    capi.check(cuda_lib.oblas_b200_solve_batch_host_ex(
        GF256, A2.ctypes.data, A2.strides[1], A2.strides[0], K, K, K, B2.ctypes.data, B2.strides[1], B2.strides[0], L,
        rank2.ctypes.data, None, nsys, 2, 1))
    assert np.array_equal(rank2, rank) and np.array_equal(B2, B)
    for s in range(nsys):
        assert np.array_equal(A2[s], A0[s] if rank[s] == K else want[s][2]), s
    # the same from PINNED host memory: a kernel writes the A of the rank-deficient systems straight into the
    # host buffer (no host-side wait for the ranks); two deficient systems in different chunks
    pa = cuda_lib.oblas_b200_host_alloc(A0.nbytes)
    pb = cuda_lib.oblas_b200_host_alloc(B0.nbytes)
    try:
        A3 = np.frombuffer((C.c_uint8 * A0.nbytes).from_address(pa), dtype=np.uint8).reshape(A0.shape)
        B3 = np.frombuffer((C.c_uint8 * B0.nbytes).from_address(pb), dtype=np.uint8).reshape(B0.shape)
        A3[:], B3[:] = A0, B0
        A3[5, 7] = A3[5, 8]
        w5 = ref_solve(GF256, A3[5].copy(), B3[5].copy(), K)
        in5 = A3[5].copy()
        rank3 = np.zeros(nsys, dtype=np.int32)
        capi.check(cuda_lib.oblas_b200_solve_batch_host_ex(
            GF256, A3.ctypes.data, A3.strides[1], A3.strides[0], K, K, K, B3.ctypes.data, B3.strides[1], B3.strides[0], L,
            rank3.ctypes.data, None, nsys, 2, 1))
        assert rank3[5] == w5[0] == K - 1 and rank3[2] == K - 1
        for s in range(nsys):
            if s == 5:
                assert np.array_equal(A3[5], w5[2]) and np.array_equal(B3[5], w5[3])
            else:
                assert rank3[s] == rank[s] and np.array_equal(B3[s], B[s]), s
                assert np.array_equal(A3[s], A0[s] if rank[s] == K else want[s][2]), s
        assert not np.array_equal(A3[5], in5)
    finally:
        cuda_lib.oblas_b200_host_free(pa)
        cuda_lib.oblas_b200_host_free(pb)


def test_barrier_time_limit_is_reported_and_the_device_recovers(cuda_lib):
    """fault injection: the operand producer of the tcgen05 kernels never delivers, so the MMA warp's
    wait runs into the (shortened) WALL-time limit, every role drains, tensor memory is released and
    the status word turns oblas_b200_sync into an error -- for the apply, the elimination's update
    kernel and the host-streaming call.  Afterwards the same calls give the right answer again (a
    kernel that had kept its 512 TMEM columns would block the next allocation on that SM)."""
    M, Kc, N = 320, 96, 960
    Cm = dev(splitmix_bytes(9600, M * Kc).reshape(M, Kc))
    D = dev(splitmix_bytes(9601, Kc * N).reshape(Kc, N))
    cuda_lib.oblas_b200_set_tuning(3, 3)
    good = device.apply(Cm, D)
    device.sync()
    K, L = 256, 320
    A = np.stack([random_matrix(GF256, K, K, 9610 + s) for s in range(2)])
    B = splitmix_bytes(9612, 2 * K * L).reshape(2, K, L).copy()
    cuda_lib.oblas_b200_set_wait_limit_ms(50)
    cuda_lib.oblas_b200_set_fault_injection(8)
    try:
        device.apply(Cm, D)
        assert cuda_lib.oblas_b200_sync() != 0
        assert b"time limit" in cuda_lib.oblas_b200_last_error()
        assert cuda_lib.oblas_b200_sync() == 0          # collected: the flag is not sticky
        dA, dB = dev(A), dev(B)
        device.solve_batch(GF256, dA, K, dB)
        assert cuda_lib.oblas_b200_sync() != 0
        # uncollected failure of an asynchronous call: the next call on the stream reports it once
        device.apply(Cm, D)
        torch.cuda.synchronize()
        rc = cuda_lib.oblas_b200_gf256_apply(C.c_void_p(Cm.data_ptr()), Kc, M, Kc, C.c_void_p(D.data_ptr()), N,
                                             C.c_void_p(good.data_ptr()), N, N, 1)
        assert rc != 0 and b"never" in cuda_lib.oblas_b200_last_error()
        hA, hB = A.copy(), B.copy()
        rank = np.zeros(2, dtype=np.int32)
        rc = cuda_lib.oblas_b200_solve_batch_host(GF256, hA.ctypes.data, K, K * K, K, K, K, hB.ctypes.data, L, K * L, L,
                                                  rank.ctypes.data, None, 2, 1)
        assert rc != 0 and b"time limit" in cuda_lib.oblas_b200_last_error()
    finally:
        cuda_lib.oblas_b200_set_fault_injection(0)
        cuda_lib.oblas_b200_set_wait_limit_ms(0)
    again = device.apply(Cm, D)
    device.sync()
    assert torch.equal(again, good)
    dA, dB = dev(A), dev(B)
    rank, _ = device.solve_batch(GF256, dA, K, dB)
    device.sync()
    cuda_lib.oblas_b200_set_tuning(-1, -1)
    for s in range(2):
        rk, pv, wa, wb = ref_solve(GF256, A[s], B[s], K)
        assert int(rank[s]) == rk and np.array_equal(host(dA)[s], wa) and np.array_equal(host(dB)[s], wb)


def test_gf2_scalars_above_one_are_no_ops_in_batches_too(cuda_lib):
    """gfmat_axpy(GF2_1, u > 1) is a no-op in the single call; a batch must behave like n single calls"""
    rows, st = 8, 64
    a = splitmix_bytes(9700, rows * st).reshape(rows, st).copy()
    b = splitmix_bytes(9701, rows * st).reshape(rows, st).copy()
    da, db = dev(a), dev(b)
    idx = torch.arange(rows, dtype=torch.int32, device="cuda")
    u = dev(np.array([0, 1, 2, 3, 1, 255, 0, 1], dtype=np.uint8))
    device.rowops(GF2, capi.OP_AXPY, da, db, idx, idx, u)
    want = a.copy()
    for k in (1, 4, 7):
        want[k] ^= b[k]
    assert np.array_equal(host(da), want)


def test_multi_device_entry_points(cuda_lib):
    """one process, one host thread + context per visible GPU (oblas_b200_set_device underneath):
    oblas_b200_solve_batch_host_multi shards a batch by systems, oblas_b200_gf256_apply_sharded shards
    D / Y by 256-byte aligned symbol-column slices and sends C to the other devices by peer copies.
    On a one-GPU box both run with a single shard (same code path, no peer copy)."""
    ndev = cuda_lib.oblas_b200_device_count()
    assert ndev >= 1
    K, L, nsys = 96, 272, 11
    A = np.stack([random_matrix(GF256, K, K, 9800 + s) for s in range(nsys)])
    A[4, 50] = A[4, 7]
    B = np.stack([splitmix_bytes(9900 + s, K * L).reshape(K, L) for s in range(nsys)]).copy()
    A0, B0 = A.copy(), B.copy()
    want = [oracle_solve(GF256, A0[s], B0[s], K) for s in range(nsys)]
    rank = np.zeros(nsys, dtype=np.int32)
    piv = np.zeros((nsys, K), dtype=np.uint32)
    capi.check(cuda_lib.oblas_b200_solve_batch_host_multi(
        GF256, A.ctypes.data, A.strides[1], A.strides[0], K, K, K, B.ctypes.data, B.strides[1], B.strides[0], L,
        rank.ctypes.data, piv.ctypes.data, nsys, None, 0, 0))
    assert rank.tolist() == [w[0] for w in want]
    assert np.array_equal(A, A0) and np.array_equal(B, B0)
    # sharded apply against the single-device kernel (itself pinned to the oracle in test_gpu_apply.py)
    M, Kc, N = 300, 200, 256 * 5 + 48
    Cm = np.zeros((M, 208), dtype=np.uint8)
    Cm[:, :Kc] = splitmix_bytes(9950, M * Kc).reshape(M, Kc)
    D = splitmix_bytes(9951, Kc * N).reshape(Kc, N).copy()
    Y = np.full((M, N), 0xEE, dtype=np.uint8)
    secs = (C.c_double * 3)()
    capi.check(cuda_lib.oblas_b200_gf256_apply_sharded(Cm.ctypes.data, Cm.strides[0], M, Kc, D.ctypes.data, D.strides[0],
                                                       Y.ctypes.data, Y.strides[0], N, None, 0, secs))
    device.bind(0)
    ref = device.apply(dev(Cm), dev(D))
    device.sync()
    assert np.array_equal(Y, host(ref))
    Y0 = np.zeros((M, N), dtype=np.uint8)
    oracle().orc_gf256_apply(ptr(Cm), Cm.strides[0], M, Kc, ptr(D), N, ptr(Y0), N, N)
    assert np.array_equal(Y, Y0)
    if ndev >= 2:   # an explicit device list in another order
        devs = (C.c_int * 2)(1, 0)
        Y2 = np.zeros_like(Y)
        capi.check(cuda_lib.oblas_b200_gf256_apply_sharded(Cm.ctypes.data, Cm.strides[0], M, Kc, D.ctypes.data, D.strides[0],
                                                           Y2.ctypes.data, Y2.strides[0], N, devs, 2, None))
        assert np.array_equal(Y2, Y0)
    assert cuda_lib.oblas_b200_set_device(0) == 0


def test_tall_systems_on_the_tensor_core_path_vs_reference(cuda_lib):
    """systems with more rows than the shared-memory row state of the outer-panel kernel holds (~2700):
    panel_kernel<.., GROWS> keeps the window chunk, the coefficient vectors, the row map and the candidate
    flags in a global workspace.  K = 3008, two systems (one rank deficient), every byte against the
    reference's loop (octmat.c:24-63)."""
    K, L = 3008, 320
    A = np.stack([random_matrix(GF256, K, K, 9960 + s) for s in range(2)])
    A[1, 2000] = A[1, 3]
    A[1, :, 1500] = 0
    B = np.stack([splitmix_bytes(9970 + s, K * L).reshape(K, L) for s in range(2)]).copy()
    path = C.c_int(0)
    capi.check(cuda_lib.oblas_b200_solve_plan(GF256, K, K, K, L, 2, C.byref(path), None, None))
    assert path.value == 1
    ranks = check_batch(GF256, A, B, K)
    assert ranks[0] == K and ranks[1] == K - 1   # a duplicated row and a zero column: one dependency each, rank K - 1


def test_raptorq_scale_solve_k10000_roundtrip():
    """one K = 10000 GF(256) system with 1280-byte symbols (the apply of BASELINE configs[4] run
    backwards): rank K, A -> I, and A_orig * X == B_orig on sampled rows recomputed on the CPU."""
    K, L = 10000, 1280
    a_bytes = 10000
    dA = torch.empty((1, K, a_bytes), dtype=torch.uint8, device="cuda")
    dB = torch.empty((1, K, L), dtype=torch.uint8, device="cuda")
    device.fill_random(dA, 9980)
    device.fill_random(dB, 9981)
    A0, B0 = dA.clone(), dB.clone()
    rank, _ = device.solve_batch(GF256, dA, K, dB)
    device.sync()
    assert int(rank[0]) == K      # P(singular) = 0.4 %; this seed is regular
    assert torch.equal(dA[0], torch.eye(K, dtype=torch.uint8, device="cuda"))
    X, Ao, Bo = host(dB[0]), host(A0[0]), host(B0[0])
    rng = np.random.default_rng(1)
    for r in rng.choice(K, 5, replace=False):
        assert np.array_equal(gf256_vecmat(Ao[r], X), Bo[r]), r


def test_lazy_outer_panels_counters_and_equality(cuda_lib):
    """lazy outer panels (compact sets of 192 / 128 rows) against the full outer-panel kernel on the same batch:
    identical bytes, ranks and pivot columns; the counters say which (system, outer panel) pairs the compact
    pass factorised and which it gave up -- the rank-deficient panels and the one whose pivots lie far down."""
    K, L, nsys = 512, 272, 40
    A = np.stack([random_matrix(GF256, K, K, 9600 + s) for s in range(nsys)])
    A[5, :, 200] = 0              # free column in outer panel 1
    A[9, 300] = A[9, 2]           # duplicate row: a zero row from outer panel 0 on (no give-up: it is never needed)
    A[11, :400, 130] = 0          # column 130: first non-zero entry in row >= 400, far below the compact set of panel 1
    A[11, 400, 130] = 9
    B = np.stack([splitmix_bytes(9700 + s, K * L).reshape(K, L) for s in range(nsys)]).copy()
    out = {}
    try:
        cuda_lib.oblas_b200_set_tuning(3, -1)
        for lazy in (0, 192, 128):
            cuda_lib.oblas_b200_set_lazy_panels(lazy)
            st = (C.c_uint64 * 2)()
            cuda_lib.oblas_b200_lazy_stats(st, 1)
            dA, dB = dev(A), dev(B)
            rank, piv = device.solve_batch(GF256, dA, K, dB, want_pivcols=True)
            device.sync()
            cuda_lib.oblas_b200_lazy_stats(st, 1)
            out[lazy] = (host(dA), host(dB), host(rank), host(piv), (int(st[0]), int(st[1])))
    finally:
        cuda_lib.oblas_b200_set_tuning(-1, -1)
        cuda_lib.oblas_b200_set_lazy_panels(-1)
    assert out[0][4] == (0, 0)
    for lazy in (192, 128):
        for k in range(4):
            assert np.array_equal(out[0][k], out[lazy][k]), (lazy, k)
        compact, given_up = out[lazy][4]
        assert compact + given_up == nsys * 4
        # system 5: panel 1 (free column) and every later panel (127 + 128 k pivots: a panel of 128 columns with only ...
        # rows left is still fine) -- at least the one panel; system 11: panel 1
        assert 2 <= given_up <= 8, out[lazy][4]
    A0, B0 = A.copy(), B.copy()
    for s in (0, 5, 9, 11):
        rk, pv = oracle_solve(GF256, A0[s], B0[s], K)
        assert int(out[192][2][s]) == rk
        assert np.array_equal(out[192][0][s], A0[s]) and np.array_equal(out[192][1][s], B0[s])
        assert np.array_equal(out[192][3][s][:rk].astype(np.uint32), pv[:rk])



@pytest.mark.parametrize("field,rows,cols,bbytes", [(GF16, 4000, 512, 256), (GF4, 6000, 1024, 0), (GF2, 10000, 2048, 32),
                                                    (GF256, 3000, 320, 64)],
                         ids=["gf16_4000x512", "gf4_6000x1024", "gf2_10000x2048", "gf256_3000x320_table_kernel"])
def test_tall_systems_on_the_table_kernel(cuda_lib, field, rows, cols, bbytes):
    """rows whose state does not fit shared memory beside the 128-bit-panel tables: the GROWS variant (per-row
    state in global memory; GF(2) beyond the ~9800 rows the 64-bit panels could take) against the reference loop,
    and against the 64-bit-panel variant where that one fits; two systems, one rank deficient"""
    A = np.stack([random_matrix(field, rows, cols, 9900 + 7 * field + s) for s in range(2)])
    A[1, rows - 3] = A[1, 5]
    A[1, :, 1] = 0
    B = np.stack([splitmix_bytes(9950 + s, rows * bbytes).reshape(rows, bbytes) for s in range(2)]).copy() if bbytes else None
    try:
        if field == GF256:
            cuda_lib.oblas_b200_set_tuning(1, -1)   # the table kernel (the tensor-core path takes GF(256) by default)
        cuda_lib.oblas_b200_set_tall_mode(1)
        plan, chunk, geo = C.c_int(), C.c_size_t(), (C.c_int * 3)()
        capi.check(cuda_lib.oblas_b200_solve_plan(field, rows, cols, A.shape[2], bbytes, 2, C.byref(plan), C.byref(chunk), geo))
        assert plan.value == 0 and list(geo) == [4, 6, 128]
        dA, dB = dev(A), (dev(B) if B is not None else None)
        rank, piv = device.solve_batch(field, dA, cols, dB, want_pivcols=True)
        device.sync()
        for s in range(2):
            rk, pv, wa, wb = ref_solve(field, A[s], B[s] if B is not None else None, cols)
            assert int(rank[s]) == rk, s
            assert np.array_equal(host(piv)[s][:rk].astype(np.uint32), pv[:rk]), s
            assert np.array_equal(host(dA)[s], wa) and (B is None or np.array_equal(host(dB)[s], wb)), s
        assert int(rank[1]) < min(rows, cols)
        if rows <= 9000:
            cuda_lib.oblas_b200_set_tall_mode(0)
            tA, tB = dev(A), (dev(B) if B is not None else None)
            trank, tpiv = device.solve_batch(field, tA, cols, tB, want_pivcols=True)
            device.sync()
            assert torch.equal(rank, trank) and torch.equal(piv, tpiv) and torch.equal(dA, tA)
            assert B is None or torch.equal(dB, tB)
    finally:
        cuda_lib.oblas_b200_set_tall_mode(1)
        cuda_lib.oblas_b200_set_tuning(-1, -1)


def test_default_flow_of_a_big_batch_is_the_lazy_one(cuda_lib):
    """default settings: a GF(256) batch of at least 2 x (SM count) systems takes the lazy outer panels (compact sets
    of 144 rows, two compact CTAs per SM); the same batch with the full outer-panel kernel must give the same bytes,
    ranks and pivot columns, and a few systems (two of them rank deficient) must equal the oracle's result"""
    K, L = 512, 64
    nsys = 2 * cuda_lib.oblas_b200_sm_count() + 3
    dA0 = torch.empty((nsys, K, K), dtype=torch.uint8, device="cuda")
    dB0 = torch.empty((nsys, K, L), dtype=torch.uint8, device="cuda")
    device.fill_random(dA0, 9850)
    device.fill_random(dB0, 9851)
    dA0[7, 100] = dA0[7, 3]
    dA0[nsys - 2, :, 300] = 0
    out = {}
    try:
        for lazy in (-1, 0):
            cuda_lib.oblas_b200_set_lazy_panels(lazy)
            st = (C.c_uint64 * 2)()
            cuda_lib.oblas_b200_lazy_stats(st, 1)
            dA, dB = dA0.clone(), dB0.clone()
            rank, piv = device.solve_batch(GF256, dA, K, dB, want_pivcols=True)
            device.sync()
            cuda_lib.oblas_b200_lazy_stats(st, 1)
            out[lazy] = (dA, dB, rank, piv, (int(st[0]), int(st[1])))
    finally:
        cuda_lib.oblas_b200_set_lazy_panels(-1)
    assert out[0][4] == (0, 0)
    compact, given_up = out[-1][4]
    assert compact + given_up == nsys * 4 and compact >= (nsys - 2) * 4 and given_up >= 1
    for k in range(4):
        assert torch.equal(out[-1][k], out[0][k]), k
    rank = host(out[-1][2])
    assert rank[7] == K - 1 and rank[nsys - 2] == K - 1
    A0, B0 = host(dA0), host(dB0)
    gA, gB, gp = host(out[-1][0]), host(out[-1][1]), host(out[-1][3])
    for s in (0, 7, nsys // 2, nsys - 2, nsys - 1):
        a, b = A0[s].copy(), B0[s].copy()
        rk, pv = oracle_solve(GF256, a, b, K)
        assert rank[s] == rk and np.array_equal(gA[s], a) and np.array_equal(gB[s], b), s
        assert np.array_equal(gp[s][:rk].astype(np.uint32), pv[:rk]), s
