"""Dense GF(256) apply Y = C*D and the row-degree kernel on the GPU."""
import numpy as np
import pytest
import torch

from gpu_helpers import dev, gf256_vecmat, host
from helpers import GF2, GF4, GF16, GF256, FIELD_BITS, oracle, ptr, splitmix_bytes
from oblas_b200 import capi, device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True, params=[-1, 0, 1, 2, 3], ids=["auto", "geo4x256", "geo6x128", "geo7x64", "tcgen05"])
def _bind(cuda_lib, request):
    """every test of this module runs once per Four-Russians table geometry"""
    device.bind(0)
    cuda_lib.oblas_b200_set_tuning(request.param, request.param)
    yield
    cuda_lib.oblas_b200_set_tuning(-1, -1)


@pytest.mark.parametrize("M,Kc,nbytes,cs", [(10, 10, 32, 16), (40, 37, 272, 48), (300, 100, 512, 112),
                                            (33, 16, 256, 16), (5, 3, 16, 7), (257, 260, 1296, 260),
                                            (1, 1, 16, 1), (17, 9, 480, 9), (16, 8, 496, 8), (130, 71, 976, 71)])
def test_apply_vs_oracle(M, Kc, nbytes, cs):
    o = oracle()
    Cm = np.zeros((M, cs), dtype=np.uint8)
    Cm[:, :Kc] = splitmix_bytes(1, M * Kc).reshape(M, Kc)
    D = splitmix_bytes(2, Kc * nbytes).reshape(Kc, nbytes).copy()
    Y0 = np.zeros((M, nbytes), dtype=np.uint8)
    o.orc_gf256_apply(ptr(Cm), cs, M, Kc, ptr(D), nbytes, ptr(Y0), nbytes, nbytes)
    Y = device.apply(dev(Cm)[:, :], dev(D))
    capi.check(capi.lib().oblas_b200_sync(), "sync")  # also surfaces the tensor-core path's status word
    assert np.array_equal(host(Y), Y0)
    Y2 = device.apply(dev(Cm), dev(D), Y=Y.clone(), accumulate=True)
    capi.check(capi.lib().oblas_b200_sync(), "sync")
    assert not host(Y2).any()  # Y ^ Y == 0


def test_apply_medium_linearity_and_spot_rows():
    M = Kc = 1500
    N = 4096
    Cm = torch.empty((M, Kc), dtype=torch.uint8, device="cuda")
    D1 = torch.empty((Kc, N), dtype=torch.uint8, device="cuda")
    D2 = torch.empty((Kc, N), dtype=torch.uint8, device="cuda")
    Cp = torch.zeros((M, 1504), dtype=torch.uint8, device="cuda")
    tmp = torch.empty(M * 1504, dtype=torch.uint8, device="cuda")
    device.fill_random(tmp, 11)
    Cp[:, :Kc] = tmp.reshape(M, 1504)[:, :Kc]
    device.fill_random(D1, 12)
    device.fill_random(D2, 13)
    Y1, Y2, Y12 = device.apply(Cp, D1), device.apply(Cp, D2), device.apply(Cp, D1 ^ D2)
    assert torch.equal(Y1 ^ Y2, Y12)
    Ch, Dh, Yh = host(Cp), host(D1), host(Y1)
    rng = np.random.default_rng(0)
    for r in rng.choice(M, 5, replace=False):
        assert np.array_equal(gf256_vecmat(Ch[r][:Kc], Dh), Yh[r])


@pytest.mark.parametrize("field", [GF2, GF4, GF16, GF256])
def test_nnz_batch(field, cuda_lib):
    """oblas_b200_nnz_batch counts CORRECTLY for every window [s, e).  The reference's gfmat_nnz /
    binmat_nnz (gfmat.c:183-212, binmat.c:87-116) are only right for s == 0 (after a partial first word the
    element index is reused as a word index), so the truth for s != 0 is numpy; the s == 0
    windows are pinned to the oracle's restatement of the reference as well (the bug-compatible host
    symbols gfmat_nnz / binmat_nnz are pinned to the golden vectors in test_gpu_dropin.py)."""
    import ctypes as C
    from helpers import oracle, ptr, u32p
    bits = FIELD_BITS[field]
    rows, cols = 50, 1000
    st = ((cols * bits + 31) // 32 * 4 + 15) // 16 * 16
    M = np.zeros((rows, st), dtype=np.uint8)
    used = (cols * bits + 7) // 8
    M[:, :used] = splitmix_bytes(20 + field, rows * used).reshape(rows, used)
    M[::3] &= 0x11  # sparser rows
    # element values on the CPU
    w = np.unpackbits(M, axis=1, bitorder="little")[:, :cols * bits].reshape(rows, cols, bits)
    nz = w.any(axis=2)
    idx = np.arange(rows, dtype=np.int32)
    for (s, e) in ((0, cols), (3, 20), (40, 100), (33, 999), (0, 0), (500, 5000)):
        out = torch.zeros(rows, dtype=torch.int32, device="cuda")
        dM, di = dev(M), dev(idx)
        capi.check(cuda_lib.oblas_b200_nnz_batch(field, C.c_void_p(dM.data_ptr()), st, cols, C.c_void_p(di.data_ptr()),
                                                 rows, s, e, C.c_void_p(out.data_ptr())))
        want = nz[:, s:min(e, cols)].sum(axis=1)
        assert np.array_equal(host(out), want.astype(np.int32)), (field, s, e)
        if s == 0 and 0 < e <= cols:   # where the reference is right, it agrees
            Mw = np.ascontiguousarray(M).view(np.uint32)
            ref = [oracle().orc_nnz(field, ptr(Mw, u32p), Mw.shape[1], rows, cols, r, s, e) for r in range(rows)]
            assert ref == want.tolist(), (field, s, e)


@pytest.mark.parametrize("M,Kc,ncols", [(40, 24, 512), (300, 130, 992), (17, 9, 32), (33, 20, 496)])
def test_mixed_field_apply_equals_the_oaxpy_b32_composition(M, Kc, ncols, cuda_lib):
    """row f-2: Y = C * expand(Dbits) on the tensor-core apply (the bit rows are expanded inside the operand
    expansion) against (a) the oracle's oaxpy_b32 (oblas_ref.c:19-28) called M * Kc times and (b) the same
    composition made of the library's own OP_AXPY_B32 row kernel, one batched launch per source row."""
    import ctypes as C
    from helpers import OP_AXPY_B32, u32p
    o = oracle()
    Cm = splitmix_bytes(41, M * Kc).reshape(M, Kc).copy()
    words = (ncols + 31) // 32
    Db = splitmix_bytes(42, Kc * words * 4).view(np.uint32).reshape(Kc, words).copy()
    Y0 = np.zeros((M, words * 32), dtype=np.uint8)    # the reference's loop always works on whole mask words
    for i in range(M):
        for j in range(Kc):
            o.orc_axpy_b32(ptr(Y0[i]), ptr(Db[j], u32p), words * 32, int(Cm[i, j]))
    Y0 = np.ascontiguousarray(Y0[:, :ncols])          # the library touches the ncols columns it was given
    dC, dD = dev(Cm), torch.from_numpy(Db.view(np.int32)).cuda()
    Y = torch.full((M, ncols), 0x5A, dtype=torch.uint8, device="cuda")
    capi.check(cuda_lib.oblas_b200_gf256_apply_b32(C.c_void_p(dC.data_ptr()), Kc, M, Kc, C.c_void_p(dD.data_ptr()), words,
                                                  C.c_void_p(Y.data_ptr()), ncols, ncols, 0), "apply_b32")
    device.sync()
    assert np.array_equal(host(Y), Y0)
    # accumulate mode: Y ^= the same product -> zero
    capi.check(cuda_lib.oblas_b200_gf256_apply_b32(C.c_void_p(dC.data_ptr()), Kc, M, Kc, C.c_void_p(dD.data_ptr()), words,
                                                  C.c_void_p(Y.data_ptr()), ncols, ncols, 1), "apply_b32")
    device.sync()
    assert not host(Y).any()
    if ncols % 32 == 0:   # the row kernel reads whole mask words
        Yr = torch.zeros((M, ncols), dtype=torch.uint8, device="cuda")
        rows_i = torch.arange(M, dtype=torch.int32, device="cuda")
        dDb = torch.from_numpy(Db.view(np.uint8).reshape(Kc, words * 4)).cuda()
        for j in range(Kc):
            device.rowops(GF256, OP_AXPY_B32, Yr, dDb, rows_i, torch.full((M,), j, dtype=torch.int32, device="cuda"),
                          dev(Cm[:, j].copy()))
        device.sync()
        assert np.array_equal(host(Yr), Y0)
