"""ONE process, all visible GPUs through the C ABI (oblas_b200_solve_batch_host_multi,
oblas_b200_gf256_apply_sharded): BASELINE configs[1] (batch sharded by systems, host buffers) and
configs[4] (apply 10000^2 x 64 KB, symbol columns sharded, coefficients broadcast by peer copies).
    python tools/multi_gpu_capi.py [max_devices] [--small]
One JSON line per (config, device count)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from oblas_b200 import capi

L = capi.lib()
small = "--small" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
have = L.oblas_b200_device_count()
maxdev = min(have, int(args[0])) if args else have
counts = [n for n in (1, 2, 4, 8) if n <= maxdev]
capi.init(0)


def pinned(nbytes):
    p = L.oblas_b200_host_alloc(nbytes)
    assert p, "pinned allocation failed"
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))


def fill(arr, seed):
    from helpers import oracle, ptr
    step = 1 << 28
    for off in range(0, arr.size, step):
        n = min(step, arr.size - off)
        oracle().orc_fill_random(ptr(arr[off:off + n]), n, seed, off)


# ---- configs[4]
K, N = (2000, 8192) if small else (10000, 65536)
cs = (K + 63) // 64 * 64
pC, hC = pinned(K * cs)
pD, hD = pinned(K * N)
pY, hY = pinned(K * N)
fill(hC, 7)
fill(hD, 8)
ref = None
for nd in counts:
    secs = (C.c_double * 3)()
    best = None
    for it in range(2):
        capi.check(L.oblas_b200_gf256_apply_sharded(pC, cs, K, K, pD, N, pY, N, N, None, nd, secs), "apply_sharded")
        best = list(secs) if best is None or secs[2] < best[2] else best
    Yv = hY.reshape(K, N)
    digest = int(np.bitwise_xor.reduce(Yv[::97].view(np.uint64).ravel()))
    ref = digest if ref is None else ref
    print(json.dumps({"config": "configs[4] apply C %dx%d, D %d x %d bytes, HOST buffers, one process" % (K, K, K, N), "devices": nd,
                      "coefficient_upload_and_peer_broadcast_s": best[0], "slowest_device_h2d_apply_d2h_s": best[1], "call_s": best[2],
                      "T_GF_MAC_per_s_call": K * K * N / best[2] / 1e12, "same_result_as_1_device": digest == ref}), flush=True)
for p in (pC, pD, pY):
    L.oblas_b200_host_free(p)

# ---- configs[1]
Ks, Ls = 1024, 1280
per_dev = 256 if small else 2048
for nd in counts:
    nsys = per_dev * nd
    pA, hA = pinned(nsys * Ks * Ks)
    pB, hB = pinned(nsys * Ks * Ls)
    rank = np.zeros(nsys, dtype=np.int32)
    best = None
    for it in range(2):
        fill(hA, 1000)
        fill(hB, 1001)
        t0 = time.perf_counter()
        capi.check(L.oblas_b200_solve_batch_host_multi(capi.GF256, pA, Ks, Ks * Ks, Ks, Ks, Ks, pB, Ls, Ks * Ls, Ls,
                                                       rank.ctypes.data, None, nsys, None, nd, 0), "solve_host_multi")
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    print(json.dumps({"config": "configs[1] K=1024, L=1280, %d systems per device, HOST buffers, one process" % per_dev, "devices": nd,
                      "systems": nsys, "call_s": best, "solves_per_s": nsys / best, "full_rank": int((rank == Ks).sum())}), flush=True)
    L.oblas_b200_host_free(pA)
    L.oblas_b200_host_free(pB)
