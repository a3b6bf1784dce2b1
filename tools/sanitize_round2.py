"""Small workload for compute-sanitizer (memcheck) over the round-2 code paths: the tensor-core elimination with
even column tiles (full and lazy outer panels) and the tall-system table kernel (per-row state in global memory),
each checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import GF2, GF4, GF16, GF256, oracle_solve, random_matrix, splitmix_bytes
from oblas_b200 import capi, device
device.bind(0)
lib = capi.lib()


def check(field, K, cols, L, nsys, seed):
    A = np.stack([random_matrix(field, K, cols, seed + s, rank_deficient=s % 2) for s in range(nsys)])
    B = np.stack([splitmix_bytes(seed + 50 + s, K * L).reshape(K, L) for s in range(nsys)]).copy() if L else None
    dA, dB = torch.from_numpy(A).cuda(), (torch.from_numpy(B).cuda() if L else None)
    rank, _ = device.solve_batch(field, dA, cols, dB)
    torch.cuda.synchronize()
    capi.check(lib.oblas_b200_sync(), "sync")
    A0, B0 = A.copy(), (B.copy() if L else None)
    want = [oracle_solve(field, A0[s], B0[s] if L else None, cols)[0] for s in range(nsys)]
    assert rank.cpu().tolist() == want, (rank, want)
    assert np.array_equal(dA.cpu().numpy(), A0) and (not L or np.array_equal(dB.cpu().numpy(), B0))


lib.oblas_b200_set_tuning(3, -1)          # tensor-core elimination, 3 outer panels, ranges 400 / 272 / 144 columns
for lazy in (0, 144):
    lib.oblas_b200_set_lazy_panels(lazy)
    check(GF256, 384, 384, 144, 3, 100)
lib.oblas_b200_set_lazy_panels(-1)
lib.oblas_b200_set_tuning(-1, -1)
check(GF2, 4000, 512, 32, 2, 200)         # tall table-kernel systems: rows x 35 B do not fit beside the 6-bit tables
check(GF16, 3600, 128, 16, 2, 300)
lib.oblas_b200_set_tuning(1, -1)
check(GF256, 2800, 48, 32, 1, 400)
lib.oblas_b200_set_tuning(-1, -1)
print("sanitize round-2 workload ok")
