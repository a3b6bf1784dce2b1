"""Tiny workload for compute-sanitizer (memcheck / racecheck): one elimination per field,
row ops, apply.  Results are checked against the oracle as in smoke()."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import GF2, GF4, GF16, GF256, OP_AXPY, OP_SWAP, oracle_solve, random_matrix, splitmix_bytes
from oblas_b200 import capi, device
device.bind(0)
for general in (0, 1):
    capi.lib().oblas_b200_set_panel_path(general)
    for field, K in ((GF256, 80), (GF16, 96), (GF4, 128), (GF2, 200)):
        A = np.stack([random_matrix(field, K, K, 10 + s, rank_deficient=s) for s in range(2)])
        B = np.stack([splitmix_bytes(20 + s, K * 144).reshape(K, 144) for s in range(2)]).copy()
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        rank, _ = device.solve_batch(field, dA, K, dB)
        torch.cuda.synchronize()
        A0, B0 = A.copy(), B.copy()
        want = [oracle_solve(field, A0[s], B0[s], K)[0] for s in range(2)]
        assert rank.cpu().tolist() == want and np.array_equal(dA.cpu().numpy(), A0) and np.array_equal(dB.cpu().numpy(), B0)
a = torch.from_numpy(splitmix_bytes(1, 64 * 1280).reshape(64, 1280).copy()).cuda()
idx = torch.arange(64, dtype=torch.int32, device="cuda")
u = torch.arange(64, dtype=torch.uint8, device="cuda") + 2
device.rowops(GF256, OP_AXPY, a, a.clone(), idx, idx.flip(0), u)
device.rowops(GF256, OP_SWAP, a, a, idx[:32], idx[32:], None)
Cm = torch.from_numpy(splitmix_bytes(3, 300 * 112).reshape(300, 112).copy()).cuda()
D = torch.from_numpy(splitmix_bytes(4, 100 * 272).reshape(100, 272).copy()).cuda()
device.apply(Cm, D)
torch.cuda.synchronize()
print("sanitize workload ok")
