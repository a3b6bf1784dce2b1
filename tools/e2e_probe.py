import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oblas_b200 import capi, device
device.bind(0)
L = capi.lib()
K, Lb, n = 1024, 1280, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
A0 = torch.empty((n, K, K), dtype=torch.uint8, device="cuda"); B0 = torch.empty((n, K, Lb), dtype=torch.uint8, device="cuda")
device.fill_random(A0, 1); device.fill_random(B0, 2)
hA = L.oblas_b200_host_alloc(A0.numel()); hB = L.oblas_b200_host_alloc(B0.numel())
rank = np.zeros(n, dtype=np.int32)
for it in range(3):
    L.oblas_b200_copy(hA, A0.data_ptr(), A0.numel()); L.oblas_b200_copy(hB, B0.data_ptr(), B0.numel()); L.oblas_b200_sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    capi.check(L.oblas_b200_solve_batch_host(capi.GF256, hA, K, K * K, K, K, K, hB, Lb, K * Lb, Lb, rank.ctypes.data, None, n, chunk))
    dt = time.perf_counter() - t0
    print("e2e n=%d chunk=%d: %.3f s  %.0f solves/s" % (n, chunk, dt, n / dt), flush=True)
