"""Time the GF(256) elimination paths on one shape: python tools/solve_probe.py [nsys K L geo]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device

nsys, K, L, geo = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (592, 1024, 1280, 3)))
device.bind(0)
lib = capi.lib()
A0 = torch.empty((nsys, K, K), dtype=torch.uint8, device="cuda")
B0 = torch.empty((nsys, K, L), dtype=torch.uint8, device="cuda")
device.fill_random(A0, 1)
device.fill_random(B0, 2)
lib.oblas_b200_set_tuning(geo, -1)
for it in range(2):
    A, B = A0.clone(), B0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r, _ = device.solve_batch(capi.GF256, A, K, B)
    e1.record()
    torch.cuda.synchronize()
    capi.check(lib.oblas_b200_sync(), "sync")
    ms = e0.elapsed_time(e1)
import numpy as np
ph = np.zeros(8, dtype=np.uint64)
lib.oblas_b200_solve_phases(1, None)
A, B = A0.clone(), B0.clone()
device.solve_batch(capi.GF256, A, K, B)
torch.cuda.synchronize()
lib.oblas_b200_solve_phases(0, ph.ctypes.data)
tot = float(ph[:5].sum()) or 1.0
print("inside factorisation (M clk per system): candidates %.3f, tables %.3f, V product %.3f" % tuple(float(x) / nsys / 1e6 for x in ph[5:8]))
print("phase cycles per system (M clk): " + ", ".join("%.3f" % (float(x) / nsys / 1e6) for x in ph[:5]) +
      "  shares: " + ", ".join("%.1f%%" % (100 * float(x) / tot) for x in ph[:5]))
print("solve geo=%d nsys=%d K=%d L=%d: %.2f ms, %.0f solves/s, full rank %d" % (geo, nsys, K, L, ms, nsys / ms * 1e3, int((r == K).sum())))
