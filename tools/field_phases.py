"""Phase shares of the table kernel on a packed-field batch: python tools/field_phases.py FIELD K NSYS [TALL]
(FIELD 0 = GF(2), 1 = GF(4), 2 = GF(16), 3 = GF(256); cycles of thread 0 per phase, summed over CTAs)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oblas_b200 import capi, device

field, K, nsys = (int(x) for x in sys.argv[1:4])
tall = int(sys.argv[4]) if len(sys.argv) > 4 else 1
device.bind(0)
lib = capi.lib()
lib.oblas_b200_set_tall_mode(tall)
a_bytes = K * capi.FIELD_BITS[field] // 8
A0 = torch.empty((nsys, K, a_bytes), dtype=torch.uint8, device="cuda")
device.fill_random(A0, 100 + field)
for it in range(2):
    A = A0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r, _ = device.solve_batch(field, A, K, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
ph = np.zeros(8, dtype=np.uint64)
lib.oblas_b200_solve_phases(1, None)
A = A0.clone()
device.solve_batch(field, A, K, None)
torch.cuda.synchronize()
lib.oblas_b200_solve_phases(0, ph.ctypes.data)
tot = float(ph[:4].sum()) or 1.0
print("field %d K=%d nsys=%d tall=%d: %.2f ms, %.0f solves/s | M clk per system: load %.2f, factorisation %.2f, update %.2f, permutation %.2f | shares %s" % (
    field, K, nsys, tall, ms, nsys / ms * 1e3, *(float(x) / nsys / 1e6 for x in ph[:4]),
    ", ".join("%.1f%%" % (100 * float(x) / tot) for x in ph[:4])))
