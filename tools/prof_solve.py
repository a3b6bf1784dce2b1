"""Small fixed workload for ncu: a few launches of the elimination kernel (296 systems of
K=1024, L=1280) with the table geometry given on the command line."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device

geo = int(sys.argv[1]) if len(sys.argv) > 1 else -1
nsys = int(sys.argv[2]) if len(sys.argv) > 2 else 296
field = int(sys.argv[3]) if len(sys.argv) > 3 else capi.GF256
K = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
L = int(sys.argv[5]) if len(sys.argv) > 5 else 1280
device.bind(0)
capi.lib().oblas_b200_set_tuning(geo, geo)
capi.lib().oblas_b200_set_panel_path(int(os.environ.get("PANEL_PATH", "0")))   # bit 0: general panel path, bit 1: no look-ahead
a_bytes = (K * capi.FIELD_BITS[field] // 8 + 15) // 16 * 16
A0 = torch.empty((nsys, K, a_bytes), dtype=torch.uint8, device="cuda")
device.fill_random(A0, 3)
B0 = None
if L:
    B0 = torch.empty((nsys, K, L), dtype=torch.uint8, device="cuda")
    device.fill_random(B0, 4)
import ctypes, numpy as np
if os.environ.get("PHASES"):
    capi.lib().oblas_b200_solve_phases(1, None)
for _ in range(3):
    A, B = A0.clone(), (B0.clone() if L else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rank, _ = device.solve_batch(field, A, K, B)
    e1.record()
    torch.cuda.synchronize()
    print("geo", geo, "ms", e0.elapsed_time(e1), "solves/s", nsys / e0.elapsed_time(e1) * 1e3, "full", int((rank == K).sum()))
if os.environ.get("PHASES"):
    ph = np.zeros(8, dtype=np.uint64)
    capi.lib().oblas_b200_solve_phases(0, ph.ctypes.data)
    ph = ph[:4]
    tot = ph.sum()
    print("phases (share of CTA cycles): load %.3f  panel %.3f  update %.3f  permute %.3f" % tuple(ph / tot))
