"""One tensor-core apply (for ncu): python tools/tc_prof.py [M Kc N cluster]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device

M, Kc, N, cs = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (2048, 2048, 65536, 1)))
device.bind(0)
lib = capi.lib()
Cm = torch.empty((M, (Kc + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
D = torch.empty((Kc, N), dtype=torch.uint8, device="cuda")
device.fill_random(Cm, 1)
device.fill_random(D, 2)
lib.oblas_b200_set_tuning(-1, 3)
Y = device.apply(Cm, D)
capi.check(lib.oblas_b200_sync(), "sync")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
device.apply(Cm, D, Y=Y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("apply_tc M=%d Kc=%d N=%d cluster=%d: %.3f ms, %.2f T GF-MAC/s" % (M, Kc, N, cs, ms, M * Kc * N / ms / 1e9))
