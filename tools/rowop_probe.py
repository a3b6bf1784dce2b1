"""GF(256) oaxpy / oaddrow / oscal streaming GB/s on the bench's working set: python tools/rowop_probe.py
(OB2_LIB=... selects another build of the library)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device
if os.environ.get("OB2_LIB"):
    capi.LIB_PATH = os.environ["OB2_LIB"]
device.bind(0)
nmat, rows, stride = 512, 1024, 1280
R = nmat * rows
a = torch.empty((R, stride), dtype=torch.uint8, device="cuda")
b = torch.empty((R, stride), dtype=torch.uint8, device="cuda")
device.fill_random(a, 1)
device.fill_random(b, 2)
dst = torch.randperm(R, device="cuda").to(torch.int32)
src = torch.randint(0, R, (R,), device="cuda", dtype=torch.int32)
u = torch.randint(2, 256, (R,), device="cuda", dtype=torch.uint8)
res = []
for name, op, mult in (("oaxpy", capi.OP_AXPY, 3), ("oaddrow", capi.OP_ADD, 3), ("oscal", capi.OP_SCAL, 2)):
    for _ in range(3):
        device.rowops(capi.GF256, op, a, b, dst, src, u)
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        device.rowops(capi.GF256, op, a, b, dst, src, u)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
    t = sorted(ts)[len(ts) // 2]
    res.append("%s %.0f GB/s" % (name, mult * R * stride / t / 1e9))
print(os.environ.get("OB2_LIB", "default").split("/")[-1], " | ".join(res))
