"""One batched oscal / oaxpy launch on the bench's working set (for ncu): python tools/rowop_one.py [scal|axpy]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oblas_b200 import capi, device
device.bind(0)
R, stride = 512 * 1024, 1280
a = torch.empty((R, stride), dtype=torch.uint8, device="cuda"); b = torch.empty((R, stride), dtype=torch.uint8, device="cuda")
device.fill_random(a, 1); device.fill_random(b, 2)
dst = torch.randperm(R, device="cuda").to(torch.int32); src = torch.randint(0, R, (R,), device="cuda", dtype=torch.int32)
u = torch.randint(2, 256, (R,), device="cuda", dtype=torch.uint8)
op = capi.OP_SCAL if (len(sys.argv) < 2 or sys.argv[1] == "scal") else capi.OP_AXPY
for _ in range(2):
    device.rowops(capi.GF256, op, a, b, dst, src, u)
torch.cuda.synchronize()
print("done")
