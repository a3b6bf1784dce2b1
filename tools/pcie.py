"""Host<->device copy bandwidth with pinned memory (explains the e2e number)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oblas_b200 import capi, device
device.bind(0)
L = capi.lib()
n = 2 << 30
h = L.oblas_b200_host_alloc(n); h2 = L.oblas_b200_host_alloc(n)
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); L.oblas_b200_sync(); torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    print("H2D GB/s", n / t(lambda: L.oblas_b200_copy(d.data_ptr(), h, n)) / 1e9)
    print("D2H GB/s", n / t(lambda: L.oblas_b200_copy(h, d.data_ptr(), n)) / 1e9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
hp = torch.empty(n, dtype=torch.uint8, pin_memory=True); hp2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(hp, non_blocking=True)
with torch.cuda.stream(s2): hp2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("bidirectional: each direction GB/s", n / dt / 1e9)
