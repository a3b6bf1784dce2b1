"""Per-kernel-class device times of the tensor-core elimination on one shape:
python tools/tc_classes.py [nsys K L]   (UPROF=1 adds the role cycle counters of the update kernel.
OBLAS_B200_TC_DEBUG=1|2|4 switches the loads / the A build / the epilogue of the update kernel off for
timing experiments -- results are then invalid, only the times mean something -- and needs a library
built with  make -C oblas_b200/csrc EXTRA=-DOB2_TC_TIMING_EXPERIMENTS)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device
if os.environ.get("OB2_LIB"):   # a build made for timing experiments (make EXTRA=-DOB2_TC_TIMING_EXPERIMENTS -> other file name)
    capi.LIB_PATH = os.environ["OB2_LIB"]

nsys, K, L = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (592, 1024, 1280)))
device.bind(0)
lib = capi.lib()
A0 = torch.empty((nsys, K, K), dtype=torch.uint8, device="cuda")
B0 = torch.empty((nsys, K, L), dtype=torch.uint8, device="cuda")
device.fill_random(A0, 1)
device.fill_random(B0, 2)
lib.oblas_b200_set_tuning(3, -1)
if os.environ.get("LAZY") is not None:   # LAZY=0: full outer-panel kernel; LAZY=n: compact sets of n rows
    lib.oblas_b200_set_lazy_panels(int(os.environ["LAZY"]))
if os.environ.get("PHASES"):
    lib.oblas_b200_solve_phases(1, None)
names = ["panel", "expand", "update", "permute", "panel-redo"]
for it in range(3):
    A, B = A0.clone(), B0.clone()
    if not os.environ.get("NOTIMES"):   # NOTIMES=1: no per-class event timing
        lib.oblas_b200_solve_tc_times(1, None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r, _ = device.solve_batch(capi.GF256, A, K, B)
    e1.record()
    torch.cuda.synchronize()
    ms = (C.c_double * 5)()
    n = (C.c_uint64 * 5)()
    lib.oblas_b200_solve_tc_times_ex(0, ms, n, 5)
    rc = lib.oblas_b200_sync()
if os.environ.get("UPROF"):
    import numpy as np
    lib.oblas_b200_update_phases(1, None)
    A, B = A0.clone(), B0.clone()
    device.solve_batch(capi.GF256, A, K, B)
    torch.cuda.synchronize()
    up = np.zeros(16, dtype=np.uint64)
    lib.oblas_b200_update_phases(0, up.ctypes.data)
    tiles = float(up[12]) or 1.0
    lab = ["mma:wait A", "mma:wait acc drained", "mma:wait B", "mma:loop", "epi:wait acc full", "epi:readout", "epi:loop",
           "prod:wait free", "prod:loop", "build:wait free", "build:build", "build:loop", "tiles", "epi:wait ld", "epi:pack", "epi:store"]
    print("update_tc_kernel cycles per column tile (%d tiles): " % int(tiles) + ", ".join("%s %.0f" % (lab[k], float(up[k]) / tiles) for k in range(16) if k != 12))
tot = e0.elapsed_time(e1)
print("dbg=%s nsys=%d K=%d L=%d: %.2f ms, %.0f solves/s, full rank %d, sync rc %d | " % (
    os.environ.get("OBLAS_B200_TC_DEBUG", "0"), nsys, K, L, tot, nsys / tot * 1e3, int((r == K).sum()), rc) +
    ", ".join("%s %.2f ms/%d" % (names[k], ms[k], n[k]) for k in range(5)))
st = (C.c_uint64 * 2)()
lib.oblas_b200_lazy_stats(st, 0)
print("lazy outer panels: %d (system, panel) pairs compact, %d given up" % (st[0], st[1]))
if os.environ.get("PHASES"):
    import numpy as np
    ph = np.zeros(8, dtype=np.uint64)
    lib.oblas_b200_solve_phases(0, ph.ctypes.data)
    print("panel kernel phase cycles (sum over CTAs, 3 runs + clones): load %d fact %d inject %d pass %d pro/epilogue %d | cand %d tables %d vprod %d" % tuple(int(x) for x in ph))

