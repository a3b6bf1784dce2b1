"""Tall systems on the table kernel: 128-bit panels with the per-row state in global memory (tall mode 1, the
default) against 64-bit panels in shared memory (mode 0): python tools/tall_probe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device

device.bind(0)
lib = capi.lib()
for name, field, K, nsys, geos in (("GF2 K=8192", capi.GF2, 8192, 256, (-1, 4)), ("GF16 K=4096", capi.GF16, 4096, 148, (-1, 4)),
                                   ("GF4 K=8192", capi.GF4, 8192, 148, (-1, 4)), ("GF16 K=2048", capi.GF16, 2048, 592, (-1,)),
                                   ("GF4 K=2048", capi.GF4, 2048, 592, (-1,)), ("GF256 K=448", capi.GF256, 448, 592, (-1,))):
    bits = capi.FIELD_BITS[field]
    a_bytes = K * bits // 8
    A0 = torch.empty((nsys, K, a_bytes), dtype=torch.uint8, device="cuda")
    device.fill_random(A0, 100 + field)
    ref = None
    for mode, geo in [(0, -1)] + [(1, g) for g in geos] + [(2, -1)]:   # mode 2: 6-bit tables + global row state also where 5-bit tables fit
        lib.oblas_b200_set_tall_mode(mode)
        lib.oblas_b200_set_tuning(geo, -1)
        g = (C.c_int * 4)()
        path = C.c_int(0)
        lib.oblas_b200_solve_plan(field, K, K, a_bytes, 0, nsys, C.byref(path), None, g)
        best = None
        for it in range(3):
            A = A0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r, _ = device.solve_batch(field, A, K, None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        same = True if ref is None else bool(torch.equal(A, ref))
        ref = A if ref is None else ref
        print("%-14s tall mode %d geo %2d -> NBW=%d KB=%d WT=%3d : %8.2f ms, %8.0f solves/s, same bytes %s" % (
            name, mode, geo, g[0], g[1], g[2], best, nsys / best * 1e3, same), flush=True)
lib.oblas_b200_set_tuning(-1, -1)
lib.oblas_b200_set_tall_mode(1)
