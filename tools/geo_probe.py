"""Table-kernel geometries on the BASELINE shapes of the packed fields: python tools/geo_probe.py
(solve_geo 0 = 4-bit groups / 256-byte tiles, 1 = 6 / 128, 2 = 7 / 64, 4 = 5 / 128)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device

device.bind(0)
lib = capi.lib()
for name, field, K, nsys in (("GF2 K=8192", capi.GF2, 8192, 148), ("GF16 K=2048", capi.GF16, 2048, 592), ("GF4 K=2048", capi.GF4, 2048, 592),
                             ("GF256 K=1024 (table kernel)", capi.GF256, 1024, 296)):
    bits = capi.FIELD_BITS[field]
    a_bytes = K * bits // 8
    A0 = torch.empty((nsys, K, a_bytes), dtype=torch.uint8, device="cuda")
    device.fill_random(A0, 100 + field)
    ref = None
    for geo in (-1, 0, 1, 2, 4):
        lib.oblas_b200_set_tuning(geo if geo >= 0 else (1 if field == capi.GF256 else -1), -1)
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
        print("%-28s geo %2d -> NBW=%d KB=%d WT=%3d : %8.2f ms, %8.0f solves/s, same bytes %s" % (name, geo, g[0], g[1], g[2], best, nsys / best * 1e3, same), flush=True)
lib.oblas_b200_set_tuning(-1, -1)
