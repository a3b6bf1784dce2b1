"""Exploratory device timings (not the bench contract): row-op GB/s per back-end,
elimination and apply throughput at a few sizes.  CUDA events, warm-up, inputs > L2."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device
from oblas_b200.capi import GF2, GF4, GF16, GF256, OP_ADD, OP_AXPY, OP_SCAL, OP_SWAP


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
    return min(ts), sum(ts) / len(ts)


def out(**kw):
    print(json.dumps(kw), flush=True)


def rowops(nmat=512, rows=1024, stride=1280):
    R = nmat * rows
    a = torch.empty((R, stride), dtype=torch.uint8, device="cuda")
    b = torch.empty((R, stride), dtype=torch.uint8, device="cuda")
    device.fill_random(a, 1)
    device.fill_random(b, 2)
    dst = torch.randperm(R, device="cuda").to(torch.int32)
    src = torch.randint(0, R, (R,), device="cuda", dtype=torch.int32)
    u = torch.randint(2, 256, (R,), device="cuda", dtype=torch.uint8)
    t, _ = timeit(lambda: b.copy_(a))
    out(what="torch copy", GBs=2 * R * stride / t / 1e9)
    for name, op, mul, mult in (("oaxpy lin", OP_AXPY, capi.MUL_LIN, 3), ("oaxpy lut", OP_AXPY, capi.MUL_LUT, 3),
                                ("oaddrow", OP_ADD, -1, 3), ("oscal lin", OP_SCAL, capi.MUL_LIN, 2),
                                ("oscal lut", OP_SCAL, capi.MUL_LUT, 2)):
        t, tav = timeit(lambda: device.rowops(GF256, op, a, b, dst, src, u, mul=mul))
        out(what=name, rows=R, stride=stride, ms=t * 1e3, ms_avg=tav * 1e3, GBs=mult * R * stride / t / 1e9)
    # sequential rows (dst = identity, src = identity): pure streaming
    ident = torch.arange(R, device="cuda", dtype=torch.int32)
    t, _ = timeit(lambda: device.rowops(GF256, OP_AXPY, a, b, ident, ident, u, mul=capi.MUL_LIN))
    out(what="oaxpy lin (identity rows)", GBs=3 * R * stride / t / 1e9)
    # cfg1 as given: one matrix pair, L2 resident
    a1, b1 = a[:rows], b[:rows]
    d1 = torch.randperm(rows, device="cuda").to(torch.int32)
    s1 = torch.randint(0, rows, (rows,), device="cuda", dtype=torch.int32)
    t, _ = timeit(lambda: device.rowops(GF256, OP_AXPY, a1, b1, d1, s1, u[:rows]), iters=20, warm=5)
    out(what="oaxpy cfg1 single matrix (L2 resident, launch bound)", us=t * 1e6, GBs=3 * rows * stride / t / 1e9)


def solve(field, K, L, nsys, cols=None):
    bits = capi.FIELD_BITS[field]
    a_bytes = (K * bits // 8 + 15) // 16 * 16
    A0 = torch.empty((nsys, K, a_bytes), dtype=torch.uint8, device="cuda")
    device.fill_random(A0, 3)
    B0 = None
    if L:
        B0 = torch.empty((nsys, K, L), dtype=torch.uint8, device="cuda")
        device.fill_random(B0, 4)
    A, B = A0.clone(), (B0.clone() if L else None)

    def run():
        A.copy_(A0)
        if L:
            B.copy_(B0)
        return device.solve_batch(field, A, K, B)

    def restore_only():
        A.copy_(A0)
        if L:
            B.copy_(B0)

    tr, _ = timeit(restore_only, iters=3, warm=1)
    t, tav = timeit(run, iters=3, warm=1)
    rank, _ = run()
    torch.cuda.synchronize()
    full = int((rank == K).sum())
    out(what="solve", field=field, K=K, L=L, nsys=nsys, s=t - tr, solves_per_s=nsys / (t - tr), full_rank=full)


def apply(M, Kc, N):
    Cm = torch.empty((M, (Kc + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
    D = torch.empty((Kc, N), dtype=torch.uint8, device="cuda")
    device.fill_random(Cm, 5)
    device.fill_random(D, 6)
    Y = torch.empty((M, N), dtype=torch.uint8, device="cuda")
    t, _ = timeit(lambda: device.apply(Cm, D, Y), iters=3, warm=1)
    out(what="apply", M=M, Kc=Kc, N=N, s=t, gmacs_per_s=M * Kc * N / t / 1e9)


if __name__ == "__main__":
    device.bind(0)
    which = sys.argv[1:] or ["rowops", "solve", "apply"]
    out(gpu=torch.cuda.get_device_name(0), sms=capi.lib().oblas_b200_sm_count())
    if "rowops" in which:
        rowops()
    geos = [int(x) for x in os.environ.get("GEOS", "0,1,2").split(",")]
    for geo in geos:
        capi.lib().oblas_b200_set_tuning(geo, geo)
        out(geo=geo)
        if "solve" in which:
            solve(GF256, 1024, 1280, 592)
            solve(GF256, 256, 1280, 592)
            solve(GF16, 2048, 0, 148)
            solve(GF4, 2048, 0, 148)
            solve(GF2, 2048, 0, 148)
            solve(GF2, 8192, 0, 37)
        if "apply" in which:
            apply(2048, 2048, 65536)
            apply(10000, 10000, 8192)
