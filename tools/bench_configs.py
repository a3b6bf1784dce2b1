"""Throughput + parity of the OTHER BASELINE.json configs (0, 2, 3, 4) -- the default
bench.py line is configs[1].  One JSON line per config; `--gpus N` under torchrun shards
batches by systems and the apply by symbol-column slices (one NCCL broadcast of C).

    python tools/bench_configs.py [cfg1] [cfg3] [cfg4] [cfg5] [--small]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from gpu_helpers import gf256_vecmat
from helpers import GF2, GF4, GF16, GF256, FIELD_BITS, mat_view, oracle_solve, ptr, reference, u32p
from oblas_b200 import capi, device, shard

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def out(**kw):
    if RANK == 0:
        print(json.dumps(kw), flush=True)


def sync_max(ms):
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if WORLD > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def timed(fn, steps=3, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return sync_max(e0.elapsed_time(e1) / steps)


def ref_lib():
    for v in ("avx512", "avx2", "ref"):
        lib = reference(v)
        if lib is not None:
            return lib, v
    return None, "oracle port"


def cfg1():
    """configs[0]: 1024 rows x 1280-byte rows, single matrix pair (L2 resident) -- latency figure"""
    rows, st = 1024, 1280
    a = torch.empty((rows, st), dtype=torch.uint8, device="cuda")
    b = torch.empty((rows, st), dtype=torch.uint8, device="cuda")
    device.fill_random(a, 1)
    device.fill_random(b, 2)
    dst = torch.randperm(rows, device="cuda").to(torch.int32)
    src = torch.randint(0, rows, (rows,), device="cuda", dtype=torch.int32)
    u = torch.randint(2, 256, (rows,), device="cuda", dtype=torch.uint8)
    res = {}
    for name, op, mult in (("oaxpy", capi.OP_AXPY, 3), ("oaddrow", capi.OP_ADD, 3), ("oscal", capi.OP_SCAL, 2)):
        ms = timed(lambda: device.rowops(GF256, op, a, b, dst, src, u), steps=50, warm=10)
        res[name] = {"us_per_launch": ms * 1e3, "GBs": mult * rows * st / (ms / 1e3) / 1e9}
    # packed fields on config-4 shaped rows (K=2048 -> 1024 / 512-byte rows), HBM streaming
    for field, nm, rb in ((GF16, "gf16", 1024), (GF4, "gf4_fixed", 512), (GF4, "gf4_shipped", 512)):
        R = 1 << 20
        A = torch.empty((R, rb), dtype=torch.uint8, device="cuda")
        B = torch.empty((R, rb), dtype=torch.uint8, device="cuda")
        device.fill_random(A, 3)
        device.fill_random(B, 4)
        d2 = torch.randperm(R, device="cuda").to(torch.int32)
        s2 = torch.randint(0, R, (R,), device="cuda", dtype=torch.int32)
        u2 = torch.randint(2, 1 << FIELD_BITS[field], (R,), device="cuda", dtype=torch.uint8)
        variant = 0 if nm == "gf4_shipped" else 1
        ms = timed(lambda: device.rowops(field, capi.OP_AXPY, A, B, d2, s2, u2, variant=variant), steps=5, warm=3)
        res["gfmat_axpy_" + nm] = {"GBs": 3 * R * rb / (ms / 1e3) / 1e9, "rows": R, "row_bytes": rb}
        del A, B
    out(config="configs[0]: GF(256) oaxpy/oaddrow/oscal, one 1024x1280 matrix pair per launch (L2 resident, "
               "launch bound) + packed-field axpy streaming", results=res)


def solve_cfg(name, field, K, nsys_total, cpu_systems=1):
    bits = FIELD_BITS[field]
    a_bytes = (K * bits // 8 + 15) // 16 * 16
    lo, hi = shard.batch_range(nsys_total, RANK, WORLD)
    n = hi - lo
    A0 = torch.empty((n, K, a_bytes), dtype=torch.uint8, device="cuda")
    device.fill_random(A0, 100 + field, lo * K * a_bytes)
    A = torch.empty_like(A0)
    state = {}

    def step():
        A.copy_(A0)
        state["rank"], state["piv"] = device.solve_batch(field, A, K, None, want_pivcols=True)

    t_restore = timed(lambda: A.copy_(A0), steps=2, warm=1)
    ms = timed(step, steps=2, warm=1)
    rank = state["rank"].cpu().numpy()
    # parity: system 0 of rank 0 bit for bit against the reference loop on the CPU
    parity = None
    cpu = None
    if RANK == 0:
        lib, vname = ref_lib()
        t0 = time.perf_counter()
        a_h = A0[0].cpu().numpy()
        if lib is not None:
            if field == GF2:
                m = lib.binmat_new(K, K)
            else:
                m = lib.gfmat_new(field, K, K)
            assert mat_view(m).shape[1] == a_bytes
            mat_view(m)[:] = a_h
            pv = np.zeros(K, dtype=np.uint32)
            t0 = time.perf_counter()
            rk = lib.refdrv_binmat_rref(m, None, ptr(pv, u32p)) if field == GF2 else \
                lib.refdrv_gfmat_solve(m, None, ptr(pv, u32p), 1)
            dt = time.perf_counter() - t0
            want = mat_view(m).copy()
            (lib.binmat_free if field == GF2 else lib.gfmat_free)(m)
        else:
            want = a_h.copy()
            t0 = time.perf_counter()
            rk, pv = oracle_solve(field, want, None, K)
            dt = time.perf_counter() - t0
        got = A[0].cpu().numpy()
        gp = state["piv"][0].cpu().numpy().astype(np.uint32)
        parity = {"system": 0, "rank_ref": int(rk), "rank_gpu": int(rank[0]), "bit_exact": bool(np.array_equal(got, want)),
                  "pivcols_equal": bool(np.array_equal(gp[:rk], pv[:rk])), "checked_against": vname}
        cpu = {"value": 1.0 / dt, "unit": "solves/s", "cores": 1, "kind": "reference" if lib is not None else "port",
               "sample": "1 system on 1 thread (%s build + oracle/ref_driver.c): %.2f s" % (vname, dt)}
    kern_ms = ms - t_restore
    out(config=name, metric="solves/s", value=nsys_total / (kern_ms / 1e3), n_gpus=WORLD, ms_per_batch=kern_ms,
        systems=nsys_total, K=K, field_bits=bits, full_rank_fraction=float((rank == K).mean()),
        mean_rank=float(rank.mean()), parity=parity, cpu_baseline=cpu)
    del A, A0


def cfg5(small=False):
    """configs[4]: Y = C*D, C 10000x10000, D 10000 x 65536 bytes, column-sharded"""
    K = 2000 if small else 10000
    N = 8192 if small else 65536
    cs = (K + 15) // 16 * 16
    lo, hi = shard.column_slice(N, RANK, WORLD)
    Cm = torch.zeros((K, cs), dtype=torch.uint8, device="cuda")
    if RANK == 0:
        tmp = torch.empty((K, cs), dtype=torch.uint8, device="cuda")
        device.fill_random(tmp, 7)
        Cm[:, :K] = tmp[:, :K]
        del tmp
    t0 = time.perf_counter()
    shard.broadcast_coefficients(Cm, src=0)   # the one collective of this config (NCCL)
    torch.cuda.synchronize()
    t_bcast = time.perf_counter() - t0
    w = hi - lo
    D = torch.empty((K, w), dtype=torch.uint8, device="cuda")
    # column slice of the global D: row r of the slice = bytes [r*N+lo, r*N+hi) of the stream
    full_row = torch.empty(N, dtype=torch.uint8, device="cuda")
    for r in range(K):
        if WORLD == 1:
            break
        device.fill_random(full_row, 8, r * N)
        D[r] = full_row[lo:hi]
    if WORLD == 1:
        device.fill_random(D, 8)
    Y = torch.empty((K, w), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: device.apply(Cm, D, Y), steps=2, warm=1)
    # parity: sampled rows recomputed on the CPU (numpy), every rank checks its own slice
    Ch = Cm.cpu().numpy()
    rng = np.random.default_rng(RANK)
    cols = slice(0, min(w, 1024))
    Dh = D[:, cols].cpu().numpy()
    ok = True
    for r in rng.choice(K, 4, replace=False):
        ok &= bool(np.array_equal(gf256_vecmat(Ch[r][:K], Dh), Y[r, cols].cpu().numpy()))
    okt = torch.tensor([1 if ok else 0], device="cuda")
    if WORLD > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    cpu = None
    if RANK == 0:
        lib, vname = ref_lib()
        if lib is not None:
            rows_s, width = 256, 1024
            c_, d_, y_ = lib.octmat_new(rows_s, K), lib.octmat_new(K, width), lib.octmat_new(rows_s, width)
            mat_view(c_)[:, :K] = Ch[:rows_s, :K]
            mat_view(d_)[:, :width] = Dh[:, :width] if Dh.shape[1] >= width else 0
            t0 = time.perf_counter()
            lib.refdrv_octmat_apply(c_, d_, y_)
            dt = time.perf_counter() - t0
            macs = rows_s * K * width
            cpu = {"value": macs / dt / 1e9, "unit": "G GF-MAC/s", "cores": 1, "kind": "reference",
                   "sample": "%d rows x %d coefficient columns x %d symbol bytes via oaxpy (%s build): %.2f s"
                             % (rows_s, K, width, vname, dt)}
            for m in (c_, d_, y_):
                lib.octmat_free(m)
    macs = float(K) * K * N
    out(config="configs[4]: GF(256) apply, C %dx%d, D %d x %d bytes, column-sharded over %d GPU(s)" % (K, K, K, N, WORLD),
        metric="T GF-MAC/s", value=macs / (ms / 1e3) / 1e12, n_gpus=WORLD, ms_per_apply=ms, scaling="strong",
        broadcast_s=t_bcast, coefficient_bytes=int(Cm.numel()), parity_rows_ok=bool(int(okt[0])), cpu_baseline=cpu)


def main():
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    device.bind(LOCAL)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    small = "--small" in sys.argv
    which = args or ["cfg1", "cfg3", "cfg4", "cfg5"]
    if "cfg1" in which:
        cfg1()   # every rank runs it (its timing helper contains collectives); rank 0 reports
    if WORLD > 1:
        dist.barrier()
    if "cfg3" in which:
        solve_cfg("configs[2]: GF(2) binmat elimination K=8192, batch of 256", GF2, 2048 if small else 8192,
                  64 if small else 256)
    if "cfg4" in which:
        solve_cfg("configs[3]: GF(16) elimination K=2048, batch of 1024", GF16, 2048, 128 if small else 1024)
        solve_cfg("configs[3]: GF(4) elimination K=2048, batch of 1024 (corrected table)", GF4, 2048,
                  128 if small else 1024)
    if "cfg5" in which:
        cfg5(small)
    if WORLD > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
