"""Bring-up / timing probe of the tensor-core apply path (apply_tc.cuh) against the
shared-memory Four-Russians path on the same inputs.  Not part of the bench contract."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oblas_b200 import capi, device


def run(M, Kc, N, iters=3, check=True):
    lib = capi.lib()
    Cm = torch.empty((M, (Kc + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
    D = torch.empty((Kc, N), dtype=torch.uint8, device="cuda")
    device.fill_random(Cm, 1)
    device.fill_random(D, 2)
    res = {"M": M, "Kc": Kc, "N": N}
    out = {}
    for name, geo, cs in (("lut", 1, 0), ("tc", 3, 1)):
        lib.oblas_b200_set_tuning(-1, geo)
        Y = device.apply(Cm, D)
        rc = lib.oblas_b200_sync()
        if rc:
            res[name + "_error"] = capi.lib().oblas_b200_last_error().decode()
            print(json.dumps(res), flush=True)
            return False
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            device.apply(Cm, D, Y=Y)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name] = Y
        res[name + "_ms"] = min(ts)
        res[name + "_Tmacs"] = M * Kc * N / (min(ts) * 1e-3) / 1e12
    lib.oblas_b200_set_tuning(-1, -1)
    if check:
        diff = (out["lut"] != out["tc"])
        res["mismatch_bytes"] = int(diff.sum().item())
        if res["mismatch_bytes"]:
            idx = diff.nonzero()[:8].tolist()
            res["first_mismatches"] = [(i, j, int(out["lut"][i, j]), int(out["tc"][i, j])) for i, j in idx]
            res["bad_rows"] = int(diff.any(dim=1).sum().item())
            res["bad_cols"] = int(diff.any(dim=0).sum().item())
    print(json.dumps(res), flush=True)
    return res.get("mismatch_bytes", 0) == 0


if __name__ == "__main__":
    device.bind(0)
    sizes = [(16, 8, 480), (16, 64, 480), (100, 100, 1024), (300, 1000, 4096), (2048, 2048, 65536)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        sizes = [(10000, 10000, 65536)]
    ok = True
    for s in sizes:
        ok = run(*s) and ok
        if not ok:
            break
    sys.exit(0 if ok else 1)
