"""Per-launch table of an `ncu --set full` capture with several launches of one kernel:
  ncu -i x.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_launch_table.py raw.csv out.json <systems> "<source / command>" ["<note>"]
Writes, per launch, duration, DRAM bytes read / written, tensor-pipe and issue-slot activity, and
over all launches the time-weighted tensor-pipe activity and the DRAM bytes per system and launch
(what bench.py's roofline.traffic is scaled from)."""
import csv
import json
import sys

COLS = {
    "ms": ("gpu__time_duration.sum", None),
    "dram_read_bytes": ("dram__bytes_read.sum", "bytes"),
    "dram_write_bytes": ("dram__bytes_write.sum", "bytes"),
    "tensor_pipe_active_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", None),
    "issue_active_pct": ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", None),
    "warp_instructions": ("smsp__inst_executed.sum", None),
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else None


def main(src, dst, systems, source, note=""):
    rows = [r for r in csv.reader(open(src)) if r]
    # the header is the first row that names the duration metric (ncu prints ==PROF== lines before it)
    h = next(i for i, r in enumerate(rows) if "gpu__time_duration.sum" in r)
    hdr, units, data = rows[h], rows[h + 1], rows[h + 2:]
    ix = {k: hdr.index(m) for k, (m, _) in COLS.items() if m in hdr}
    kname = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    launches = []
    for n, r in enumerate(data):
        e = {"launch": n}
        if kname is not None:
            e["kernel"] = r[kname][:80]
        for k, i in ix.items():
            v = num(r[i])
            if v is not None:
                v *= SCALE.get(units[i], 1.0)
            e[k] = v
        launches.append(e)
    total_ms = sum(e["ms"] for e in launches)
    out = {"source": source, "systems": systems, "launches": launches, "total_ms": total_ms}
    if all(e.get("tensor_pipe_active_pct") is not None for e in launches):
        out["tensor_pipe_active_pct_time_weighted"] = sum(e["tensor_pipe_active_pct"] * e["ms"] for e in launches) / total_ms
    if all(e.get("dram_read_bytes") is not None and e.get("dram_write_bytes") is not None for e in launches):
        out["dram_bytes_per_system_per_launch"] = sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in launches) / len(launches) / systems
    if note:
        out["note"] = note
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "launches"}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
