"""Turn what a gpurun call brought back (gpurun_out/) into the small tracked summaries under
profiles/:

  python tools/summarize_profiles.py launches gpurun_out/launches_bench.csv profiles/r01_launches_bench_summary.json "<command>"
  python tools/summarize_profiles.py ncu gpurun_out/prof_upd3.ncu-rep profiles/r01_ncu_update_tc.json "<label>" "<note>"

`launches` sums gpu__time_duration per kernel of an `ncu --metrics gpu__time_duration.sum` launch list;
`ncu` extracts the metrics the design document quotes from an `ncu --set full` report (needs `ncu` on PATH).
"""
import csv
import json
import subprocess
import sys

NCU_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utcomma_src_fp4_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum",
]


def launches(src, dst, command):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = {}
    for r in rows:
        name = r[k].split("(")[0][-60:]
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[v].replace(",", "")) / 1e6
    total = sum(t[1] for t in tot.values())
    out = {"command": command, "kernels": {n: {"launches": t[0], "total_ms": t[1], "share": t[1] / total}
                                            for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, open(dst, "w"), indent=1)
    for n, t in list(out["kernels"].items())[:6]:
        print("%-62s %4d launches %10.2f ms %5.1f %%" % (n, t["launches"], t["total_ms"], 100 * t["share"]))


def ncu(src, dst, label, note):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[2]
    got = {}
    for h, u, v in zip(hdr, units, row):
        if h in NCU_KEYS:
            got[h] = {"value": v, "unit": u}
    got["note"] = note
    json.dump({label: got}, open(dst, "w"), indent=1)
    for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
              "dram__bytes_write.sum"):
        print(k, got.get(k))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        ncu(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
