#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep OUT.md [--traffic profiles/roofline_traffic.json --names 'sass substring=bench kernel name' ...]

Turns an `ncu --set full` report into the markdown summary kept under profiles/ and (optionally) records the measured
DRAM bytes per launch of named kernels in profiles/roofline_traffic.json, which bench.py reports as roofline.traffic."""
import csv
import io
import json
import subprocess
import sys

WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_path, names, title = None, {}, "ncu --set full --clock-control none summary"
    args = sys.argv[3:]
    while args:
        a = args.pop(0)
        if a == "--traffic":
            traffic_path = args.pop(0)
        elif a == "--title":
            title = args.pop(0)
        elif a == "--names":
            while args and not args[0].startswith("--"):
                k, v = args.pop(0).split("=")
                names[k] = v
    if rep.endswith(".csv"):  # the raw page exported on the GPU box (ncu -i REPORT --page raw --csv) when the report itself is too large to keep
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource report: `{rep}` (B200, sm_100a).\n\n")
        for d in data:
            kn = d[idx["Kernel Name"]]
            f.write(f"## {kn}\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {d[idx[w]]} | {units[idx[w]]} |\n")
            stalls = [(float(d[idx[k]]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                      for k in hdr if "issue_stalled" in k and "ratio" in k and "not_issued" not in k and d[idx[k]] not in ("", "n/a")]
            f.write("| top stalls (warps per issue) | " + ", ".join(f"{k}={v:.2f}" for v, k in sorted(stalls, reverse=True)[:6]) + " | |\n\n")
            rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            for sub, bench_name in names.items():
                if sub in kn and bench_name not in traffic:
                    traffic[bench_name] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "grid": d[idx["Grid Size"]], "report": rep}
    if traffic_path:
        try:
            old = json.load(open(traffic_path))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
