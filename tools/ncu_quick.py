#!/usr/bin/env python
"""tools/ncu_quick.py REPORT.ncu-rep [top] -- per kernel of an `ncu --set full --import-source on` report: the headline metrics, the
warp-stall breakdown per issued instruction and the `top` instructions holding the most stall samples (what to fix first)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def page(rep, name):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout.splitlines()))


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10
    raw = page(rep, "raw")
    hdr = raw[0]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        print("==", d["Kernel Name"][:90])
        for k in KEYS:
            if k in d:
                print("    %-72s %s" % (k, d[k]))
        st = []
        for k, v in d.items():
            if "issue_stalled" in k and "per_issue_active" in k:
                try:
                    if float(v) > 0.2:
                        st.append((round(float(v), 2), k.split("stalled_")[-1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("    stalls per issue:", sorted(st, reverse=True)[:8])
    src = page(rep, "source")
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if b["name"] in seen or not b["rows"]:
            continue
        seen.add(b["name"])
        h = b["rows"][0]
        ix = {x: i for i, x in enumerate(h)}
        data = [r for r in b["rows"][1:] if len(r) > ix["# Samples"] and r[ix["# Samples"]].isdigit()]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        print("== hot instructions of", b["name"][:70], "(samples %d)" % tot)
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
            print("    %5.1f%%  %s" % (100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Source"]][:100]))


if __name__ == "__main__":
    main()
