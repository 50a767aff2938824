#!/usr/bin/env python3
"""Turns an ncu launch list (csv) and/or an `ncu --set full` report into the markdown summaries
committed under profiles/.  Usage: summarize_ncu.py <tag> [launches.csv] [full.ncu-rep]"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | mean ns | share of device time |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append("| %s | %d | %.0f | %.3f |" % (k, len(v), sum(v) / len(v), sum(v) / tot))
    return "\n".join(out)


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append("### %s  grid %s block %s" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        out.append("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in hdr:
                j = hdr.index(key)
                out.append("| %s (`%s`) | %s | %s |" % (label, key, r[j], units[j]))
        stalls = []
        for j, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[j].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("")
        out.append("warp stall cycles per issued instruction (top): " + ", ".join("%s %.2f" % (k, v) for v, k in stalls[:7]))
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    tag = sys.argv[1]
    print("# ncu summary %s\n" % tag)
    for a in [a for a in sys.argv[2:] if a]:
        if a.endswith(".csv"):
            print("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
            print(launches(a) + "\n")
        else:
            print("## `ncu --set full --clock-control none` (one launch per kernel)\n")
            print(full(a))
