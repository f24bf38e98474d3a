#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` report: headline metrics, stall mix and the
SASS instructions that collect the most warp-stall samples.

    python profiles/tools/ncu_top.py REPORT.ncu-rep [N_TOP]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[hdr.index("Kernel Name")])
    stalls = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            print(f"  {h} = {v} {u}")
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            stalls[h.split("stalled_")[1]] = int(v)
    tot = sum(stalls.values()) or 1
    print("  stall mix:", ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(stalls.items(), key=lambda t: -t[1])[:8]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"  {len(data)} SASS instructions, {tot} samples; top {ntop}:")
    top = sorted(enumerate(data), key=lambda t: -int(t[1][ix["# Samples"]]))[:ntop]
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for i, r in sorted(top):
        st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]])}
        print(f"   {i:5d} {int(r[ix['# Samples']]):5d} x{r[ix['Instructions Executed']]:>7} {r[ix['Source']].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main()
