#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) of the land kernels as markdown: launch metrics, stall reasons, opcode mix
and the hottest source lines.  Used to write profiles/*.md from gpurun_out/*.ncu-rep (the raw reports are not committed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [N hottest lines] > profiles/rX_ncu_summary.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__block_size", "launch__grid_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s\n" % rep.split("/")[-1])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## launch %s: `%s`\n\n| metric | value | unit |\n|---|---|---|" % (d["ID"], d.get("Kernel Name")))
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        print()
    src = ncu(rep, "--page", "source", "--csv", "--print-source=cuda,sass")
    cur, agg, st, op = None, defaultdict(lambda: [0, 0, ""]), defaultdict(int), defaultdict(int)
    iI = iS = None
    for r in csv.reader(src.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if len(r) > 10 and r[0].isdigit():
            try:
                ie, smp = int(r[iI]), int(r[iS])
            except ValueError:
                continue
            a = agg[(cur, int(r[0]))]
            a[0] += ie
            a[1] += smp
            a[2] = r[1].strip()[:100]
        elif len(r) > 10 and r[2].startswith("0x"):
            try:
                ie = int(r[iI])
            except ValueError:
                continue
            m = r[3].split()
            op[(m[1] if m[0].startswith("@") else m[0]).split(".")[0]] += ie
            for k in hdr:
                if k.startswith("stall_") and "Not Issued" not in k:
                    try:
                        st[k] += int(r[hdr.index(k)])
                    except ValueError:
                        pass
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    tt = sum(st.values()) or 1
    to = sum(op.values()) or 1
    print("## all captured launches\n")
    print("warp instructions executed: %d\n" % tot)
    print("stall reasons (share of samples): " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / tt)
                                                          for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]) + "\n")
    print("opcode mix (share of executed warp instructions): " + ", ".join("%s %.1f%%" % (k, 100 * v / to)
                                                                           for k, v in sorted(op.items(), key=lambda x: -x[1])[:24]) + "\n")
    print("hottest source lines (share of instructions / of stall samples):\n\n```")
    for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("%5.2f%% %5.2f%%  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / ts, f, l, a[2]))
    print("```")


if __name__ == "__main__":
    main()
