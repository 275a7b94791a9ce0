"""Turns the ncu reports brought back in gpurun_out/ into text summaries under profiles/.

    python tools/summarize_profile2.py <report.ncu-rep> [...]   -> profiles/<report>_summary.txt
    python tools/summarize_profile2.py --launches gpurun_out/x.csv -> profiles/<x>_summary.txt
"""
import collections
import csv
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[ik].split("(")[0][-80:]
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["== launch list %s (gpu__time_duration.sum ns; cold-cache, serialised: compare shares)" % os.path.basename(path)]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        out.append("%-82s n=%4d total=%12.0f share=%5.1f%% avg=%10.0f" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    return out


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("== ncu --set full: %s grid %s block %s" % (d.get("Kernel Name", "")[:110], d.get("Grid Size"), d.get("Block Size")))
        for w in WANT:
            if w in d:
                out.append("   %-82s %s %s" % (w, d[w], units[hdr.index(w)]))
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--launches":
        for p in args[1:]:
            lines = launches(p)
            name = os.path.splitext(os.path.basename(p))[0]
            open("profiles/%s_summary.txt" % name, "w").write("\n".join(lines) + "\n")
            print("\n".join(lines))
    else:
        for p in args:
            lines = report(p)
            name = os.path.splitext(os.path.basename(p))[0]
            open("profiles/%s_summary.txt" % name, "w").write("\n".join(lines) + "\n")
            print("\n".join(lines))
