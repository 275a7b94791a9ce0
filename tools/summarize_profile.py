"""Turns the ncu outputs brought back in gpurun_out/ into the text summaries kept under profiles/.

    python tools/summarize_profile.py <tag>      # e.g. r1_knn_phases
"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
out = open("profiles/%s_summary.txt" % tag, "w")


def p(*a):
    print(*a, file=out)
    print(*a)


# ---- launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)
rows = list(csv.reader(open("gpurun_out/launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[ik].split("(")[0][-70:]
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
import os
p("== launch list of `python bench.py --steps %s --warmup 3 --no-extra --no-cpu --no-e2e` (ns, share of all launches incl. the one-off build)" % os.environ.get("PROF_STEPS", "4"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    p("%-72s n=%4d total=%12.0f share=%5.1f%% avg=%10.0f" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
loop = {k: a for k, a in agg.items() if any(s in k for s in ("knn_kernel", "knn_overflow", "accumulate", "finalize", "hist_kernel", "pick_kernel", "init_limits"))}
lt = sum(a[1] for a in loop.values())
p("-- share inside the iteration loop only")
for k, a in sorted(loop.items(), key=lambda x: -x[1][1]):
    p("%-72s share=%5.1f%%" % (k, 100 * a[1] / lt))

# ---- full capture of the dominant kernel
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_knn.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    p("== ncu --set full: %s grid %s block %s" % (d.get("Kernel Name", "")[:70], d.get("Grid Size"), d.get("Block Size")))
    for w in want:
        if w in d:
            p("   %-82s %s %s" % (w, d[w], units[hdr.index(w)]))
out.close()
