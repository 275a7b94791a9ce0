"""Per-source-line cost of one captured kernel from an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/hot_lines.py gpurun_out/prof_knn.ncu-rep [top] [launch-index]

Prints the lines with the most executed warp instructions, their share, the average active lanes
and the stall samples — the view used to decide what to change in K2.
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr, launch, seen = "?", None, 0, set()
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        if fname in seen:  # sections repeat per captured launch
            launch, seen = launch + 1, set()
        seen.add(fname)
        continue
    if r[0] == "Function Name":
        if launch == which and len(seen) == 1:
            print("kernel:", r[1][:100])
        continue
    if r[0] == "Line No":
        hdr = r
        ii, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or launch != which or r[0] == "":
        continue
    try:
        key = (fname, int(r[0]))
        inst, tinst, samp = float(r[ii]), float(r[it]), float(r[isamp])
    except ValueError:
        continue
    agg[key] = [r[1].strip()[:90], inst, tinst, samp]
tot_i = sum(a[1] for a in agg.values())
tot_s = sum(a[3] for a in agg.values())
print("total warp instructions %.0f, lanes/inst %.2f, samples %.0f" % (tot_i, sum(a[2] for a in agg.values()) / max(tot_i, 1), tot_s))
print("%-12s %5s %6s %6s %5s  %s" % ("file", "line", "inst%", "samp%", "lanes", "source"))
by = 3 if "--by-samples" in sys.argv else 1
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][by])[:top]:
    print("%-12s %5d %6.2f %6.2f %5.1f  %s" % (f[:12], ln, 100 * a[1] / tot_i, 100 * a[3] / max(tot_s, 1), a[2] / max(a[1], 1), a[0]))
