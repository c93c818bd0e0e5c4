"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel for the LAST step.

    python scripts/summarize_launches.py gpurun_out/launches.csv [--steps-in-log 4] > profiles/rNN_launch_summary.txt
"""
import argparse
import csv
import collections
import re
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--steps-in-log", type=int, default=4, help="how many identical steps the log holds (the last one is used)")
a = ap.parse_args()

rows = []
with open(a.csv, newline="") as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
reader = csv.DictReader(lines)
for r in reader:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
    name = r["Kernel Name"]
    rows.append((name, ns))
if not rows:
    sys.exit("no gpu__time_duration rows found")
# step boundaries: the mask sampler runs exactly once per step, first thing in the forward; use the last COMPLETE step
marks = [i for i, (name, _) in enumerate(rows) if "mask_sampler_kernel" in name]
if len(marks) >= 2:
    last = rows[marks[-2]:marks[-1]]
    per_step = len(last)
else:
    per_step = len(rows) // a.steps_in_log
    last = rows[-per_step:]
tot = sum(ns for _, ns in last)
agg = collections.OrderedDict()
for name, ns in last:
    clean = name.replace("(anonymous namespace)::", "")
    m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)(<[^()]*>)?\(", clean)
    short = (m.group(1) + (m.group(2) or "")) if m else clean[:60]
    short = short.replace("(bool)", "")
    k = agg.setdefault(short, [0, 0.0])
    k[0] += 1
    k[1] += ns
print("launches in log: %d; per step: %d; last-step total kernel time: %.3f ms (ncu per-launch times are cold-cache and"
      " serialised: compare SHARES)" % (len(rows), per_step, tot / 1e6))
print("%-62s %8s %10s %7s" % ("kernel", "launches", "ms", "share"))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %8d %10.3f %6.1f%%" % (name[:62], n, ns / 1e6, 100 * ns / tot))
