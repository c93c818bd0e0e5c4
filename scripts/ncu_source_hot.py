"""Aggregate `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` (SASS rows) into: total warp instructions, warp-stall
samples per opcode, and the hottest instructions with their dominant stall reason.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --kernel-name regex:attn_tc_fwd_kernel | python scripts/ncu_source_hot.py [top]
"""
import csv
import sys
from collections import defaultdict

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(ln for ln in sys.stdin if not ln.startswith("==")))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = defaultdict(lambda: [0, 0])
insts = []
tot_s = tot_i = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name"):
        continue
    sass = r[col["Source"]].strip()
    toks = sass.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("MUFU", "UTCHMMA", "LDTM", "SYNCS", "STS", "LDS", "STG", "LDG", "BAR") else "")
    s = int(float(r[col["# Samples"]] or 0))
    n = int(float(r[col["Instructions Executed"]] or 0))
    by_op[op][0] += s
    by_op[op][1] += n
    tot_s += s
    tot_i += n
    reasons = sorted(((int(float(r[col[c]] or 0)), c) for c in stall_cols), reverse=True)[:2]
    insts.append((s, n, sass, reasons))
print("total warp-instructions %d, stall samples %d" % (tot_i, tot_s))
print("%-14s %9s %7s %12s %7s" % ("opcode", "samples", "share", "warp-insts", "share"))
for op, (s, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s %9d %6.1f%% %12d %6.1f%%" % (op, s, 100.0 * s / max(tot_s, 1), n, 100.0 * n / max(tot_i, 1)))
print("\nhottest instructions")
for s, n, sass, reasons in sorted(insts, key=lambda t: -t[0])[:top]:
    print("%6d %5.1f%%  x%-8d %-70s %s" % (s, 100.0 * s / max(tot_s, 1), n, sass[:70], " ".join("%s=%d" % (c[6:], v) for v, c in reasons if v)))
