"""CPU: per-kernel SASS instruction counts of the shipped library (no GPU needed).

    python scripts/sass_table.py [multimae_b200/libmultimae_b200.so] > profiles/r02_sass_instruction_table.txt

For every kernel in the .so (cuobjdump -sass): total instructions and the counts of the mnemonics that tell a
Blackwell-native kernel from a recompiled sm_80-style one (B200_PROFILING.md):
  UTCHMMA / UTCQMMA (tcgen05.mma, `.2CTA` = cta_group::2), LDTM / STTM (tcgen05.ld/st, TMEM), UTMALDG / UTMASTG / UTMAREDG
  (TMA tile load / store / reduce-add), UBLKCP (bulk copy), SYNCS (mbarrier) -- versus HMMA (mma.sync), LDSM (ldmatrix),
  LDGSTS (cp.async)."""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimae_b200", "libmultimae_b200.so")
MNEMS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "LDSM",
         "LDGSTS", "MUFU"]

sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
it = iter(names)
rows = []
cur = None
op_re = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in sass.split("\n"):
    if "Function :" in line:
        full = next(it)
        short = re.sub(r"\(anonymous namespace\)::", "", full)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\(.*$", "", short)          # drop the argument list
        short = short.replace("mmae::", "")
        cur = [short, 0, collections.Counter()]
        rows.append(cur)
        continue
    m = op_re.match(line)
    if m and cur is not None:
        op = m.group(1)
        cur[1] += 1
        base = op.split(".")[0]
        cur[2][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            cur[2]["UTCHMMA.2CTA"] += 1

try:
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except Exception:
    commit = "?"
print("# SASS instruction counts per kernel of %s (sm_100a), source tree at commit %s" % (os.path.basename(so), commit))
print("# cuobjdump -sass | per-function mnemonic counts; tcgen05 = UTCHMMA(+.2CTA)/LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAREDG,")
print("# sm_80 idiom = HMMA (mma.sync) / LDSM (ldmatrix) / LDGSTS (cp.async)")
hdr = "%-72s %7s " % ("kernel", "instr") + " ".join("%7s" % (m if len(m) <= 7 else m[-7:]) for m in MNEMS)
print(hdr)
tot = collections.Counter()
rows.sort(key=lambda r: r[0])
for short, n, c in rows:
    print("%-72s %7d " % (short[:72], n) + " ".join("%7d" % c[m] for m in MNEMS))
    for m in MNEMS:
        tot[m] += c[m]
print("%-72s %7d " % ("TOTAL (%d kernels)" % len(rows), sum(r[1] for r in rows)) + " ".join("%7d" % tot[m] for m in MNEMS))
hm = [r[0] for r in rows if r[2]["HMMA"]]
tc = [r[0] for r in rows if r[2]["UTCHMMA"]]
print("\n# kernels issuing tcgen05.mma (UTCHMMA): %d" % len(tc))
print("# kernels issuing mma.sync (HMMA): %d" % len(hm))
for k in hm:
    print("#   " + k)
