"""Per-kernel counts of the Blackwell-native SASS mnemonics in the in-tree library (evidence that the hot path is tcgen05 / TMEM / TMA
code, not a recompiled mma.sync kernel).  usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nerf_sampling_b200", "libb200nerf.so")
PATTERNS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "STTM", "UTMALDG.2D.2CTA", "UTMALDG", "UBLKCP", "UTCATOMSWS",
            "SYNCS", "HMMA", "IMMA", "FFMA", "MUFU"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
usage = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
res = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", usage):
    res[m.group(1)] = m.group(2)
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    ins = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not ins:
        continue
    op = ins.group(1)
    counts[cur]["_total"] += 1
    for p in PATTERNS:
        if op == p or op.startswith(p + "."):
            counts[cur][p] += 1
            break
print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("# UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA/IMMA = legacy mma.sync")
print()
cols = [p for p in PATTERNS if any(counts[f][p] for f in order)]
print("| kernel | instr | " + " | ".join(cols) + " | regs / smem |")
print("|---|---|" + "---|" * (len(cols) + 1))
for f in order:
    c = counts[f]
    name = demangle(f)
    name = re.sub(r"\(.*", "", name)[:70]
    r = res.get(f, "")
    rs = " ".join(re.findall(r"(REG:\d+|SHARED:\d+)", r))
    print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[p]) if c[p] else "" for p in cols) + f" | {rs} |")
tot = collections.Counter()
for f in order:
    tot.update(counts[f])
print()
print("totals: " + ", ".join(f"{p} {tot[p]}" for p in PATTERNS if tot[p]))
assert tot["HMMA"] == 0 and tot["IMMA"] == 0, "legacy tensor-core instructions found"
