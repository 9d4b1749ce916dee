"""profiles/rNN_sass_summary.txt: per-kernel counts of the SASS mnemonics that evidence the hardware paths
(tcgen05 = UTC*MMA / LDTM / UTCBAR, TMA = UTMALDG, NVLink multicast = LDGMC / ST*MC, legacy tensor core = HMMA ...)
from `cuobjdump -sass gdn_b200/libgdn_b200.so`.     python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gdn_b200", "libgdn_b200.so")
PATS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "LDGMC",
        "MULTIMEM", "REDG", "RED", "ATOMG", "ATOMS", "HMMA", "MATCH", "LDGSTS", "LDSM", "DFMA", "FFMA2", "FFMA", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
demangle = dict(zip(re.findall(r"Function : (\S+)", sass), names))
fn, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
    if fn is None or not mm:
        continue
    op = mm.group(1)
    total[fn] += 1
    for p in PATS:
        if op == p or op.startswith(p + "."):
            counts[fn][op if p in ("LDGMC", "UTCHMMA", "UTMALDG", "LDTM", "HMMA", "REDG", "RED") else p] += 1
            break
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)\n# kernel: total SASS instructions; selected mnemonics\n")
for f in sorted(total, key=lambda k: demangle.get(k, k)):
    short = re.sub(r"\(.*", "", demangle.get(f, f)).replace("void ", "").replace("gdn::", "")
    sel = "  ".join(f"{k}x{v}" for k, v in sorted(counts[f].items()))
    print(f"{short:60s} {total[f]:6d}   {sel}")
agg = collections.Counter()
for f in counts:
    agg.update(counts[f])
print("\n# whole library:", "  ".join(f"{k}x{v}" for k, v in sorted(agg.items())))
