"""Per-kernel counts of the Blackwell-native SASS mnemonics in maskedsst_b200/libmsst.so (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAPF = TMA load / store / L2 prefetch, HMMA = legacy mma.sync.
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "maskedsst_b200", "libmsst.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2"]
cur, tab = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("msst::", "")
        name = re.sub(r"\(.*", "", name)
        cur = tab.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k), ln):
            cur[k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (sm_100a)")
print(f"{'kernel':58s}" + "".join(f"{k:>9s}" for k in keys))
for name, c in tab.items():
    if any(c[k] for k in keys[:6]) or c["HMMA"]:
        print(f"{name[:58]:58s}" + "".join(f"{c[k]:9d}" for k in keys))
