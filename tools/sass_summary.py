"""Per-kernel counts of the SASS mnemonics that show which hardware paths the built library uses
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTCHMMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS =
cp.async, DMMA = FP64 mma.sync.   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "contractn_b200", "libctn_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "DMMA", "HMMA", "FFMA", "DFMA"]

proc = subprocess.Popen(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True)
counts = collections.OrderedDict()
cur = None
for line in proc.stdout:
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("ctn::", "")
        cur = counts.setdefault(name, collections.Counter())
        cur["instantiations"] += 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur[k] += 1
# group template instantiations by kernel base name
groups = collections.OrderedDict()
for name, c in counts.items():
    base = re.sub(r"<.*", "", name)
    g = groups.setdefault(base, collections.Counter())
    g.update(c)
    g["instantiations"] += 1
print(f"SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions summed over")
print("the template instantiations of each kernel\n")
print(f"{'kernel':42s} {'inst':>4s} {'SASS':>8s} " + " ".join(f"{k:>8s}" for k in KEYS))
for base, g in groups.items():
    print(f"{base[:42]:42s} {g['instantiations']:4d} {g['_total']:8d} " + " ".join(f"{g[k]:8d}" for k in KEYS))
