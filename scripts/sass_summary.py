"""Per-kernel SASS evidence for the tcgen05 / TMA / bulk-copy paths: counts of the relevant mnemonics in every kernel
of libmcd_sm100.so (cuobjdump -sass, no GPU needed).   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multichannel-semseg-with-uda_b200", "libmcd_sm100.so")
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMAPF", "HMMA",
        "ATOMS", "RED", "ATOMG", "SHFL", "MUFU", "LDGSTS", "STL", "LDL", "R2UR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        if op in WANT:
            counts[kern][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of libmcd_sm100.so (sm_100a): UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit,")
print("# LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier")
print("# ops, ATOMS / RED = shared / global atomics, STL / LDL = local-memory (spill) stores / loads")
for k, name in zip(counts, demangle):
    name = re.sub(r"\(.*$", "", name.replace("mcd::", "").replace("void ", ""))[:90]
    c = counts[k]
    print("%-92s %6d instr | %s" % (name, total[k], " ".join("%s %d" % (o, c[o]) for o in WANT if c[o])))
