"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel -> profiles/*.csv

    python scripts/ncu_launch_list.py gpurun_out/launches.csv profiles/r01b_launch_list_step_b16.csv "note"
"""
import collections
import csv
import re
import sys

src, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
lines = [ln for ln in open(src) if ln.startswith('"')]
rows = list(csv.DictReader(lines))


def kernel_short_name(n):
    """ncu's demangled name -> the name bench.py / DESIGN.md use."""
    n = re.sub(r"^void ", "", n)
    n = n.replace("mcd::", "")
    n = re.sub(r"\(bool\)", "", n)
    m = re.match(r"conv_umma_fprop_kernel<(\d+), *(\w+), *(\d+), *(\w+)(?:, *(\w+))?>", n)
    if m:
        bn, pair, occ, halo = m.group(1), m.group(2) in ("1", "true"), m.group(3), m.group(4) in ("1", "true")
        inst = "" if m.group(5) is None else (" [dgrad-epilogue inst.]" if m.group(5) in ("1", "true") else " [forward inst.]")
        return "conv_umma_fprop_kernel<%s%s%s>%s" % (bn, ",pair" if pair else "", ",halo" if halo else "", inst)
    m = re.match(r"conv_umma_wgrad_kernel<(\d+), *(\d+)>", n)
    if m:
        return "conv_umma_wgrad_kernel<%s>" % m.group(1)
    m = re.match(r"(conv_umma_\w+)<(\d+)>", n)
    if m:
        return "%s<%s>" % (m.group(1), m.group(2))
    return re.sub(r"\(.*$", "", n)[:70]


short = kernel_short_name


agg = collections.OrderedDict()
total = 0.0
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    us = float(r["Metric Value"].replace(",", "")) / (1e3 if r["Metric Unit"] == "ns" else 1.0)
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
n = sum(a[0] for a in agg.values())
with open(out, "w") as f:
    for ln in note.split("\\n"):
        if ln:
            f.write("# " + ln + "\n")
    f.write("# total %.0f us over %d launches\n" % (total, n))
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.0f,%.2f,%.1f\n" % (k.replace(",", ";"), c, t, 100.0 * t / total, t / c))
print("wrote", out, "total_us", round(total), "launches", n)
