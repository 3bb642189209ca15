"""Summarise an `ncu --set full` report (read with `ncu -i rep --page raw --csv`) into a compact CSV under
profiles/ and refresh profiles/ncu_traffic.json (DRAM bytes per launch per kernel, read by bench.py's
roofline.traffic).

    python scripts/ncu_summarise.py gpurun_out/prof_conv.ncu-rep profiles/r01b_ncu_full_conv_bn.csv
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_dim_x", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg"]


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


short_name = kernel_short_name


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):      # already exported on the GPU box: ncu -i rep --page raw --csv
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    traffic = {}
    with open(out, "w") as f:
        if note:
            f.write("# " + note + "\n")
        f.write(",".join("%s [%s]" % (c, units[idx[c]]) for c in cols) + "\n")
        for r in data:
            vals = [r[idx[c]].replace(",", ";") for c in cols]
            kname = short_name(r[idx["Kernel Name"]])[:60]
            vals[0] = kname.replace(",", ";")
            f.write(",".join(vals) + "\n")
            try:
                def tobytes(c):
                    v, u = float(r[idx[c]].replace(",", "")), units[idx[c]].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                t = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
                ent = traffic.setdefault(kname, {"dram_bytes_per_launch": [], "grid": []})   # per instantiation
                ent["dram_bytes_per_launch"].append(t)
                ent["grid"].append(r[idx["Grid Size"]])
            except (KeyError, ValueError):
                pass
    tpath = os.path.join(os.path.dirname(out), "ncu_traffic.json")
    prev = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for k, v in traffic.items():
        prev[k] = {"dram_bytes_per_launch": v["dram_bytes_per_launch"], "grid": v["grid"],
                   "source": os.path.basename(out)}
    json.dump(prev, open(tpath, "w"), indent=1, sort_keys=True)
    print("wrote", out, "and", tpath, "(%d kernels)" % len(data))


if __name__ == "__main__":
    main()
