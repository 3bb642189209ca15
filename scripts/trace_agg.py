"""aggregate gpurun_out/trace_kernels.csv by kernel name (count, total us) and print concurrency summary."""
import csv, collections, sys
rows = list(csv.DictReader(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_kernels.csv")))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["name"][:52]; agg[k][0] += 1; agg[k][1] += float(r["dur_us"])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-54s %4d %9.1f us" % (k, v[0], v[1]))
ev = []
for r in rows:
    t0 = float(r["ts_us"]); ev.append((t0, 1)); ev.append((t0 + float(r["dur_us"]), -1))
ev.sort()
busy = conc = 0.0; n = 0; last = ev[0][0]
for t, d in ev:
    if n >= 1: busy += t - last
    if n >= 2: conc += t - last
    n += d; last = t
print("span %.1f ms busy %.1f ms >=2 concurrent %.1f ms" % ((ev[-1][0] - ev[0][0]) / 1e3, busy / 1e3, conc / 1e3))
