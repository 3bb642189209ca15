# round-2e final numbers with the separate BatchNorm-backward reduction as default: tests, smoke, BASELINE training configs,
# ncu launch list of one serialised iteration
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=${STEPS:-10}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/t_all.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    cb = d.get("cpu_baseline") or {}
    rf = d.get("roofline") or {}
    print(sys.argv[1].split("/")[-1], round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms | e2e", round(d["e2e"]["value"], 2),
          "| util exec", round((d.get("flops") or {}).get("tensor_util_executed", 0), 3), "| roof", rf.get("kernel"), round(rf.get("frac", 0), 3),
          "| clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "| cpu", cb.get("value"), cb.get("kind"), (d.get("extras") or {}).get("error"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 900 python bench.py --steps $S --warmup 3 > gpurun_out/bench_early.log 2>&1; echo "bench early (with cpu baseline) rc=$?"
tail -n 1 gpurun_out/bench_early.log > gpurun_out/r02f_bench_early_b22.json; summ gpurun_out/r02f_bench_early_b22.json
for w in mfnet-add mfnet-scoreadd multitask triple; do
  timeout 600 python bench.py --workload $w --steps $S --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"
  tail -n 1 gpurun_out/bench_$w.log > gpurun_out/r02f_bench_${w}_b22.json; summ gpurun_out/r02f_bench_${w}_b22.json
done
for b in 1 8; do
  timeout 300 python bench.py --batch $b --steps $S --warmup 3 --no-cpu-baseline > gpurun_out/bench_early_b$b.log 2>&1
  tail -n 1 gpurun_out/bench_early_b$b.log > gpurun_out/r02f_bench_early_b$b.json; summ gpurun_out/r02f_bench_early_b$b.json
done
if [ -z "$NO_NCU" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 22 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
fi
