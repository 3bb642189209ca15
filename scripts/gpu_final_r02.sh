# round-2 final measurement sweep (run through gpurun): tests, smoke, every BASELINE config, batch set, inference sweep,
# reference arm, ncu launch list + full capture.  Outputs under gpurun_out/ (copied to profiles/ by hand).
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
    print(sys.argv[1].split("/")[-1], round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms | e2e", round(d["e2e"]["value"], 2),
          "| util exec", round((d.get("flops") or {}).get("tensor_util_executed", 0), 3), "| roof", round((d.get("roofline") or {}).get("frac", 0), 3),
          "| clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "| cpu", cb.get("value"), cb.get("kind"), (d.get("extras") or {}).get("error"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 900 python bench.py --steps $S --warmup 3 > gpurun_out/bench_early.log 2>&1; echo "bench early (with cpu baseline) rc=$?"
tail -n 1 gpurun_out/bench_early.log > gpurun_out/r02_bench_early_b22.json; summ gpurun_out/r02_bench_early_b22.json
for w in mfnet-add mfnet-scoreadd multitask triple; do
  timeout 600 python bench.py --workload $w --steps $S --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"
  tail -n 1 gpurun_out/bench_$w.log > gpurun_out/r02_bench_${w}_b22.json; summ gpurun_out/r02_bench_${w}_b22.json
done
for b in 1 2 4 8; do
  timeout 300 python bench.py --batch $b --steps $S --warmup 3 --no-cpu-baseline > gpurun_out/bench_early_b$b.log 2>&1
  tail -n 1 gpurun_out/bench_early_b$b.log > gpurun_out/r02_bench_early_b$b.json; summ gpurun_out/r02_bench_early_b$b.json
done
timeout 300 python bench.py --input u8 --steps $S --warmup 3 --no-cpu-baseline > gpurun_out/bench_early_u8.log 2>&1
tail -n 1 gpurun_out/bench_early_u8.log > gpurun_out/r02_bench_early_u8_b22.json; summ gpurun_out/r02_bench_early_u8_b22.json
timeout 900 python bench.py --workload infer --sweep --steps 5 > gpurun_out/bench_infer.log 2>&1; echo "bench infer rc=$?"
tail -n 1 gpurun_out/bench_infer.log > gpurun_out/r02_bench_infer_sweep.json; summ gpurun_out/r02_bench_infer_sweep.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_infer_sweep.json"))
    for r in d["extras"]["sweep"]:
        print("infer B=%d %.1f img/s (e2e %.1f) %.3f ms util %.3f launches %d" % (r["batch_per_gpu"], r["images_per_s"], r["e2e_images_per_s"], r["ms_per_batch"], r["tensor_util"], r["gpu_launches_per_batch"]))
except Exception as e:
    print("infer FAILED", e)
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference arm rc=$?"
tail -n 1 gpurun_out/bench_reference.log > gpurun_out/r02_bench_reference_cpu.json; cut -c1-600 gpurun_out/r02_bench_reference_cpu.json
if [ -z "$NO_NCU" ]; then
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 22 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/prof_r02 python scripts/profile_r02.py 22 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la /tmp/prof_r02.ncu-rep
ncu -i /tmp/prof_r02.ncu-rep --page raw --csv > gpurun_out/prof_r02_raw.csv 2>/dev/null
if [ $(stat -c %s /tmp/prof_r02.ncu-rep) -lt 40000000 ]; then cp /tmp/prof_r02.ncu-rep gpurun_out/; fi
fi
du -sh gpurun_out
