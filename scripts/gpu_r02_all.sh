# round-2 validation sweep: tests, smoke, one bench line per BASELINE config (run through gpurun)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
for w in early mfnet-add mfnet-scoreadd multitask triple; do
  timeout 600 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"
  tail -n 1 gpurun_out/bench_$w.log > gpurun_out/bench_r02_$w.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r02_$w.json"))
    print("$w", round(d["value"], 2), "pairs/s", round(d["ms_per_step"], 2), "ms | e2e", round(d["e2e"]["value"], 2), "| exec TF/pair",
          round(d["flops"]["executed_tflop_per_pair"], 3), "util", round(d["flops"]["tensor_util_executed"], 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["extras"].get("error"))
except Exception as e:
    print("$w FAILED", e)
    print(open("gpurun_out/bench_$w.log").read()[-1500:])
PY
done
timeout 900 python bench.py --workload infer --sweep --steps 5 --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1; echo "bench infer rc=$?"
tail -n 1 gpurun_out/bench_infer.log > gpurun_out/bench_r02_infer.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r02_infer.json"))
    for r in d["extras"]["sweep"]:
        print("infer B=%d %.1f img/s (e2e %.1f) %.3f ms util %.3f launches %d" % (r["batch_per_gpu"], r["images_per_s"], r["e2e_images_per_s"], r["ms_per_batch"], r["tensor_util"], r["gpu_launches_per_batch"]))
except Exception as e:
    print("infer FAILED", e)
    print(open("gpurun_out/bench_infer.log").read()[-1500:])
PY
