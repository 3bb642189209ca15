# weak-scaling check on ONE box: N = 1 and N = NGPU back to back (same box, same clocks), early-fusion workload
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${NGPU:-2}
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n1_of$N.log 2>&1; echo "bench n1 rc=$?"
tail -n 1 gpurun_out/bench_scale_n1_of$N.log > gpurun_out/r02_bench_scale_n1_of$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_scale_n$N.log 2>&1; echo "bench n$N rc=$?"
tail -n 1 gpurun_out/bench_scale_n$N.log > gpurun_out/r02_bench_scale_n$N.json
python - <<PY
import json
a = json.load(open("gpurun_out/r02_bench_scale_n1_of$N.json")); b = json.load(open("gpurun_out/r02_bench_scale_n$N.json"))
print("N=1 %.2f pairs/s %.2f ms | N=%d %.2f pairs/s %.2f ms | efficiency %.4f | e2e %.2f | clocks %s %s" % (
    a["value"], a["ms_per_step"], b["n_gpus"], b["value"], b["ms_per_step"], b["value"] / (b["n_gpus"] * a["value"]),
    b["e2e"]["value"], a["clocks"]["sm_mhz"], b["clocks"]["sm_mhz"]))
PY
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_steps_gpu.py -q --timeout 500 -p no:cacheprovider -k two_gpu > gpurun_out/t_two_gpu.log 2>&1; echo "two-gpu parity rc=$?"; grep -E "passed|failed|skipped" gpurun_out/t_two_gpu.log | tail -n 2
fi
