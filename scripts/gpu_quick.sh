mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log > gpurun_out/bench_head.json; python -c "
import json; d=json.load(open('gpurun_out/bench_head.json')); print(round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], round(d['tensor_util_of_step'],4))"
