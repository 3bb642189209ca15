mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/smoke.log; tail -n 5 gpurun_out/bench.log | cut -c1-1500
grep -E "passed|failed" gpurun_out/t2_parity.log
