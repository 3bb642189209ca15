mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t1_kernels.log 2>&1; echo "kernels rc=$?"; grep -E "passed|failed" gpurun_out/t1_kernels.log
timeout 1200 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log | cut -c1-3000
timeout 600 python scripts/trace_step.py 16 graph > gpurun_out/trace.log 2>&1; tail -2 gpurun_out/trace.log
python scripts/trace_agg.py gpurun_out/trace_kernels.csv 45
