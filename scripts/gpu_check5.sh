mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t1_kernels.log 2>&1; echo "kernels rc=$?" > gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed" gpurun_out/t1_kernels.log gpurun_out/t2_parity.log
tail -n 1 gpurun_out/bench.log | cut -c1-2500
