mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t1_kernels.log 2>&1; echo "kernels rc=$?" > gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_modules_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/t1_kernels.log
tail -n 30 gpurun_out/t2_parity.log
