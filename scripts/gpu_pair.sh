mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "conv" -x > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -30; tail -5 gpurun_out/t1_kernels.log; exit 1; fi
timeout 300 python scripts/bench_conv.py 2>&1 | grep -v "streamk=1" | tail -12
