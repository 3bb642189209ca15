mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -40; exit 1; fi
timeout 1500 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench_b8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['tensor_util_of_step'], {k:(v['ms'],v['tflops']) for k,v in d['roofline']['families'].items()})"
