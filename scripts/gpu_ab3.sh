mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider -k "bn or batchnorm or BatchNorm" > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -40; exit 1; fi
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_modules_gpu.py -q --timeout 900 -p no:cacheprovider -x > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
run() {
  env $1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $2 2>/dev/null | tail -n 1 > gpurun_out/bench_$3.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$3.json'))
print('$3', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])
"
}
run MCD_SGD_BLOCKS=64 22 sgd64_b22
run MCD_SGD_BLOCKS=192 22 sgd192_b22
run MCD_SGD_BLOCKS=192 37 sgd192_b37
timeout 600 python scripts/trace_step.py 22 graph > gpurun_out/trace.log 2>&1; tail -1 gpurun_out/trace.log
python scripts/trace_agg.py gpurun_out/trace_kernels.csv 40
