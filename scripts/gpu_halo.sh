mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -40; exit 1; fi
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_modules_gpu.py -q --timeout 900 -p no:cacheprovider -x > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
run() {
  env $1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $2 2>/dev/null | tail -n 1 > gpurun_out/bench_$3.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$3.json'))
k=d['roofline']['kernels']
print('$3', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {n[10:42]:v['ms'] for n,v in k.items()})
"
}
run MCD_HALO=0 22 halo0_b22
run MCD_HALO=1 22 halo1_b22
run MCD_HALO=1 16 halo1_b16
run MCD_HALO=1 30 halo1_b30
