mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
./build_exp/exp_swizzle_shift > gpurun_out/exp_swizzle_shift.txt 2>&1; echo "exp rc=$?"; cat gpurun_out/exp_swizzle_shift.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -30; fi
run() {
  env $1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $2 2>/dev/null | tail -n 1 > gpurun_out/bench_$3.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$3.json'))
k=d['roofline']['kernels']
print('$3', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {n[10:40]:v['ms'] for n,v in k.items()})
"
}
run MCD_THIN_OCC2=0 16 occ1_b16
run MCD_THIN_OCC2=1 16 occ2_b16
run MCD_THIN_OCC2=1 18 occ2_b18
run MCD_THIN_OCC2=1 22 occ2_b22
run MCD_THIN_OCC2=0 22 occ1_b22
