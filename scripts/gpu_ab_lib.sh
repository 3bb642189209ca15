mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider -k "conv" > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -40; exit 1; fi
timeout 900 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider -x > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
run() {
  env $1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 > gpurun_out/bench_$2.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$2.json'))
k=d['roofline']['kernels']
print('$2', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {n[10:42]:(v['ms'],v['tflops']) for n,v in k.items() if '<16>' in n or '<32>' in n})
"
}
run MCD_LIB_PATH=$PWD/multichannel-semseg-with-uda_b200/libmcd_sm100_prev.so prev
run MCD_X=1 new
run MCD_LIB_PATH=$PWD/multichannel-semseg-with-uda_b200/libmcd_sm100_prev.so prevb
run MCD_X=1 newb
