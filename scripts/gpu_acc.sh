mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=$PWD/multichannel-semseg-with-uda_b200
for V in acc32 acc64; do
MCD_LIB_PATH=$L/libmcd_sm100_$V.so timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider -k "conv" > gpurun_out/t1_$V.log 2>&1; echo "kernels $V rc=$?"; grep -E "passed|failed" gpurun_out/t1_$V.log
done
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider -x > gpurun_out/t2_parity.log 2>&1; echo "default kernels+parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
MCD_LIB_PATH=$L/libmcd_sm100_acc32.so timeout 900 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider -x > gpurun_out/t2_parity32.log 2>&1; echo "acc32 parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity32.log | head
run() {
  env $1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 > gpurun_out/bench_$2.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$2.json'))
k=d['roofline']['kernels']
print('$2', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], {n[10:42]:v['ms'] for n,v in k.items() if ('<16>' in n or '<32>' in n or '<64' in n) and 'wgrad' not in n})
"
}
run MCD_X=1 base
run MCD_LIB_PATH=$L/libmcd_sm100_acc32.so acc32
run MCD_LIB_PATH=$L/libmcd_sm100_acc64.so acc64
run MCD_X=1 baseb
run MCD_LIB_PATH=$L/libmcd_sm100_acc32.so acc32b
run MCD_LIB_PATH=$L/libmcd_sm100_acc64.so acc64b
