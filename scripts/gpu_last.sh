bash scripts/gpu_validate.sh
run() {
  env $1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 > gpurun_out/bench_$2.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$2.json'))
print('$2', round(d['value'],2), round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])
"
}
run MCD_DIFF2D_BWD_REG=0 d2d0
run MCD_DIFF2D_BWD_REG=1 d2d1
run MCD_DIFF2D_BWD_REG=0 d2d0b
run MCD_DIFF2D_BWD_REG=1 d2d1b
grep -E "diff2d" gpurun_out/launches.csv | head -3 | cut -c1-60
