mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$tag.log 2>&1
  echo "$tag $(tail -n 1 gpurun_out/ab_$tag.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.1f pairs/s %.2f ms %.0f MHz" % (d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))')"; }
run default A=0
run thin_occ2_dgrad MCD_THIN_OCC2_DGRAD=1
run bn_bulk0 MCD_BN_BULK=0
run default_again A=0
