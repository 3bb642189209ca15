mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cfg in "1 1" "0 1" "1 0"; do
  set -- $cfg
  MCD_FUSE_BN_BWD=$1 MCD_CTA_PAIRS=$2 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse=$1 pairs=$2', d['value'], d['ms_per_step'], d['e2e']['value'], d['tensor_util_of_step'], d['clocks'])"
done
timeout 600 python scripts/trace_step.py 16 graph > gpurun_out/trace.log 2>&1; tail -2 gpurun_out/trace.log
