mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for c in 0 128 256 512 100000; do
  MCD_FUSE_BN_BWD=1 MCD_FUSE_BN_BWD_MIN_C=$c timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_fuse_$c.log 2>&1
  echo "min_c=$c $(tail -n 1 gpurun_out/ab_fuse_$c.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])')"
done
MCD_FUSE_BN_BWD=1 MCD_FUSE_BN_BWD_MIN_C=0 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_fuse_0b.log 2>&1
echo "min_c=0 again $(tail -n 1 gpurun_out/ab_fuse_0b.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])')"
