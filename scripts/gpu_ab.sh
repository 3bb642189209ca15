mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider -k "conv" > gpurun_out/t1_kernels.log 2>&1; rc=$?; echo "kernels rc=$rc"; grep -E "passed|failed" gpurun_out/t1_kernels.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|E  )" gpurun_out/t1_kernels.log | head -40; exit 1; fi
timeout 300 python scripts/bench_conv.py 2>&1 | tail -12
for cfg in "1 1" "1 0" "0 1" "0 0"; do
  set -- $cfg
  MCD_FUSE_BN_BWD=$1 MCD_STREAMK=$2 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch ${BATCH:-8} 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse=$1 streamk=$2', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
