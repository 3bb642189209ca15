# round-2e final validation: the whole GPU suite twice (flakiness), smoke, headline bench, MFNet-add bench, micro-benchmark
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_variants.txt
for i in 1 2; do
  timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all_$i.log 2>&1; echo "pytest -m gpu run $i rc=$?"
  grep -E "passed|failed|^FAILED" gpurun_out/t_all_$i.log | head -10
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log > gpurun_out/r02e_bench_early_b22.json; cut -c1-330 gpurun_out/r02e_bench_early_b22.json
timeout 600 python bench.py --workload mfnet-add > gpurun_out/bench_mfnet.log 2>&1; echo "bench mfnet-add rc=$?"
tail -n 1 gpurun_out/bench_mfnet.log > gpurun_out/r02e_bench_mfnet-add_b22.json; cut -c1-330 gpurun_out/r02e_bench_mfnet-add_b22.json
timeout 300 python scripts/bench_variants.py 8 > gpurun_out/bench_variants.txt 2>&1; cat gpurun_out/bench_variants.txt
