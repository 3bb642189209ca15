# round-2 re-entry validation: the whole GPU suite, smoke, and the headline bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_variants.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | head -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log > gpurun_out/r02e_bench_early_b22.json; cut -c1-900 gpurun_out/r02e_bench_early_b22.json
cat gpurun_out/parity_variants.txt
