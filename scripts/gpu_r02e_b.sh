# re-run of the adjusted tests, BatchNorm large-tensor test repeated, sanitizer passes over the variants kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_variants.txt
timeout 300 python -m pytest tests/test_variants_gpu.py -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/t_variants.log 2>&1; echo "variants rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_variants.log | head -20
for i in 1 2 3 4; do timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "bn_act" -p no:cacheprovider 2>&1 | tail -n 1; done
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 400 $CS --tool memcheck --leak-check no --print-limit 20 --error-exitcode 86 python -m pytest -q -p no:cacheprovider --timeout 350 tests/test_variants_gpu.py -k "kernels_vs_torch or fusion_heads or arch_c_units" > gpurun_out/sanitizer_memcheck_variants.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_variants.log | tail -n 3
timeout 300 $CS --tool racecheck --racecheck-report all --print-limit 20 --error-exitcode 86 python -m pytest -q -p no:cacheprovider --timeout 250 tests/test_variants_gpu.py -k "kernels_vs_torch" > gpurun_out/sanitizer_racecheck_variants.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_variants.log | tail -n 3
cat gpurun_out/parity_variants.txt | tail -n 14
