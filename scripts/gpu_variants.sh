mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_variants.txt
timeout 600 python -m pytest tests/test_variants_gpu.py -m gpu -q --timeout 150 -p no:cacheprovider > gpurun_out/t_variants.log 2>&1; echo "variants rc=$?"
grep -E "passed|failed|^E  |^FAILED|Error" gpurun_out/t_variants.log | head -60
cat gpurun_out/parity_variants.txt
