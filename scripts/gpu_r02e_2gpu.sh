mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_two_gpu.txt gpurun_out/parity_variants.txt
timeout 500 python -m pytest tests/test_steps_gpu.py -m gpu -q -k two_gpu --timeout 400 -p no:cacheprovider > gpurun_out/t_two_gpu.log 2>&1; echo "two_gpu rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_two_gpu.log | cut -c1-300 | head -20
cat gpurun_out/parity_two_gpu.txt
timeout 200 python -m pytest tests/test_variants_gpu.py -m gpu -q -k "mfnet_tester or segbd_tester" --timeout 150 -p no:cacheprovider > gpurun_out/t_variants2.log 2>&1; echo "testers rc=$?"
grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_variants2.log | cut -c1-300 | head -20
cat gpurun_out/parity_variants.txt
