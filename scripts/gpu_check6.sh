mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -p no:cacheprovider -k "mfnet or triple" > gpurun_out/t3_cfg.log 2>&1; echo "cfg rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t3_cfg.log | head -20
