mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | head
