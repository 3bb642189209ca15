mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "not umma" --timeout 300 -p no:cacheprovider > gpurun_out/t1_direct.log 2>&1; echo "direct rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "umma and fprop" --timeout 120 -p no:cacheprovider > gpurun_out/t2_umma_fprop.log 2>&1; echo "umma_fprop rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "umma and dgrad" --timeout 120 -p no:cacheprovider > gpurun_out/t3_umma_bwd.log 2>&1; echo "umma_bwd rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_modules_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/t4_modules.log 2>&1; echo "modules rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t1_direct.log gpurun_out/t2_umma_fprop.log gpurun_out/t3_umma_bwd.log gpurun_out/t4_modules.log
