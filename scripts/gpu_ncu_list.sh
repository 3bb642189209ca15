mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 13000 -c 3400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
