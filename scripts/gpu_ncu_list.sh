mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 7100 -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv; tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
