mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_graph.log 2>&1; echo "bench graph rc=$?"
tail -n 3 gpurun_out/bench_graph.log | cut -c1-1800
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_eager.log 2>&1; echo "bench eager rc=$?"
tail -n 1 gpurun_out/bench_eager.log | cut -c1-400
