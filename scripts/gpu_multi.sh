mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 "$@" > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N rc=$?"
tail -n 1 gpurun_out/bench_n$N.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])" || tail -n 20 gpurun_out/bench_n$N.log | cut -c1-300
