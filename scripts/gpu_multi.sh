mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench n$N rc=$?"
tail -n 1 gpurun_out/bench_n$N.log | cut -c1-900
