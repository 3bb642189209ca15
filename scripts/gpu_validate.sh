mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/t_all.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log > gpurun_out/bench_final.json; cut -c1-1500 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"conv_umma|bn_|wgrad_reduce" -f -o /tmp/prof_conv python scripts/profile_conv.py 22 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 22 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
du -sh gpurun_out
