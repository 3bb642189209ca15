mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_umma|bn_|wgrad_reduce" -f -o gpurun_out/prof_conv python scripts/profile_conv.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_conv.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_stem python scripts/profile_stem.py > gpurun_out/ncu_stem.log 2>&1
echo "ncu stem rc=$?"; ls -la gpurun_out/prof_stem.ncu-rep
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 16 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log | cut -c1-3500
