mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"conv_umma|bn_|wgrad_reduce" -f -o /tmp/prof_conv python scripts/profile_conv.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la /tmp/prof_conv.ncu-rep
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_stem python scripts/profile_stem.py > gpurun_out/ncu_stem.log 2>&1
echo "ncu stem rc=$?"; ls -la /tmp/prof_stem.ncu-rep
ncu -i /tmp/prof_stem.ncu-rep --page raw --csv > gpurun_out/prof_stem_raw.csv 2>/dev/null
for f in /tmp/prof_conv.ncu-rep /tmp/prof_stem.ncu-rep; do
  if [ $(stat -c %s $f) -lt 12000000 ]; then cp $f gpurun_out/; fi
done
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 16 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_modules_gpu.py -q --timeout 900 -p no:cacheprovider -x > gpurun_out/t2_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t2_parity.log | head
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench.log | cut -c1-1200
du -sh gpurun_out
