mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_umma_fprop -f -o /tmp/p64 python scripts/profile_one.py 16 64 64 120 160 1 > gpurun_out/ncu_p64.log 2>&1
ncu -i /tmp/p64.ncu-rep --page source --csv > gpurun_out/src_p64.csv 2>gpurun_out/src_p64.err
ncu -i /tmp/p64.ncu-rep --page raw --csv > gpurun_out/raw_p64.csv 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_umma_fprop -f -o /tmp/p256 python scripts/profile_one.py 16 256 256 60 80 2 > gpurun_out/ncu_p256.log 2>&1
ncu -i /tmp/p256.ncu-rep --page source --csv > gpurun_out/src_p256.csv 2>gpurun_out/src_p256.err
ncu -i /tmp/p256.ncu-rep --page raw --csv > gpurun_out/raw_p256.csv 2>/dev/null
ls -la gpurun_out/src_* gpurun_out/raw_p*
