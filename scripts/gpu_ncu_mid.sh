mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_umma" -s 9 -c 9 -f -o gpurun_out/prof_mid python scripts/profile_mid.py > gpurun_out/ncu_mid.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_mid.ncu-rep
