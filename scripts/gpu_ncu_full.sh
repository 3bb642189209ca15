mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_umma|bn_reduce|bn_bwd_apply" -s 8 -c 12 -f -o gpurun_out/prof_conv python scripts/profile_conv.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_conv.ncu-rep
