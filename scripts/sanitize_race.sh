#!/bin/bash
# racecheck, split by kernel family (run on the GPU box through gpurun):
#   (a) every kernel EXCEPT the tcgen05 / TMA convolutions: plain shared-memory staging with __syncthreads
#   (b) the tcgen05 / TMA convolutions only, all reports kept: their shared memory is synchronised with mbarriers and the
#       async proxy (TMA complete_tx, tcgen05.commit), which racecheck reports as "potential" hazards on the barrier words
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
PY=(python -m pytest -q -p no:cacheprovider --timeout 900 tests/test_kernels_gpu.py tests/test_pipeline_gpu.py)
timeout 900 $CS --tool racecheck --racecheck-report all --print-limit 200 --error-exitcode 86 \
  --kernel-name-exclude kernel_substring=conv_umma \
  "${PY[@]}" -k "head_ or bn_act or input_transform or fast_hist or deconv16s8 or bilinear or ce2d or diff2d or relabel or resize or sgd_step or layout or argmax" \
  > gpurun_out/sanitizer_racecheck_plain.log 2>&1
echo "racecheck (non-tcgen05 kernels) rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_plain.log | tail -n 3
timeout 900 $CS --tool racecheck --racecheck-report all --print-limit 100000 --show-backtrace no --error-exitcode 86 \
  --kernel-name kernel_substring=conv_umma \
  "${PY[@]}" -k "(conv_fprop and umma and (shape0 or shape5 or shape12 or shape20)) or (conv_dgrad_wgrad and umma and (shape0 or shape3)) or streamk" \
  > gpurun_out/sanitizer_racecheck_umma.log 2>&1
echo "racecheck (tcgen05 kernels) rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_umma.log | tail -n 3
python - <<'PY'
import collections, re
c = collections.Counter()
kinds = collections.Counter()
cur = None
for ln in open("gpurun_out/sanitizer_racecheck_umma.log", errors="replace"):
    m = re.match(r"========= Error: (.*?) at __shared__ (0x[0-9a-f]+)", ln)
    if m:
        cur = (m.group(1), int(m.group(2), 16) // 8 * 8)
        kinds[m.group(1)] += 1
    m = re.match(r"=========     (Write|Read) Thread.*? at (?:void )?(?:mcd::)?([A-Za-z_0-9:]+)", ln)
    if m and cur:
        c[(cur[0], m.group(1), m.group(2))] += 1
print("hazard kinds:", dict(kinds))
for k, v in c.most_common(30):
    print(v, k)
PY
