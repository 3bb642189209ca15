#!/bin/bash
# compute-sanitizer passes over the kernel-level parity tests (run on the GPU box through gpurun):
#   memcheck  - every kernel test (tcgen05 / TMA convolutions incl. stream-K and cta_group::2 pairs, BatchNorm, heads,
#               losses, fused head+loss, input pipeline, evaluation counts) + one small MCD iteration (smoke)
#   racecheck - shared-memory hazards of the kernels with hand-rolled staging (head_loss, BatchNorm, pipeline, stem convs)
#   synccheck - barrier misuse, same subset
# Writes gpurun_out/sanitizer_{memcheck,racecheck,synccheck}.log; copy the summaries to profiles/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
PY=(python -m pytest -q -p no:cacheprovider --timeout 900 tests/test_kernels_gpu.py tests/test_pipeline_gpu.py)
K_ALL=(-k "not mcdstep")
K_RACE=(-k "head_ or bn_act or input_transform or fast_hist or deconv16s8 or (conv_fprop and umma and shape0) or (conv_dgrad_wgrad and umma and shape0) or streamk")
run() {  # tool, log, extra sanitizer args..., then pytest selection
  tool=$1; log=$2; shift 2
  timeout ${SAN_TIMEOUT:-1500} $CS --tool $tool --print-limit 20 --error-exitcode 86 "$@" > gpurun_out/$log 2>&1
  rc=$?
  echo "$tool rc=$rc"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/$log | tail -n 4
}
run memcheck sanitizer_memcheck.log --leak-check no "${PY[@]}" "${K_ALL[@]}"
run memcheck sanitizer_memcheck_smoke.log --leak-check no python __graft_entry__.py smoke
run racecheck sanitizer_racecheck.log --racecheck-report all "${PY[@]}" "${K_RACE[@]}"
run synccheck sanitizer_synccheck.log "${PY[@]}" "${K_RACE[@]}"
