# head + loss kernel: where the time goes (compile-time experiment builds, see HL_EXP in csrc/headloss.cu)
export PYTHONUNBUFFERED=1
P=$PWD/multichannel-semseg-with-uda_b200
echo "== full kernel"; timeout 200 python scripts/bench_headloss.py 2>&1 | tail -7
for k in 1 2 3; do echo "== HL_EXP=$k (1: logits only, 2: + softmax/loss/gradient math, 3: + logit-gradient stores; no phase 2)"; MCD_LIB_PATH=$P/libmcd_exp$k.so timeout 200 python scripts/bench_headloss.py 2>&1 | tail -7; done
