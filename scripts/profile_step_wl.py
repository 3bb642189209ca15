"""One eager (serialised) iteration of any bench workload between cudaProfilerStart/Stop for an ncu launch list:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... \
       python scripts/profile_step_wl.py triple 22"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

warnings.simplefilter("ignore")
wl = sys.argv[1] if len(sys.argv) > 1 else "triple"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 22
bench.use_ours()
from mcd_b200 import nn as mcd_nn
dev = torch.device("cuda", 0)
step, _ = bench.build_step(wl, dev)
src, lbl, tgt = [t.to(dev) for t in bench.synth(B, bench.FULL, 1, bench.WORKLOADS[wl]["src_ch"])]
mcd_nn.set_overlap_wgrad(False)
for _ in range(2):
    step(src, lbl, tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(src, lbl, tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
