"""One eager (serialised: wgrad on the main stream) MCD iteration between cudaProfilerStart/Stop for the ncu
launch list:  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ..."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
sys.path.insert(0, ROOT)
import torch
import bench
from mcd_b200 import nn as mcd_nn
from mcd_b200.step import MCDStep
from loss import CrossEntropyLoss2d, get_prob_distance_criterion
from models.model_util import get_models
from util import get_class_weight_from_file

warnings.simplefilter("ignore")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
models = [m.to(dev).train() for m in get_models("drn_d_38", 6, 41)]
step = MCDStep(models, CrossEntropyLoss2d(get_class_weight_from_file(41).to(dev)), get_prob_distance_criterion("diff"))
src, lbl, tgt = [t.to(dev) for t in bench.synth(B, (480, 640), 1)]
mcd_nn.set_overlap_wgrad(False)
for _ in range(2):
    step(src, lbl, tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(src, lbl, tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
