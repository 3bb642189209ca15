"""ncu source-level profile driver: one convolution shape, forward only (argv: B cin cout h w dil)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

B, cin, cout, h, w_, dil = [int(v) for v in sys.argv[1:7]]
dev = torch.device("cuda")
torch.manual_seed(0)
x = ops.to_nhwc(torch.randn(B, cin, h, w_, device=dev))
w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
g = ops.conv_geom(x.shape, cin, cout, 3, 3, 1, dil, dil)
wf = ops.pack_weight_for(w, g, 0)
for _ in range(3):
    ops.conv_fprop(x, wf, None, g, want_stats=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.conv_fprop(x, wf, None, g, want_stats=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
