"""ncu driver for the medium layers (64 ch @ 120x160, 128 ch @ 60x80, 16 image pairs): fprop+stats, fused dgrad, wgrad."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
B = 16
for (c, h, w) in ((64, 120, 160), (128, 60, 80)):
    x = ops.to_nhwc(torch.randn(B, c, h, w, device=dev))
    y = ops.to_nhwc(torch.randn(B, c, h, w, device=dev))
    wt = torch.randn(c, c, 3, 3, device=dev) * 0.02
    g = ops.conv_geom(x.shape, c, c, 3, 3, 1, 1, 1)
    wf, wd = ops.pack_weight_for(wt, g, 0), ops.pack_weight_for(wt, g, 1)
    for _ in range(3):
        ops.conv_fprop(x, wf, None, g, want_stats=True)
        ops.conv_dgrad(y, wd, g, add=x, relu_src=x, bn_y=y)
        ops.conv_wgrad(x, y, g)
torch.cuda.synchronize()
print("done")
