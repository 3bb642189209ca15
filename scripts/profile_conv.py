"""Short driver for `ncu --set full`: the dominant convolution shapes of DRN-D-38 at 8 images of 480x640
(layer6: 512->512 3x3 dilation 4 on 60x80; layer5: 256->256 dilation 2), forward / dgrad / wgrad, plus the
BatchNorm backward kernels on the same tensor."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
for (c, dil) in ((512, 4), (256, 2)):
    x = ops.to_nhwc(torch.randn(8, c, 60, 80, device=dev))
    w = torch.randn(c, c, 3, 3, device=dev) * 0.02
    g = ops.conv_geom(x.shape, c, c, 3, 3, 1, dil, dil)
    wf, wd = ops.pack_weight_for(w, g, 0), ops.pack_weight_for(w, g, 1)
    for _ in range(3):
        y, stats = ops.conv_fprop(x, wf, None, g, want_stats=True)
        dx = ops.conv_dgrad(y, wd, g)
        dw, _ = ops.conv_wgrad(x, y, g)
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    for _ in range(2):
        _, stats = ops.conv_fprop(x, wf, None, g, want_stats=True)
        z, save, _ = ops.bn_forward(y, stats, bn, True, res=x)
        ops.bn_bwd(dx, z, y, bn.weight, save, True, True, want_dres=True)
torch.cuda.synchronize()
print("done")
