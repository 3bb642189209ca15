"""Short driver for `ncu --set full`: the dominant convolution shapes of DRN-D-38 at 16 images of 480x640
(layer6: 512->512 3x3 dilation 4 on 60x80; layer5: 256->256 dilation 2), forward / dgrad (with the fused
BatchNorm-backward epilogue) / wgrad, plus the BatchNorm kernels on the same tensors.  cudaProfilerStart/Stop
bracket the measured launches (ncu --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")
torch.manual_seed(0)
cases = []
for (c, dil) in ((512, 4), (256, 2)):
    x = ops.to_nhwc(torch.randn(B, c, 60, 80, device=dev))
    w = torch.randn(c, c, 3, 3, device=dev) * 0.02
    g = ops.conv_geom(x.shape, c, c, 3, 3, 1, dil, dil)
    wf, wd = ops.pack_weight_for(w, g, 0), ops.pack_weight_for(w, g, 1)
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    cases.append((x, w, g, wf, wd, bn))


def run():
    for x, w, g, wf, wd, bn in cases:
        y, stats = ops.conv_fprop(x, wf, None, g, want_stats=True)
        z, save, _ = ops.bn_forward(y, stats, bn, True, res=x)
        dx, sums = ops.conv_dgrad(y, wd, g, relu_src=z, bn_y=y)
        ops.bn_bwd(dx, z, y, bn.weight, save, True, True, raw_sums=sums)
        ops.bn_bwd(dx, z, y, bn.weight, save, True, True, want_dres=True)
        dw, _ = ops.conv_wgrad(x, y, g)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
