"""Driver for the round-2 `ncu --set full` capture: the kernels that are new or changed in round 2 at the bench geometry
(22 images of 480x640): fused head + loss (Diff2d / cross entropy, deconv heads, with and without filter gradients),
uint8 input transform + relabel, BatchNorm forward (IEEE-half z + bfloat16 twin) / backward, the folded eval-mode unit
(conv + bias + residual + ReLU), and the wide convolution forward / dgrad / wgrad in the round-2 operand formats.
cudaProfilerStart/Stop bracket the measured launches (ncu --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import headloss, ops, pipeline

N = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dev = torch.device("cuda")
torch.manual_seed(0)
x = (torch.randn(N, 41, 60, 80, device=dev) * 2).requires_grad_(True)
wa = (torch.randn(41, 1, 16, 16, device=dev) * 0.1)
wb = (torch.randn(41, 1, 16, 16, device=dev) * 0.1)
lbl = torch.randint(0, 41, (N, 480, 640), device=dev)
cw = torch.ones(41, device=dev)
cw[40] = 0
rgb = torch.randint(0, 256, (N, 480, 640, 3), dtype=torch.uint8, device=dev)
hha = torch.randint(0, 256, (N, 480, 640, 3), dtype=torch.uint8, device=dev)
lbl8 = torch.randint(0, 41, (N, 480, 640), dtype=torch.uint8, device=dev)
cases = []
for (c, dil) in ((512, 4), (256, 2), (64, 1)):
    hw = (60, 80) if c > 64 else (120, 160)
    a = ops.to_nhwc(torch.randn(N, c, *hw, device=dev))
    w = torch.randn(c, c, 3, 3, device=dev) * 0.02
    g = ops.conv_geom(a.shape, c, c, 3, 3, 1, dil, dil)
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    cases.append((a, w, g, ops.pack_weight_for(w, g, 0), ops.pack_weight_for(w, g, 1), bn, torch.zeros(c, device=dev)))


def run():
    for train_w in (False, True):
        wa.requires_grad_(train_w), wb.requires_grad_(train_w)
        headloss.head_diff2d([x], [wa], [x], [wb])
        headloss.head_ce2d([x], [wa], lbl, cw)
    pipeline.transform_images([rgb, hha])
    pipeline.relabel(lbl8, 41)
    for a, w, g, wf, wd, bn, bias in cases:
        y, stats = ops.conv_fprop(a, wf, None, g, want_stats=True)
        z, save, _ = ops.bn_forward(y, stats, bn, True, res=a)
        dx, sums = ops.conv_dgrad(y, wd, g, relu_src=z, bn_y=y)
        ops.bn_bwd(dx, z, y, bn.weight, save, True, True, raw_sums=sums)
        ops.bn_bwd(dx, z, y, bn.weight, save, True, True, want_dres=True)
        ops.conv_wgrad(a, y, g)
        with torch.no_grad():
            ops.conv_fprop_act(a, wf, bias, g, res=a, relu=True)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
