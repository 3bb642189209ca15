"""micro-benchmark: dgrad with the fused ReLU-mask / BatchNorm-sum epilogue and the folded inference unit (addend + ReLU)
for the 64- / 128- / 256- / 512-channel 3x3 layers at N images (CUDA events, 20 launches each)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dev = torch.device("cuda")
torch.manual_seed(0)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (c, dil, hw) in ((64, 1, (120, 160)), (128, 1, (60, 80)), (256, 2, (60, 80)), (512, 4, (60, 80))):
    a = ops.to_nhwc(torch.randn(N, c, *hw, device=dev))
    w = torch.randn(c, c, 3, 3, device=dev) * 0.02
    g = ops.conv_geom(a.shape, c, c, 3, 3, 1, dil, dil)
    wf, wd = ops.pack_weight_for(w, g, 0), ops.pack_weight_for(w, g, 1)
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    y, stats = ops.conv_fprop(a, wf, None, g, want_stats=True)
    z, save, _ = ops.bn_forward(y, stats, bn, True, res=a)
    dy = ops.convert16(y)
    bias = torch.zeros(c, device=dev)
    t_f = timeit(lambda: ops.conv_fprop(a, wf, None, g, want_stats=True))
    t_d = timeit(lambda: ops.conv_dgrad(dy, wd, g, relu_src=z, bn_y=y))
    t_p = timeit(lambda: ops.conv_dgrad(dy, wd, g))
    with torch.no_grad():
        t_a = timeit(lambda: ops.conv_fprop_act(a, wf, bias, g, res=a, relu=True))
    print("%4d ch: forward %7.1f us | dgrad + mask + BN sums %7.1f | plain dgrad %7.1f | folded unit (res + relu) %7.1f"
          % (c, t_f, t_d, t_p, t_a))
