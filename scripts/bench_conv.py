"""Isolated timing of the tcgen05 convolution kernels on the DRN-D-38 layer shapes (8 images of 480x640):
fprop (+BN statistics), dgrad (plain / fused ReLU-mask + BN sums), wgrad; stream-K schedule on / off."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SHAPES = [(512, 60, 80, 4), (256, 60, 80, 2), (128, 60, 80, 1), (64, 120, 160, 1)]


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


for (c, h, w, dil) in SHAPES:
    x = ops.to_nhwc(torch.randn(B, c, h, w, device=dev))
    y = ops.to_nhwc(torch.randn(B, c, h, w, device=dev))
    wt = torch.randn(c, c, 3, 3, device=dev) * 0.02
    g = ops.conv_geom(x.shape, c, c, 3, 3, 1, dil, dil)
    wf, wd = ops.pack_weight_for(wt, g, 0), ops.pack_weight_for(wt, g, 1)
    flops = 2.0 * B * h * w * c * c * 9
    for sk in (False, True):
        ops.set_streamk(sk)
        t_f = timeit(lambda: ops.conv_fprop(x, wf, None, g, want_stats=True))
        t_d = timeit(lambda: ops.conv_dgrad(y, wd, g))
        t_da = timeit(lambda: ops.conv_dgrad(y, wd, g, add=x))
        t_df = timeit(lambda: ops.conv_dgrad(y, wd, g, add=x, relu_src=x, bn_y=y))
        print("C=%d %dx%d streamk=%d  fprop %.1f us (%.0f TF/s)  dgrad %.1f (%.0f)  dgrad+add %.1f  dgrad+add+mask+bn %.1f"
              % (c, h, w, sk, t_f, flops / t_f / 1e6, t_d, flops / t_d / 1e6, t_da, t_df))
    t_w = timeit(lambda: ops.conv_wgrad(x, y, g))
    print("C=%d wgrad %.1f us (%.0f TF/s)" % (c, t_w, flops / t_w / 1e6))
