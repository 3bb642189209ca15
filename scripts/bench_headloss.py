"""micro-benchmark of the fused head + loss kernels (csrc/headloss.cu) at the bench geometry: N images, 41 classes,
60x80 score maps -> 480x640.   python scripts/bench_headloss.py [N]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
from mcd_b200 import headloss  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = (torch.randn(N, 41, 60, 80, device=dev) * 2).requires_grad_(True)
x2 = (torch.randn(N, 41, 60, 80, device=dev) * 2).requires_grad_(True)
wa = (torch.randn(41, 1, 16, 16, device=dev) * 0.1)
wb = (torch.randn(41, 1, 16, 16, device=dev) * 0.1)
lbl = torch.randint(0, 41, (N, 480, 640), device=dev)
cw = torch.ones(41, device=dev)
cw[40] = 0


def timeit(name, fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %8.1f us" % (name, e0.elapsed_time(e1) / iters * 1e3))


for train_w in (False, True):
    wa.requires_grad_(train_w), wb.requires_grad_(train_w)
    tag = "dx+dw" if train_w else "dx"
    timeit("diff2d deconv  (%s)" % tag, lambda: headloss.head_diff2d([x], [wa], [x], [wb]))
    timeit("ce2d   deconv  (%s)" % tag, lambda: headloss.head_ce2d([x], [wa], lbl, cw))
timeit("diff2d bilinear (dx)", lambda: headloss.head_diff2d([x], None, [x2], None))
timeit("ce2d   bilinear (dx)", lambda: headloss.head_ce2d([x], None, lbl, cw))
with torch.no_grad():
    timeit("diff2d deconv  (loss only)", lambda: headloss.head_diff2d([x], [wa], [x], [wb]))
