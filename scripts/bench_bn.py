"""micro-benchmark of the BatchNorm streaming kernels at the shapes of an MCD iteration (N images of 480x640):
forward (y [+ res] -> z + twin) and backward apply; prints microseconds and achieved GB/s (algorithmic bytes).
MCD_BN_BULK=0 selects the register-staged kernels."""
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


print("bulk" if os.environ.get("MCD_BN_BULK", "1") != "0" else "register-staged")
for (c, hw, with_res) in ((512, (60, 80), True), (512, (60, 80), False), (256, (60, 80), True), (128, (60, 80), False),
                          (64, (120, 160), True), (32, (240, 320), False), (16, (480, 640), False)):
    y = ops.to_nhwc(torch.randn(N, c, *hw, device=dev))
    y16 = ops.h16(y)
    res = y if with_res else None
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    stats = torch.cat([y16.float().sum((0, 2, 3)), y16.float().pow(2).sum((0, 2, 3))]).contiguous()
    elems = y16.numel()
    t_f = timeit(lambda: ops.bn_forward(y16, stats, bn, True, res=res, twin=True))
    z, save, _ = ops.bn_forward(y16, stats, bn, True, res=res, twin=True)
    dz = ops.convert16(y16)
    sums = torch.zeros(2 * c, device=dev)
    t_b = timeit(lambda: ops.bn_bwd(dz, z, y16, bn.weight, save, True, True, raw_sums=sums))
    t_r = timeit(lambda: ops.bn_bwd(dz, z, y16, bn.weight, save, True, True, want_dres=with_res))
    bf = elems * 2 * (4 if with_res else 3)
    bb = elems * 2 * 3            # dz, y in; dy out (mask comes with dz when the sums were fused)
    print("%4d ch %s: forward %7.1f us %6.0f GB/s | backward apply (fused sums) %7.1f us %6.0f GB/s | reduce + apply %7.1f us"
          % (c, "res" if with_res else "   ", t_f, bf / t_f / 1e3, t_b, bb / t_b / 1e3, t_r))
