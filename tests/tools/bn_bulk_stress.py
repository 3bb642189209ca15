"""stress the bulk-staged BatchNorm kernels: repeat forward + backward of a 9.3 MB tensor and compare every run with the
first one (bitwise for the forward and the residual gradient, 1e-2 of max for dy, whose statistics are reduced with
atomics).  MCD_BN_BULK=0 runs the register-staged forward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
import torch
from mcd_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
n, c, h, w = 3, 64, 150, 161
y = ops.h16(ops.to_nhwc(torch.randn(n, c, h, w, device=dev) * 2 + 0.5))
r = ops.h16(ops.to_nhwc(torch.randn(n, c, h, w, device=dev)))
dz = ops.to_nhwc(torch.randn(n, c, h, w, device=dev), grad=True)
bn = torch.nn.BatchNorm2d(c).to(dev).train()
stats = torch.cat([y.float().sum((0, 2, 3)), y.float().pow(2).sum((0, 2, 3))]).contiguous()
ref = None
bad = {"z": 0, "dy": 0, "dres": 0}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for i in range(iters):
    z, save, _ = ops.bn_forward(y, stats, bn, True, res=r, twin=True)
    dy, dg, db, dres, _, _ = ops.bn_bwd(dz, z, y, bn.weight, save, True, True, res=r, want_dres=True)
    torch.cuda.synchronize()
    cur = (z.float().clone(), dy.float().clone(), dres.float().clone())
    if ref is None:
        ref = cur
        continue
    bad["z"] += int(not torch.equal(cur[0], ref[0]))
    bad["dy"] += int(float((cur[1] - ref[1]).abs().max()) > 1e-2 * float(ref[1].abs().max()))
    bad["dres"] += int(not torch.equal(cur[2], ref[2]))
print("bulk fwd =", os.environ.get("MCD_BN_BULK", "1"), "iterations", iters, "mismatching runs", bad)
