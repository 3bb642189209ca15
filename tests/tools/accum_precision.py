"""How exact is the fp32 accumulation of tcgen05.mma (kind::f16)?  (GPU test tool)

fp16-exact inputs, planar fp32 output (straight from the TMEM accumulators, no 16-bit output rounding), against an
fp64 convolution; the CUDA-core kernel (fp32 FMA chain) on the same packed weights is the yard-stick.
    python tests/tools/accum_precision.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
from mcd_b200 import abi, ops  # noqa: E402

dev = torch.device("cuda", 0)
for cin, cout, k, dil, relu_in in [(64, 64, 3, 1, True), (256, 256, 3, 2, True), (512, 512, 3, 4, True),
                                   (512, 512, 3, 4, False), (512, 512, 1, 1, True)]:
    g0 = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, 30, 40, generator=g0)
    if relu_in:
        x = torch.relu(x + 0.5)
    x = x.half().float().to(dev)
    w = (torch.randn(cout, cin, k, k, generator=g0) * (2.0 / (k * k * cin)) ** 0.5).half().float().to(dev)
    pad = dil * (k // 2)
    ref = F.conv2d(x.double(), w.double(), None, 1, pad, dil)
    with torch.no_grad():
        xn = ops.to_nhwc(x)
        gm = ops.conv_geom(xn.shape, cin, cout, k, k, 1, dil, pad)
        res = {}
        for name, algo in (("umma", abi.ALGO_UMMA), ("direct", abi.ALGO_DIRECT)):
            y, _ = ops.conv_fprop(xn, ops.pack_weight_for(w, gm, 0, algo), None, gm, planar=True, algo=algo)
            e = (y.double() - ref)
            res[name] = (float(e.abs().max() / ref.abs().max()), float(e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
                         float(e.mean() / ref.abs().mean()))
    print("cin %4d k %d relu_in %d  K=%5d | umma max %.2e rms %.2e bias %+.2e | direct max %.2e rms %.2e bias %+.2e" %
          ((cin, k, relu_in, cin * k * k) + res["umma"] + res["direct"]))
