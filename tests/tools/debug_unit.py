"""Stage-by-stage comparison of one conv-BN-ReLU unit (512 -> 512, 3x3) on the GPU against a torch fp32 emulation that
rounds at the same storage points (GPU test tool).   python tests/tools/debug_unit.py [C] [dil]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
from mcd_b200 import abi, ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dil = int(sys.argv[2]) if len(sys.argv) > 2 else 1
HF, BF = torch.float16, torch.bfloat16


def l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def rd(t, dt):
    return t.to(dt).float()


g0 = torch.Generator().manual_seed(3)
n, h, w = 2, 30, 40
x = torch.relu(torch.randn(n, C, h, w, generator=g0) + 0.3).to(dev)
W = (torch.randn(C, C, 3, 3, generator=g0) * (2.0 / (9 * C)) ** 0.5).to(dev)
gamma = (torch.rand(C, generator=g0) + 0.5).to(dev)
beta = (torch.randn(C, generator=g0) * 0.1).to(dev)
# a structured upstream gradient: smooth over pixels plus noise (like a real loss gradient), tiny magnitude
dz_up = (torch.randn(n, C, h, w, generator=g0) * 1e-6).to(dev)

for algo_name, algo in (("umma", abi.ALGO_UMMA), ("direct", abi.ALGO_DIRECT)):
    with torch.no_grad():
        # ---- emulation (fp32 math, storage rounding) -- double for the big reductions
        x16, w16 = rd(x, HF), rd(W, HF)
        y_e = rd(F.conv2d(x16.double(), w16.double(), None, 1, dil, dil).float(), HF)
        mean = y_e.double().mean((0, 2, 3)); var = y_e.double().var((0, 2, 3), unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        xhat = ((y_e.double() - mean[None, :, None, None]) * rstd[None, :, None, None])
        v_e = xhat * gamma.double()[None, :, None, None] + beta.double()[None, :, None, None]
        z_e = torch.relu(v_e).float()
        g_e = rd(dz_up, BF).double() * (v_e > 0)
        m1 = g_e.mean((0, 2, 3)); m2 = (g_e * xhat).mean((0, 2, 3))
        dy_e = rd(((gamma.double() * rstd)[None, :, None, None] * (g_e - m1[None, :, None, None] - xhat * m2[None, :, None, None])).float(), BF)
        dx_e = torch.nn.grad.conv2d_input(x.shape, rd(W, BF).double(), dy_e.double(), 1, dil, dil)
        dW_e = torch.nn.grad.conv2d_weight(rd(x, BF).double(), W.shape, dy_e.double(), 1, dil, dil)
    # ---- library
    xn = ops.to_nhwc(x)
    gm = ops.conv_geom(xn.shape, C, C, 3, 3, 1, dil, dil)
    y, stats = ops.conv_fprop(xn, ops.pack_weight_for(W, gm, 0, algo), None, gm, want_stats=True, algo=algo)
    bn = torch.nn.BatchNorm2d(C).to(dev)
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
    z, save, _ = ops.bn_forward(y, stats, bn, True)
    dzn = ops.to_nhwc(dz_up, grad=True)
    dy, dgam, dbet, _, _, _ = ops.bn_bwd(dzn, z, y, bn.weight, save, True, True)
    dx = ops.conv_dgrad(dy, ops.pack_weight_for(W, gm, 1, algo), gm, algo=algo)
    dW, _ = ops.conv_wgrad(xn, dy, gm, algo=algo)
    # the same kernels fed the EMULATION's dy: isolates dgrad / wgrad
    dyn_e = ops.to_nhwc(dy_e, grad=True)
    dx2 = ops.conv_dgrad(dyn_e, ops.pack_weight_for(W, gm, 1, algo), gm, algo=algo)
    dW2, _ = ops.conv_wgrad(xn, dyn_e, gm, algo=algo)
    torch.cuda.synchronize()
    yg = ops.to_nchw_f32(y); zg = ops.to_nchw_f32(z); dyg = ops.to_nchw_f32(dy)
    flips = float(((zg > 0) != (z_e > 0)).float().mean())
    print("[%s C=%d dil=%d] y %.2e (exact-equal %.4f) mean %.2e rstd %.2e z %.2e flips %.2e | dy %.2e | dx %.2e dW %.2e | "
          "dx(given dy_e) %.2e dW(given dy_e) %.2e | dgamma %.2e dbeta %.2e" % (
              algo_name, C, dil, l2(yg, y_e), float((yg == y_e).float().mean()), l2(save[0], mean), l2(save[1], rstd),
              l2(zg, z_e), flips, l2(dyg, dy_e), l2(ops.to_nchw_f32(dx), dx_e), l2(dW, dW_e),
              l2(ops.to_nchw_f32(dx2), dx_e), l2(dW2, dW_e), l2(dgam, (g_e * xhat).sum((0, 2, 3))),
              l2(dbet, g_e.sum((0, 2, 3)))))
