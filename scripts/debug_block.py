import sys, os, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200")); sys.path.insert(0, ROOT)
import torch
from oracle import mcd_oracle as O
from mcd_b200 import ops
from models.drn import BasicBlock
warnings.simplefilter("ignore")
dev = torch.device("cuda")
torch.manual_seed(0)
C = 64
blk = BasicBlock(C, C).to(dev).train()
sd = {("u." + k): v.detach().clone() for k, v in blk.state_dict().items()}
O.fill_state_dict_(sd, 3)
blk.load_state_dict({k[2:]: v for k, v in sd.items()})
sd = O.to_device(sd, dev)
unit = ("block", "u", 1, 1, 1, False)
g = torch.Generator().manual_seed(1)
x = torch.relu(torch.randn(2, C, 60, 80, generator=g)).to(dev)
dz = (torch.randn(2, C, 60, 80, generator=g) * 1e-3).to(dev)
def run_oracle(st):
    s = {k: v.clone().requires_grad_(torch.is_floating_point(v) and "running" not in k) for k, v in sd.items()}
    xe = x.clone().requires_grad_(True)
    taps = {}
    with O.storage(st):
        out = O.unit_forward(s, unit, O._q(xe), True, taps)
    pk = [k for k in s if s[k].requires_grad]
    gr = torch.autograd.grad(out, [xe] + [s[k] for k in pk], dz)
    return out.detach(), gr[0], dict(zip(pk, gr[1:])), taps
o32, dx32, gp32, t32 = run_oracle(None)
o16, dx16, gp16, t16 = run_oracle(torch.bfloat16)
xin = ops.to_nhwc(x).requires_grad_(True)
out = blk(xin)
out.backward(ops.to_nhwc(dz))
torch.cuda.synchronize()
def nerr(a, b): return float((a.float()-b.float()).abs().max()/b.float().abs().max())
def l2(a, b): return float((a.float()-b.float()).norm()/b.float().norm())
o = ops.to_nchw_f32(out.detach()); dx = ops.to_nchw_f32(xin.grad)
print("act  vs32 %.3e vs16 %.3e   (16 vs 32: %.3e)" % (nerr(o, o32), nerr(o, o16), nerr(o16, o32)))
print("dx   vs32 max %.3e l2 %.3e | vs16 max %.3e l2 %.3e | 16vs32 max %.3e l2 %.3e" % (nerr(dx, dx32), l2(dx, dx32), nerr(dx, dx16), l2(dx, dx16), nerr(dx16, dx32), l2(dx16, dx32)))
print("mask diff ours vs 32: %.5f  ours vs 16: %.5f" % (float(((o > 0) != (o32 > 0)).float().mean()), float(((o > 0) != (o16 > 0)).float().mean())))
for k, p in blk.named_parameters():
    print("%-12s vs32 %.3e vs16 %.3e  (16vs32 %.3e)" % (k, nerr(p.grad, gp32["u." + k]), nerr(p.grad, gp16["u." + k]), nerr(gp16["u."+k], gp32["u."+k])))
# isolate: same thing with direct algo
prev = ops.set_conv_algo(1)
xin2 = ops.to_nhwc(x).requires_grad_(True); blk.zero_grad()
out2 = blk(xin2); out2.backward(ops.to_nhwc(dz)); torch.cuda.synchronize()
ops.set_conv_algo(prev)
dx2 = ops.to_nchw_f32(xin2.grad)
print("direct algo: dx vs32 max %.3e l2 %.3e | vs16 max %.3e | umma-vs-direct max %.3e" % (nerr(dx2, dx32), l2(dx2, dx32), nerr(dx2, dx16), nerr(dx, dx2)))
