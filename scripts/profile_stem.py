"""Driver for `ncu --set full` on the HBM-bound stem layers of DRN-D-38 (models/drn.py:126-139) at B images of
480x640: layer0 7x7 6->16, layer1 3x3 16->16, layer2 3x3 s2 16->32, layer3.0 conv1 3x3 s2 32->64 and the 64-channel
3x3 convolutions of layer3: forward, dgrad, wgrad."""
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
for (cin, cout, k, stride, pad, h, w_) in ((6, 16, 7, 1, 3, 480, 640), (16, 16, 3, 1, 1, 480, 640),
                                            (16, 32, 3, 2, 1, 480, 640), (32, 64, 3, 2, 1, 240, 320),
                                            (64, 64, 3, 1, 1, 120, 160), (128, 128, 3, 1, 1, 60, 80)):
    x = ops.to_nhwc(torch.randn(B, cin, h, w_, device=dev))
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    g = ops.conv_geom(x.shape, cin, cout, k, k, stride, 1, pad)
    wf, wd = ops.pack_weight_for(w, g, 0), ops.pack_weight_for(w, g, 1)
    cases.append((x, w, g, wf, wd))


def run():
    for i, (x, w, g, wf, wd) in enumerate(cases):
        y, stats = ops.conv_fprop(x, wf, None, g, want_stats=True)
        if i > 0:
            ops.conv_dgrad(y, wd, g)
        ops.conv_wgrad(x, y, g)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
