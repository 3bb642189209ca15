"""CUDA-event micro-benchmark of the option-surface kernels (csrc/variants.cu) at the sizes they see on 480x640 inputs:
achieved algorithmic GB/s against the measured HBM copy bandwidth (MEASURED_PEAKS.json, fallback 6544 GB/s).

    python scripts/bench_variants.py [N] > gpurun_out/bench_variants.txt
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200"))
from mcd_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
peak = 6544.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print("%-34s %9.1f MB  %8.3f ms  %7.0f GB/s  %5.1f %% of %.0f" % (name, nbytes / 1e6, ms, gbs, 100 * gbs / peak, peak))


full = (N, 41, 480, 640)
E = 4 * N * 41 * 480 * 640
x1, x2, a, d = (torch.randn(full, device=dev) for _ in range(4))
report("gate_fuse_fwd", 4 * E, timed(lambda: ops.gate_fuse_fwd(x1, x2, a)))
report("gate_fuse_bwd", 7 * E, timed(lambda: ops.gate_fuse_bwd(x1, x2, a, d)))
p = ops.softmax_ch_fwd(x1)
report("softmax_ch_fwd", 2 * E, timed(lambda: ops.softmax_ch_fwd(x1)))
report("softmax_ch_bwd", 3 * E, timed(lambda: ops.softmax_ch_bwd(p, d)))
report("cat2_f32", 4 * E, timed(lambda: ops.cat2_f32(x1, x2)))
cat = ops.cat2_f32(x1, x2)
report("split2_f32", 4 * E, timed(lambda: ops.split2_f32(cat, 41)))
del cat
t = torch.randint(0, 41, (N, 480, 640), device=dev)
w = torch.ones(41, device=dev)
report("prob_ce2d_fwd", N * 480 * 640 * 12, timed(lambda: ops.prob_ce2d_fwd(p, t, w, -100)))
acc = ops.prob_ce2d_fwd(p, t, w, -100).clone()
gs = torch.ones(1, device=dev)
report("prob_ce2d_bwd", E + N * 480 * 640 * 12, timed(lambda: ops.prob_ce2d_bwd(p, t, w, -100, acc, gs)))
report("add3_f32", 4 * E, timed(lambda: ops.add3_f32(x1, x2, a)))
report("sigmoid_fwd", 2 * E, timed(lambda: ops.sigmoid_fwd(x1)))
s = torch.randn(N, 41, 60, 80, device=dev)
report("bilinear_ac_up_fwd (fp32 out)", E + E // 64, timed(lambda: ops.bilinear_ac_up_fwd(s, 8, True)))
report("bilinear_ac_up_fwd (bf16 out)", E // 2 + E // 64, timed(lambda: ops.bilinear_ac_up_fwd(s, 8, False)))
report("bilinear_ac_up_bwd (fp32 in)", E + E // 64, timed(lambda: ops.bilinear_ac_up_bwd(d, 8)))
report("bilinear_up_fwd (fp32 out, ref.)", E + E // 64, timed(lambda: ops.bilinear_up_fwd(s, 8, True)))
report("bilinear_up_bwd (fp32 in, ref.)", E + E // 64, timed(lambda: ops.bilinear_up_bwd(d, 8)))
del x1, x2, a, d, p
u = ops.to_nhwc(torch.randn(4 * N, 512, 60, 80, device=dev), twin=False)
v = ops.to_nhwc(torch.randn(4 * N, 512, 60, 80, device=dev), twin=False)
nb = u.numel() * 2
report("add_nhwc (fp16 + bf16 twin out)", 4 * nb, timed(lambda: ops.add_nhwc(u, v, twin=True)))
report("add_nhwc (fp16 out)", 3 * nb, timed(lambda: ops.add_nhwc(u, v, twin=False)))
