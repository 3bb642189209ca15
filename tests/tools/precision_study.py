"""Which STORAGE points decide the per-unit gradient parity?  (test tool, CPU, uses the oracle)

Replays tests/test_parity_gpu.py::test_per_layer_forward_backward_vs_oracle entirely inside the fp32 oracle: every
DRN unit is fed the fp32 oracle's input and upstream gradient and run with storage emulation
(`oracle.storage(...)`) under several configurations; prints, per configuration, the worst relative-L2 error of
dx and of the parameter gradients against fp32, the worst activation max-norm error, and the ReLU-mask flip rate.

    python tests/tools/precision_study.py [H W N]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mcd_oracle as O  # noqa: E402

BF, HF = torch.bfloat16, torch.float16


def l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    h, w, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (120, 160, 2)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    G = O.fill_state_dict_(O.init_seg_base("drn_d_38", 6, 41), 1)
    F1 = O.fill_state_dict_(O.init_head(41), 2)
    g = torch.Generator().manual_seed(5)
    src = torch.randn(n, 6, h, w, generator=g)
    lbl = torch.randint(0, 41, (n, h, w), generator=g)
    wgt = O.class_weight(41)
    O._req([G, F1])
    taps = {}
    feat = O.seg_base_forward(G, src, taps=taps)
    o1 = O.head_forward(F1, feat)
    loss = O.ce2d(o1, lbl, wgt)
    keys = [k for k in taps if k.endswith(":out")]
    pn = O.trainable(G)
    grads = torch.autograd.grad(loss, [taps[k] for k in keys] + [G[k] for k in pn])
    d_out = dict(zip(keys, grads[:len(keys)]))
    gG = dict(zip(pn, grads[len(keys):]))
    units = [u for st in O.trunk_spec("drn_d_38", "base.") for u in st]

    configs = {
        "bf16 all": dict(dtype=BF),
        "fp16 all": dict(dtype=HF),
        "bf16, y fp32": dict(dtype=BF, y=None),
        "bf16, y fp16": dict(dtype=BF, y=HF),
        "bf16, grad fp32": dict(dtype=BF, grad=None),
        "bf16, w fp32": dict(dtype=BF, w=None),
        "bf16, act fp32": dict(dtype=BF, act=None),
        "only y bf16": dict(dtype=None, y=BF),
        "only act bf16": dict(dtype=None, act=BF),
        "only w bf16": dict(dtype=None, w=BF),
        "only grad bf16": dict(dtype=None, grad=BF),
        "fp16 fwd, bf16 grad": dict(dtype=HF, grad=BF),
        "fp16 fwd, fp32 y, bf16 grad": dict(dtype=HF, grad=BF, y=None),
    }
    print("# %dx%d n=%d: worst over units of  act-maxnorm | dx rel-L2 | param-grad rel-L2   (vs fp32)" % (h, w, n))
    for name, cfg in configs.items():
        cfg = dict(cfg)
        dt = cfg.pop("dtype")
        worst = [0.0, 0.0, 0.0]
        per = []
        x_in, prev = src, None
        for unit in units:
            key = O.unit_key(unit)
            prefix = key[:-4] + "."
            if unit[0] == "cbr":
                prefix = key.split(":")[0][:-1]
            sd_u = {k: v.detach().clone().requires_grad_(k in gG) for k, v in G.items() if k.startswith(prefix)}
            first = prev is None
            xe = x_in.detach().clone().requires_grad_(not first)
            with O.storage(dt, **cfg):
                oe = O.unit_forward(sd_u, unit, O._q(xe), True)
            pk = [k for k in sd_u if sd_u[k].requires_grad]
            ge = torch.autograd.grad(oe, ([] if first else [xe]) + [sd_u[k] for k in pk], d_out[key])
            e_act = float((oe - taps[key]).abs().max() / taps[key].abs().max())
            e_dx = 0.0 if first else l2(ge[0], d_out[prev])
            e_p = max(l2(a, gG[k]) for k, a in zip(pk, ge[0 if first else 1:]))
            per.append((key, e_act, e_dx, e_p))
            worst = [max(a, b) for a, b in zip(worst, (e_act, e_dx, e_p))]
            x_in, prev = taps[key], key
        print("%-30s %.3e | %.3e | %.3e" % ((name,) + tuple(worst)))
        if os.environ.get("VERBOSE"):
            for r in per:
                print("    %-16s %.3e %.3e %.3e" % r)


if __name__ == "__main__":
    main()
