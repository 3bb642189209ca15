#!/usr/bin/env python
"""bench.py - MCD-step throughput of the B200-native path (and of the reference's CPU path).

    python bench.py --gpus N --steps K --warmup W            # our arm; N>1 under torchrun, one rank per GPU
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

Metric (BASELINE.json): "MCD-step images/s" = source/target image PAIRS per second through one full MCD
iteration (phase A + B + num_k=4 x C, adapt_trainer.py:162-212) of DRN-D-38, input_ch=6, n_class=41,
480x640, SGD(lr 1e-3, momentum .9, wd 2e-5), random-init weights, synthetic N(0,1) images and uniform labels.
A "step" is one MCD iteration over a batch of `--batch` pairs per GPU (weak scaling).

One JSON line is printed by rank 0; see README/DESIGN.md for the keys (value, e2e, roofline, cpu_baseline,
clocks, gpu_launches).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multichannel-semseg-with-uda_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_CLASS = 41
METRIC = "MCD-step images/s, DRN-D-38 6ch 480x640"
UNIT = "image pairs/s"
FULL = (480, 640)
# algorithmic conv FLOPs per image pair and MCD iteration: 7 fwd + 5 bwd of G
G_FWD_GF = 260.33
ITER_TFLOP_PER_PAIR = (7 * G_FWD_GF + 5 * (2 * G_FWD_GF - 2.89)) / 1e3


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"],
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json, written by scripts/ncu_summarise.py); None when no capture covers it."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    ent = json.load(open(path)).get(kernel.split(" [")[0])
    if not ent or not ent.get("dram_bytes_per_launch"):
        return None
    v = ent["dram_bytes_per_launch"]
    return sum(v) / len(v)


# --------------------------------------------------------------------------------------------------
def synth(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(batch, 6, *size, generator=g)
    tgt = torch.randn(batch, 6, *size, generator=g)
    lbl = torch.randint(0, N_CLASS, (batch, *size), generator=g)
    return src, lbl, tgt


def run_reference(args):
    """The reference's algorithm on the host cores: the oracle port (fp32, stock torch CPU ops), all threads.
    Each step = one MCD iteration on a bounded sample of the workload (see `sample`)."""
    from oracle import mcd_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # choose the largest sample whose projected run time stays within ~4 minutes
    budget_s, steps = 240.0, args.steps + args.warmup
    G = O.init_seg_base("drn_d_38", 6, N_CLASS, torch.Generator().manual_seed(0))
    F1, F2 = O.init_head(N_CLASS, gen=torch.Generator().manual_seed(1)), O.init_head(N_CLASS, gen=torch.Generator().manual_seed(2))
    w = O.class_weight(N_CLASS)
    og, of = O.SGD(), O.SGD()

    def one(size):
        src, lbl, tgt = synth(1, size, 3)
        t0 = time.perf_counter()
        O.mcd_step_early(G, F1, F2, src, lbl, tgt, w, og, of, num_k=4)
        return time.perf_counter() - t0

    t_small = min(one((120, 160)), one((120, 160)))
    size = (120, 160)
    for cand, scale in (((480, 640), 16.0), ((240, 320), 4.0)):
        if t_small * scale * steps <= budget_s:
            size = cand
            break
    for _ in range(args.warmup):
        one(size)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one(size)
    dt = (time.perf_counter() - t0) / args.steps
    frac = (size[0] * size[1]) / float(FULL[0] * FULL[1])
    value = frac / dt   # 480x640-equivalent pairs per second
    sample = ("1 pair per step at %dx%d (%.4f of a 480x640 pair by pixel count; value is in 480x640-pair "
              "equivalents), fp32, %d torch threads" % (size[0], size[1], frac, cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "early-fusion MCD iteration (A+B+4xC), DRN-D-38 6ch, n_class 41", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_sample():
    """bounded CPU sample for the N=1 line: one MCD iteration of the oracle port on the host cores."""
    from oracle import mcd_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G = O.init_seg_base("drn_d_38", 6, N_CLASS, torch.Generator().manual_seed(0))
    F1, F2 = O.init_head(N_CLASS, gen=torch.Generator().manual_seed(1)), O.init_head(N_CLASS, gen=torch.Generator().manual_seed(2))
    w, og, of = O.class_weight(N_CLASS), O.SGD(), O.SGD()
    size = (240, 320)
    src, lbl, tgt = synth(1, size, 3)
    O.mcd_step_early(G, F1, F2, src, lbl, tgt, w, og, of, num_k=4)   # warm-up
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        O.mcd_step_early(G, F1, F2, src, lbl, tgt, w, og, of, num_k=4)
    dt = (time.perf_counter() - t0) / n
    frac = (size[0] * size[1]) / float(FULL[0] * FULL[1])
    return {"value": frac / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d MCD iterations of 1 pair at %dx%d (%.2f of a 480x640 pair; value in 480x640-pair "
                      "equivalents), fp32 oracle port, %d torch threads, %.2f s/iteration"
                      % (n, size[0], size[1], frac, cores, dt)}


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    from mcd_b200 import abi, ops, parallel
    from mcd_b200.step import MCDStep
    from loss import CrossEntropyLoss2d, get_prob_distance_criterion
    from models.model_util import get_models
    from util import get_class_weight_from_file

    rank, local, world = parallel.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    abi.check(abi.lib().mcd_check_device(local), "mcd_check_device")
    pk = peaks()
    B, size = args.batch, FULL

    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        models = [m.to(dev).train() for m in get_models("drn_d_38", 6, N_CLASS, method="MCD")]
    criterion = CrossEntropyLoss2d(get_class_weight_from_file(N_CLASS).to(dev))
    criterion_d = get_prob_distance_criterion("diff")
    step = MCDStep(models, criterion, criterion_d, num_k=4)

    src_h, lbl_h, tgt_h = [t.pin_memory() for t in synth(B, size, 100 + rank)]
    src_d, lbl_d, tgt_d = src_h.to(dev), lbl_h.to(dev), tgt_h.to(dev)
    h2d = sum(t.numel() * t.element_size() for t in (src_h, lbl_h, tgt_h))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t)
        return ms

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    use_graph = not args.no_graph
    for _ in range(args.warmup):
        step(src_d, lbl_d, tgt_d)
    # ---- roofline instrumentation: one eager iteration with a CUDA-event pair (on the launching stream) around
    #      every convolution launch; the timed region below replays the same kernels from a CUDA graph, where
    #      individual launches cannot be bracketed.
    from mcd_b200 import nn as mcd_nn
    prev_overlap = mcd_nn.set_overlap_wgrad(False)   # serialise dgrad / wgrad so each family is timed alone
    prof = ops.ConvProfiler()
    with prof:
        step(src_d, lbl_d, tgt_d)
    torch.cuda.synchronize()
    mcd_nn.set_overlap_wgrad(prev_overlap)
    fam = prof.summary()
    if use_graph:
        step.capture(src_d, lbl_d, tgt_d, warmup=1)

    def resident_step():
        if use_graph:
            step.graph.replay()        # static device-resident inputs: pure hot-path time
        else:
            step(src_d, lbl_d, tgt_d)

    def e2e_step():
        if use_graph:
            # pinned host -> staging (copy stream, started one step ahead so that it overlaps the running
            # iteration) -> static device buffers -> one graph launch -> losses back to the host.  Every step's
            # H2D copy and D2H read happen inside the timed region.
            c, d = step.replay_prefetched()
            step.prefetch(src_h, lbl_h, tgt_h)
        else:
            c, d = step(src_h.to(dev, non_blocking=True), lbl_h.to(dev, non_blocking=True),
                        tgt_h.to(dev, non_blocking=True))
        return float(c), float(d)      # device -> host read of the step's result

    for _ in range(3):
        resident_step()
    def measure():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        n0 = abi.launch_count()
        t = timed(resident_step, args.steps)
        n = step.launches_per_replay * args.steps if use_graph else abi.launch_count() - n0
        return t, n, (sampler.stop() if rank == 0 else None)

    def rejected(c):
        """thermal / hardware slowdown, or SM clocks far below max with no reason given (a leftover clock lock)"""
        if not c or c.get("sm_mhz") is None:
            return False
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c.get("reasons", []))
        stuck = not c.get("reasons") and c.get("sm_max_mhz") and c["sm_mhz"] < 0.6 * c["sm_max_mhz"]
        return bool(bad) or bool(stuck)

    ms, launches, clocks = measure()
    redo = torch.tensor([1 if (rank == 0 and rejected(clocks)) else 0], device=dev)
    if world > 1:
        torch.distributed.broadcast(redo, src=0)
    if int(redo):        # re-measured ONCE, the first attempt is kept in the line for the record
        first = {"ms_per_step": ms, "clocks": clocks}
        ms, launches, clocks = measure()
        if rank == 0:
            clocks["rejected_first_attempt"] = first
    if use_graph:
        step.prefetch(src_h, lbl_h, tgt_h)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # dominant kernel = the one with the largest summed launch time in the instrumented (serialised) iteration
    dom = max(fam, key=lambda k: fam[k]["ms"]) if fam else None
    roof = None
    if dom:
        ach = fam[dom]["flop"] / (fam[dom]["ms"] * 1e-3) / 1e12
        serial_ms = sum(v["ms"] for v in fam.values())
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["sustained"], "traffic": ncu_traffic(dom),
                "traffic_note": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the launches of this "
                                "kernel in profiles/r01b_ncu_full_conv_bn.csv (512->512 and 256->256 3x3 layers, 22 "
                                "images, forward and dgrad); algorithmic bytes of the 512-channel forward: 221 MB",
                "peak_source": pk["src"] + ", sustained",
                "launches": fam[dom]["n"], "avg_launch_ms": fam[dom]["ms"] / fam[dom]["n"],
                "flop_per_launch": fam[dom]["flop"] / fam[dom]["n"],
                "share_of_conv_time": fam[dom]["ms"] / serial_ms,
                "how": "CUDA-event pair on the launching stream around every convolution launch of one eager, "
                       "fully serialised MCD iteration run inside this process right before the timed region "
                       "(a CUDA-graph replay cannot be bracketed per kernel); algorithmic FLOPs = 2*N*Ho*Wo*Cout*Cin*R*S",
                "kernels": {k: {"ms": round(v["ms"], 3), "tflops": round(v["flop"] / (v["ms"] * 1e-3) / 1e12, 1),
                                "n": v["n"]} for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}
    # ---- extras (per-phase and inference throughput); never allowed to break the main line ----------
    extras = {}
    try:
        step(src_d, lbl_d, tgt_d)           # untimed: re-populates the eager allocator pool after the graph capture
        torch.cuda.synchronize()
        step.phase_events = []
        step(src_d, lbl_d, tgt_d)           # one eager iteration, wgrad overlapped as in the graph
        torch.cuda.synchronize()
        ev, step.phase_events = step.phase_events, None
        ph = {}
        for (n0_, e0_), (n1_, e1_) in zip(ev[:-1], ev[1:]):
            ph[n1_] = e0_.elapsed_time(e1_)
        c_ms = [v for k, v in ph.items() if k.startswith("C")]
        extras["phases_eager_ms"] = {k: round(v, 3) for k, v in ph.items()}
        extras["phases_pairs_per_s"] = {"A": B / (ph["A"] * 1e-3), "B": B / (ph["B"] * 1e-3),
                                        "C": B / (sum(c_ms) / len(c_ms) * 1e-3)}
        extras["phases_note"] = ("device time between CUDA events at the phase boundaries of ONE eager iteration on "
                                 "this rank; B includes the target forward that phase C[0] re-uses, so C0 is backward only")
        # inference (adapt_tester.py:104-124): eval-mode forward of G + both heads, argmax over the 40 valid classes
        import util as mcd_util
        for m in models:
            m.eval()
        mg, mf1, mf2 = models

        def infer():
            with torch.no_grad():
                feat = mg(tgt_d)
                out = mf1(feat) + mf2(feat)
                return mcd_util.predict_labels(out, N_CLASS - 1), mcd_util.calc_entropy(out)
        for _ in range(2):
            infer()
        ms_inf = timed(infer, 3)
        for m in models:
            m.train()
        extras["inference"] = {"images_per_s": B * world / (ms_inf * 1e-3), "ms_per_batch": ms_inf, "batch_per_gpu": B,
                               "what": "eval-mode DRN-D-38 6ch forward + 2 heads + argmax(40 classes) + entropy, eager "
                                       "launches, inputs resident", "algorithmic_gflop_per_image": G_FWD_GF,
                               "tensor_util": G_FWD_GF * 1e-3 * B / (ms_inf * 1e-3) / pk["sustained"]}
    except Exception as exc:      # noqa: BLE001
        extras["error"] = repr(exc)[:300]
    if rank != 0:
        return
    pairs = B * world
    value = pairs / (ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "early-fusion MCD iteration (A+B+4xC), DRN-D-38 input_ch=6 n_class=41 480x640, "
                               "SGD momentum .9 wd 2e-5, random init", "pairs_per_gpu": B, "global_pairs": pairs,
                   "parallelism": "dp%d" % world, "l2": "per-step working set (activations > 1 GB) exceeds the 126 MB L2",
                   "dead_phaseB_backward_skipped": True, "cuda_graph": bool(use_graph)},
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "tensor_util_of_step": ITER_TFLOP_PER_PAIR * B / (ms * 1e-3) / pk["sustained"],
        "algorithmic_tflop_per_pair": ITER_TFLOP_PER_PAIR,
        "extras": extras,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=22,
                    help="image pairs per GPU and step (22 x 40 = 880 pixel tiles of 128 = 5.95 / 11.9 full waves of "
                         "the 74 CTA pairs for the 256- / 512-channel layers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of one CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
