"""torch.profiler (CUPTI) kernel timeline of one MCD iteration: per-kernel start/duration/stream -> compact CSV
(gpurun_out/trace_kernels.csv) for overlap / idle-gap analysis."""
import os, sys, json, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multichannel-semseg-with-uda_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from mcd_b200.step import MCDStep
from loss import CrossEntropyLoss2d, get_prob_distance_criterion
from models.model_util import get_models
from util import get_class_weight_from_file
warnings.simplefilter("ignore")
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
torch.manual_seed(0)
models = [m.to(dev).train() for m in get_models("drn_d_38", 6, 41)]
step = MCDStep(models, CrossEntropyLoss2d(get_class_weight_from_file(41).to(dev)), get_prob_distance_criterion("diff"))
src, lbl, tgt = [t.to(dev) for t in bench.synth(B, (480, 640), 1)]
for _ in range(3):
    step(src, lbl, tgt)
if graph:
    step.capture(src, lbl, tgt, warmup=1)
    run = lambda: step.graph.replay()
else:
    run = lambda: step(src, lbl, tgt)
run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/trace.json")
ev = json.load(open("gpurun_out/trace.json"))["traceEvents"]
rows = [(e["ts"], e["dur"], e["args"].get("stream", -1), e["name"][:60].replace(",", ";")) for e in ev
        if e.get("cat") == "kernel"]
rows.sort()
with open("gpurun_out/trace_kernels.csv", "w") as f:
    f.write("ts_us,dur_us,stream,name\n")
    for r in rows:
        f.write("%.3f,%.3f,%s,%s\n" % r)
os.remove("gpurun_out/trace.json")
print("kernels", len(rows), "span_ms", (rows[-1][0] + rows[-1][1] - rows[0][0]) / 1e3)
