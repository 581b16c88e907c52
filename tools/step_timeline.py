"""Kernel timeline of one captured training step (CUPTI via torch.profiler): per-kernel start/end and stream,
written to gpurun_out/step_kernels.csv, plus a summary of where the step's wall time goes."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200.core import model_spec as ms
from monopsr_b200.core.engine import Engine

dev = torch.device("cuda:0")
eng = Engine(dev, params=ms.init_params(0))
S = ms.synthetic_sample(0)
eng.set_inputs(S)
for _ in range(5):
    eng.train_step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        eng.train_step()
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "step_trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ks.sort(key=lambda e: e["ts"])
print("kernel events:", len(ks))
# split into steps by the optimizer kernel
ends = [i for i, e in enumerate(ks) if "round_copy" in e["name"] or "fold_bn_multi" in e["name"]]
# take the middle third of events as one step
n = len(ks) // 3
step = ks[n:2 * n]
t0 = step[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in step)
print("step span: %.1f us, %d kernels" % (t1 - t0, len(step)))
with open(os.path.join(ROOT, "gpurun_out", "step_kernels.csv"), "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in step:
        f.write("%.3f,%.3f,%s,%s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream", -1), e["name"][:90].replace(",", ";")))
# coverage: time with >= 1 kernel running, and concurrency histogram
pts = []
for e in step:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
cur, last, busy, hist = 0, pts[0][0], 0.0, {}
for t, d in pts:
    if t > last:
        hist[cur] = hist.get(cur, 0.0) + (t - last)
        if cur > 0:
            busy += t - last
        last = t
    cur += d
print("busy (>=1 kernel): %.1f us; idle gaps: %.1f us" % (busy, (t1 - t0) - busy))
print("concurrency histogram (us):", {k: round(v, 1) for k, v in sorted(hist.items())})
tot = {}
for e in step:
    k = e["name"].split("(")[0][:60]
    tot[k] = tot.get(k, 0.0) + e["dur"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print("%9.1f us  %s" % (v, k))
