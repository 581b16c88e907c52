#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MonoPSR hot path.

Metric (BASELINE.json): instance-crops/sec, forward+backward(+train-op), on BASELINE config 2:
one sample = 32 synthetic 48x48x3 crops + one 160x608 full image + 2304-pt targets, the full
``monopsr_model_000`` step (two ResNet-101 towers, squash, map decoder, FC heads, losses,
per-variable clip + Adam + EMA).  N GPUs = N samples per step (weak scaling), one NCCL
all-reduce of the 401 MB fp32 gradient arena per step.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the restated reference graph on the host CPU cores

Prints ONE JSON line (see the keys at the bottom).  ``value`` is timed with inputs resident
in HBM; ``e2e`` goes through the public API with HOST inputs (H2D of the sample and D2H of
the losses inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CROPS_PER_SAMPLE = 32
FLOP_PER_CROP_FWD_BWD = 68.5e9      # 22.84 GFLOP fwd x 3


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, gpu=0):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        self.gpu = gpu
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.15)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def ncu_traffic_per_launch():
    """DRAM bytes per launch of the GEMM family from the committed ncu launch list (None if absent)."""
    import re
    try:
        txt = open(os.path.join(ROOT, "profiles", "r1_launch_shares.txt")).read()
        m = re.search(r"([0-9.]+) MB per launch", txt)
        return float(m.group(1)) * 1e6 if m else None
    except OSError:
        return None


def cpu_threads():
    # torch's CPU convolutions on 12x12 maps stop scaling (and then regress) beyond a few dozen
    # threads; use what helps and report the count actually used
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_reference_step(P, S, threads, steps=1):
    """The restated reference graph (oracle/network.py, fp32) fwd+bwd on the host cores.
    TensorFlow 1.8 itself cannot be installed here."""
    from oracle import network as onet
    torch.set_num_threads(threads)
    Pt = onet.to_torch(P, torch.float32)
    for v in Pt.values():
        v.requires_grad_(v.dtype == torch.float32)
    St = onet.to_torch(S, torch.float32)
    ts = []
    for _ in range(steps):
        t0 = time.time()
        out, _ = onet.forward(Pt, St, train=True)
        _, tot = onet.loss(out, St)
        tot.backward()
        for v in Pt.values():
            v.grad = None
        ts.append(time.time() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from monopsr_b200.core import model_spec as ms
    threads = cpu_threads()
    P, S = ms.init_params(0), ms.synthetic_sample(0)
    steps = max(1, min(args.steps, 3))       # one step = 32 crops = ~10-20 s of CPU work
    cpu_reference_step(P, S, threads, 1) if args.warmup > 0 else None
    ts = cpu_reference_step(P, S, threads, steps)
    ms_per_step = float(np.mean(ts)) * 1e3
    value = CROPS_PER_SAMPLE / (ms_per_step / 1e3)
    line = {
        "impl": "reference", "metric": "instance-crops/sec fwd+bwd", "value": value, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 32 crops 48x48x3 + 160x608 full image, monopsr_model_000 fwd+bwd",
                   "note": "restated TF1 graph in torch-CPU fp32 (TensorFlow 1.8 not installable); bounded sample"},
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": threads, "kind": "port",
                         "sample": "%d full step(s) of 32 crops" % steps},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def tfops_micro(dev):
    """nn_distance / approxmatch us per batch (BASELINE configs 3 and 4), CUDA events, L2 flushed."""
    from monopsr_b200 import lib as mlib
    L = mlib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def t(fn, it=10, wu=3):
        for _ in range(wu):
            fn()
        ts = []
        for _ in range(it):
            flush.sum()      # READ-flush of L2: a write-flush leaves 126 MB of dirty lines whose write-back
                             # would be charged to the HBM-bound kernel timed next
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        return float(np.median(ts))

    out = {}
    g = torch.Generator(device="cpu").manual_seed(100)
    for b in (32, 256):
        n = 2048
        x, y = torch.randn(b, n, 3, generator=g).to(dev), torch.randn(b, n, 3, generator=g).to(dev)
        d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, n, device=dev)
        i1 = torch.empty(b, n, device=dev, dtype=torch.int32)
        i2 = torch.empty(b, n, device=dev, dtype=torch.int32)
        g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev)
        one = torch.ones(b, n, device=dev)
        fw = lambda: L.mpb_nn_distance(b, n, x.data_ptr(), n, y.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(),
                                       i2.data_ptr(), mlib.stream_ptr())
        out["nn_distance_fwd_us_b%d" % b] = t(fw)
        bw = lambda: L.mpb_nn_distance_grad(b, n, x.data_ptr(), n, y.data_ptr(), one.data_ptr(), i1.data_ptr(),
                                            one.data_ptr(), i2.data_ptr(), g1.data_ptr(), g2.data_ptr(), mlib.stream_ptr())
        out["nn_distance_grad_us_b%d" % b] = t(bw)
        # algorithmic HBM bytes: b*(n+m)*(12+8) fwd
        out["nn_distance_fwd_hbm_gbs_b%d" % b] = b * 2 * n * 20 / (out["nn_distance_fwd_us_b%d" % b] * 1e-6) / 1e9
    g = torch.Generator(device="cpu").manual_seed(200)
    b, n = 32, 1024
    x, y = torch.randn(b, n, 3, generator=g).to(dev), torch.randn(b, n, 3, generator=g).to(dev)
    mt = torch.empty(b, n, n, device=dev)
    cost = torch.empty(b, device=dev)
    g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev)
    out["approxmatch_us"] = t(lambda: L.mpb_approxmatch(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), None,
                                                        mlib.stream_ptr()), it=5, wu=2)
    out["matchcost_us"] = t(lambda: L.mpb_matchcost(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), cost.data_ptr(),
                                                    mlib.stream_ptr()))
    out["matchcostgrad_us"] = t(lambda: L.mpb_matchcostgrad(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(),
                                                            g1.data_ptr(), g2.data_ptr(), mlib.stream_ptr()))
    # SURVEY 8f rank 2 (the step before the path): targets from a 375 x 1242 depth map + 32 masks, inputs from the image
    from monopsr_b200.core import model_spec as ms, targets as mtg
    import time
    S = ms.synthetic_sample(0)
    rng = np.random.RandomState(0)
    H, W = 375, 1242
    depth = torch.from_numpy(rng.uniform(3, 60, (H, W)).astype(np.float32)).to(dev)
    masks_h = rng.rand(ms.NUM_BOXES, H, W) < 0.6
    masks = torch.from_numpy(masks_h).to(dev)
    args = [torch.from_numpy(S[k]).to(dev) for k in ("boxes_2d", "boxes_3d", "est_view_angs", "cam_p")]
    out["gt_targets_us"] = t(lambda: mtg.gt_maps_from_depth(depth, masks, *args, dev))
    img = torch.from_numpy(rng.randint(0, 256, (H, W, 3)).astype(np.uint8)).to(dev)
    bn = torch.from_numpy(S["boxes_2d_norm"]).to(dev)
    out["image_inputs_us"] = t(lambda: mtg.image_inputs(img, bn, dev))
    from oracle import targets as otg          # the CPU restatement, timed once beside it (a reported baseline)
    t0 = time.perf_counter()
    otg.gt_maps(S["boxes_2d"], S["boxes_3d"], masks_h, depth.cpu().numpy(), S["est_view_angs"], S["cam_p"])
    out["gt_targets_cpu_oracle_us"] = (time.perf_counter() - t0) * 1e6
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ops", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from monopsr_b200 import lib as mlib
    from monopsr_b200.core import model_spec as ms
    from monopsr_b200.core.engine import Engine

    P = ms.init_params(0)
    S = ms.synthetic_sample(rank)                    # one sample (32 crops) per GPU
    eng = Engine(dev, params=P)
    eng.set_inputs(S)
    Spin = {k: torch.as_tensor(np.asarray(v)).pin_memory() for k, v in S.items()}
    h2d = int(sum(v.numel() * v.element_size() for v in Spin.values()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up (also captures the CUDA graph)
    for _ in range(args.warmup):
        eng.train_step()
    barrier()

    # ---- timed region A: inputs resident in HBM
    with ClockSampler(local) as clk:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for s, e in ev:
            flush.zero_()                              # L2 flush between timed iterations (outside the events)
            s.record()
            eng.train_step()
            e.record()
        barrier()
        step_ms = [s.elapsed_time(e) for s, e in ev]
        total_ms = float(sum(step_ms))
        # ---- timed region B: end to end through the public API, host inputs + loss read-back
        barrier()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        d2h = 0
        for s, e in ev2:
            flush.zero_()
            s.record()
            eng.train_step(Spin)                       # pinned host -> device copies inside
            lv = eng.h["losses"].cpu()                 # device -> host read of the step's losses
            e.record()
            d2h = lv.numel() * lv.element_size()
        barrier()
        e2e_ms = float(sum(s.elapsed_time(e) for s, e in ev2))
    clocks = clk.summary()
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])

    crops = CROPS_PER_SAMPLE * world * args.steps
    value = crops / (total_ms / 1e3)
    e2e_value = crops / (e2e_ms / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (tcgen05 implicit GEMM), measured live:
    # replay ONLY the tc_gemm launches of one step (same arguments) as a CUDA graph and time it.
    peaks, which = read_peaks()
    roof = eng.gemm_only_roofline(flush)
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0      # dense tf32 = half the bf16 rate; sustained (long step)
    roofline = {"bound": "tensor", "achieved": roof["tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": roof["tflops"] / tf32_peak, "traffic": ncu_traffic_per_launch(),
                "traffic_source": "profiles/r1_launch_shares.txt: ncu dram__bytes_read+write of the 592 GEMM launches "
                                  "of one step / 592 (bytes per launch; ncu flushes caches between launches)",
                "kernel": "tc_gemm_tma_kernel (tcgen05 kind::tf32), %d launches/step, %.2f ms alone vs %.2f ms/step" %
                          (roof["launches"], roof["ms"], total_ms / args.steps),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (%s)" % which,
                "algorithmic_gflop_per_step": roof["gflop"],
                "whole_step_frac": (value * FLOP_PER_CROP_FWD_BWD / 1e12) / (tf32_peak * world)}

    line = {
        "metric": "instance-crops/sec fwd+bwd", "value": value, "unit": "crops/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
        "config": {"workload": "cfg2: 32 crops 48x48x3 + 160x608 full image per GPU, monopsr_model_000 "
                               "fwd+bwd+train-op (clip+Adam+EMA)", "crops_per_gpu": CROPS_PER_SAMPLE,
                   "precision": "3xTF32 forward (MPB_PRECISION=x3), single-pass tf32 backward" if getattr(eng, "x3", False)
                                else "single-pass tf32",
                   "points_per_instance": 2304, "parallelism": "dp%d" % world, "l2": "flushed between timed steps",
                   "weights": "random-init, seed 0"},
        "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(getattr(eng, "launches_per_step", 0)) * args.steps * 2,
        "clocks": clocks,
        "roofline": roofline,
    }
    if not args.no_ops:
        line["ops"] = tfops_micro(dev)
    if not args.no_cpu_baseline and world == 1:
        threads = cpu_threads()
        ts = cpu_reference_step(P, S, threads, 1)
        line["cpu_baseline"] = {"value": CROPS_PER_SAMPLE / ts[0], "unit": "crops/s", "cores": threads, "kind": "port",
                                "sample": "1 full step of 32 crops (restated TF1 graph, torch-CPU fp32)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
