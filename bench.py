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
        name = "r2_launch_shares.txt" if os.path.exists(os.path.join(ROOT, "profiles", "r2_launch_shares.txt")) else "r1_launch_shares.txt"
        txt = open(os.path.join(ROOT, "profiles", name)).read()
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


FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12        # nominal: 148 SMs x 128 FMA lanes x 2 x max SM clock = 74.4
MUFU_PEAK_TOPS = 148 * 16 * 1.965e9 / 1e12               # nominal: 16 ex2 / clk / SM = 4.65 T ex2/s


def _stats(ts):
    ts = sorted(ts)
    return {"med": ts[len(ts) // 2], "p10": ts[len(ts) // 10], "p90": ts[(len(ts) * 9) // 10]}


def tfops_micro(dev, iters=100, warmup=20):
    """BASELINE configs 3 and 4 (+ the batch-256 and in-model sizes) and config 1, per SURVEY.md 8(d): CUDA events,
    L2 read-flushed between iterations, >= 20 warm-up + >= 100 timed launches (median / p10 / p90 in us); next to
    each op the reference's OWN kernel recompiled for sm_100a (`ref_gpu_us`, oracle/_ref/libtfops_ref_gpu.so: the
    kernel to beat), the reference's OWN single-threaded CPU function (`cpu_baseline`, oracle/_ref/libtfops_ref_cpu.so)
    and the roofline that bounds the op (FP32 pipe for nn_distance, MUFU for approxmatch, HBM for cost / grad)."""
    import ctypes
    from monopsr_b200 import lib as mlib
    from oracle import refgpu, tfops
    L = mlib.load()
    peaks, which = read_peaks()
    hbm = peaks["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sp = mlib.stream_ptr
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def t(fn, it=iters, wu=warmup):
        for _ in range(wu):
            fn()
        torch.cuda.synchronize()
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
        return _stats(ts)

    def cpu_time(fn):
        t0 = time.perf_counter()
        fn()
        return (time.perf_counter() - t0) * 1e6

    have_ref_gpu, have_ref_cpu = refgpu.available(), tfops.ref_available()
    out = {"protocol": {"iters": iters, "warmup": warmup, "l2": "read-flushed before every timed launch",
                        "hbm_peak_gbs": hbm, "hbm_peak_source": "MEASURED_PEAKS.json (%s)" % which,
                        "fp32_peak_tflops": FP32_PEAK_TFLOPS, "mufu_peak_tex2s": MUFU_PEAK_TOPS,
                        "fp32_mufu_peak_source": "nominal (SM count x lanes x max SM clock)",
                        "ref_gpu": "reference .cu recompiled -arch=sm_100a" if have_ref_gpu else "unavailable",
                        "cpu_baseline": "reference CPU functions, 1 thread" if have_ref_cpu else "oracle port, 1 thread"}}

    # ---------------- nn_distance: cfg3 (32 x 2048^2), the north star's batch 256, the in-model 32 x 2304^2
    for tag, b, n in (("cfg3_nn_distance_b32_n2048", 32, 2048), ("nn_distance_b256_n2048", 256, 2048),
                      ("nn_distance_b32_n2304_inmodel", 32, 2304)):
        g = torch.Generator(device="cpu").manual_seed(100)
        x, y = torch.randn(b, n, 3, generator=g), torch.randn(b, n, 3, generator=g)
        xd, yd = x.to(dev), y.to(dev)
        d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, n, device=dev)
        i1 = torch.empty(b, n, device=dev, dtype=torch.int32)
        i2 = torch.empty(b, n, device=dev, dtype=torch.int32)
        g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev)
        one = torch.ones(b, n, device=dev)
        fw = lambda: L.mpb_nn_distance(b, n, xd.data_ptr(), n, yd.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(),
                                       i2.data_ptr(), sp())
        bw = lambda: L.mpb_nn_distance_grad(b, n, xd.data_ptr(), n, yd.data_ptr(), one.data_ptr(), i1.data_ptr(),
                                            one.data_ptr(), i2.data_ptr(), g1.data_ptr(), g2.data_ptr(), sp())
        r = {"fwd_us": t(fw), "grad_us": t(bw)}
        flop = 2.0 * b * n * n * 8                      # 2 directions x 8 FP32 ops per pair
        bytes_fwd = b * 2 * n * 20                      # b(n+m)(12 in + 8 out)
        us = r["fwd_us"]["med"]
        r["roofline"] = {"bound": "fp32", "achieved": flop / (us * 1e-6) / 1e12, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": flop / (us * 1e-6) / 1e12 / FP32_PEAK_TFLOPS,
                         "note": "FP32-issue bound by construction (820 FLOP/B); the north star's 70%-of-HBM target "
                                 "is unreachable with algorithmic bytes"}
        r["hbm"] = {"algorithmic_bytes": bytes_fwd, "achieved_gbs": bytes_fwd / (us * 1e-6) / 1e9,
                    "frac": bytes_fwd / (us * 1e-6) / 1e9 / hbm}
        if have_ref_gpu:
            r["ref_gpu_us"] = {"fwd": t(lambda: refgpu.nn_distance_launch(xd, yd, d1, i1, d2, i2), it=20, wu=3),
                               "grad": t(lambda: getattr(refgpu.lib(), refgpu._NNG)(b, n, P(xd), n, P(yd), P(one), P(i1), P(one),
                                                                                  P(i2), P(g1), P(g2)), it=20, wu=3)}
            fw()
        if b == 32:                                      # reference CPU nnsearch + gradient loops, one thread, full batch
            xn, yn = x.numpy(), y.numpy()
            order = "ref" if have_ref_cpu else "cpu"
            res = {}
            cf = cpu_time(lambda: res.update(o=tfops.nn_distance(xn, yn, order)))
            o = res["o"]
            ones = np.ones((b, n), np.float32)
            cg = cpu_time(lambda: tfops.nn_distance_grad(xn, yn, ones, o[1], ones, o[3]))
            r["cpu_baseline"] = {"fwd_us": cf, "grad_us": cg, "cores": 1, "kind": "reference" if have_ref_cpu else "port",
                                 "sample": "the whole batch, once"}
        out[tag] = r

    # ---------------- approxmatch + match_cost + match_cost_grad: cfg4 (32 x 1024^2) and the in-model 32 x 2304^2
    for tag, b, n in (("cfg4_approxmatch_b32_n1024", 32, 1024), ("approxmatch_b32_n2304_inmodel", 32, 2304)):
        g = torch.Generator(device="cpu").manual_seed(200)
        x, y = torch.randn(b, n, 3, generator=g), torch.randn(b, n, 3, generator=g)
        xd, yd = x.to(dev), y.to(dev)
        mt = torch.empty(b, n, n, device=dev)
        cost = torch.empty(b, device=dev)
        g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev)
        am = lambda: L.mpb_approxmatch(b, n, n, xd.data_ptr(), yd.data_ptr(), mt.data_ptr(), None, sp())
        mc = lambda: L.mpb_matchcost(b, n, n, xd.data_ptr(), yd.data_ptr(), mt.data_ptr(), cost.data_ptr(), sp())
        mg = lambda: L.mpb_matchcostgrad(b, n, n, xd.data_ptr(), yd.data_ptr(), mt.data_ptr(), g1.data_ptr(), g2.data_ptr(), sp())
        big = n > 1024
        r = {"match_us": t(am, it=20 if big else iters, wu=3 if big else warmup), "cost_us": t(mc), "grad_us": t(mg)}
        mbytes = 4.0 * b * n * n
        nexp = 30.0 * b * n * n                           # 10 levels x 3 sweeps
        r["roofline_match"] = {"bound": "mufu", "achieved": nexp / (r["match_us"]["med"] * 1e-6) / 1e12, "peak": MUFU_PEAK_TOPS,
                               "unit": "T ex2/s", "frac": nexp / (r["match_us"]["med"] * 1e-6) / 1e12 / MUFU_PEAK_TOPS,
                               "hbm_gbs_on_the_single_match_write": mbytes / (r["match_us"]["med"] * 1e-6) / 1e9}
        for k in ("cost", "grad"):
            gbs = mbytes / (r[k + "_us"]["med"] * 1e-6) / 1e9
            r["roofline_" + k] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                  "algorithmic_bytes": mbytes}
        if have_ref_gpu and not big:
            rl = refgpu.lib()
            temp = torch.empty(32 * 4 * n, device=dev)
            r["ref_gpu_us"] = {
                "match": t(lambda: getattr(rl, refgpu._AM)(b, n, n, P(xd), P(yd), P(mt), P(temp)), it=5, wu=1),
                "cost": t(lambda: getattr(rl, refgpu._MC)(b, n, n, P(xd), P(yd), P(mt), P(cost)), it=10, wu=2),
                "grad": t(lambda: getattr(rl, refgpu._MCG)(b, n, n, P(xd), P(yd), P(mt), P(g1), P(g2)), it=10, wu=2)}
            am()
        if not big:                                       # reference CPU functions, one thread, bounded sample
            bs = 2
            xn, yn = x.numpy()[:bs], y.numpy()[:bs]
            order = "ref" if have_ref_cpu else "cpu"
            res = {}
            c_m = cpu_time(lambda: res.update(m=tfops.approx_match(xn, yn, order))) * (b / bs)
            c_c = cpu_time(lambda: tfops.match_cost(xn, yn, res["m"], order)) * (b / bs)
            c_g = cpu_time(lambda: tfops.match_cost_grad(xn, yn, res["m"], order)) * (b / bs)
            r["cpu_baseline"] = {"match_us": c_m, "cost_us": c_c, "grad_us": c_g, "cores": 1,
                                 "kind": "reference" if have_ref_cpu else "port",
                                 "sample": "%d of %d batch elements, scaled by %d" % (bs, b, b // bs)}
        out[tag] = r

    # ---------------- cfg1: ONE crop -> ResNet-101 block3 features -> centroid-z head (plumbing check; CPU = the restated graph)
    try:
        from monopsr_b200.core import model_spec as ms
        from monopsr_b200.core.engine import Engine
        from oracle import network as onet
        P1, S1 = ms.init_params(0), ms.synthetic_sample(0)
        eng = Engine(dev, params=P1)
        eng.set_inputs(S1)
        Tc = eng.towers[ms.ENCODERS[0]]

        def crop_tower():
            eng._tower_fwd(Tc, eng.inputs["rgb_crops"])
        eng.forward(train=False)
        out["cfg1_one_crop_features"] = {
            "gpu_crop_tower_us_per_32_crops": t(crop_tower, it=20, wu=3),
            "note": "the crop encoder alone on the 32 crops of one sample (the engine's granularity); per crop = / 32"}
        torch.set_num_threads(cpu_threads())
        Pt = onet.to_torch(P1, torch.float32)
        x1 = torch.from_numpy(np.asarray(S1["rgb_crops"])[:1]).float()
        with torch.no_grad():
            onet.resnet101_block3(x1, Pt, ms.ENCODERS[0])
            c1 = cpu_time(lambda: onet.resnet101_block3(x1, Pt, ms.ENCODERS[0]))
        out["cfg1_one_crop_features"]["cpu_baseline"] = {"us_per_crop": c1, "cores": cpu_threads(), "kind": "port",
                                                         "sample": "1 crop, crop encoder (restated TF1 graph, torch-CPU fp32)"}
        del eng
    except Exception as ex:      # the plumbing configuration must not cost the headline line
        out["cfg1_one_crop_features"] = {"error": repr(ex)[:200]}

    # ---------------- SURVEY 8f rank 2 (the step before the path): targets from a 375 x 1242 depth map + 32 masks
    from monopsr_b200.core import model_spec as ms, targets as mtg
    S = ms.synthetic_sample(0)
    rng = np.random.RandomState(0)
    H, W = 375, 1242
    depth = torch.from_numpy(rng.uniform(3, 60, (H, W)).astype(np.float32)).to(dev)
    masks_h = rng.rand(ms.NUM_BOXES, H, W) < 0.6
    masks = torch.from_numpy(masks_h).to(dev)
    args = [torch.from_numpy(S[k]).to(dev) for k in ("boxes_2d", "boxes_3d", "est_view_angs", "cam_p")]
    img = torch.from_numpy(rng.randint(0, 256, (H, W, 3)).astype(np.uint8)).to(dev)
    bn = torch.from_numpy(S["boxes_2d_norm"]).to(dev)
    from oracle import targets as otg          # the CPU restatement, timed once beside it (a reported baseline)
    out["gt_targets"] = {"us": t(lambda: mtg.gt_maps_from_depth(depth, masks, *args, dev), it=20, wu=3),
                         "image_inputs_us": t(lambda: mtg.image_inputs(img, bn, dev), it=20, wu=3),
                         "cpu_oracle_us": cpu_time(lambda: otg.gt_maps(S["boxes_2d"], S["boxes_3d"], masks_h, depth.cpu().numpy(),
                                                                       S["est_view_angs"], S["cam_p"]))}
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
        finish(world)
        return

    # ---- roofline of the dominant kernel family (tcgen05 implicit GEMM):
    #   frac      = WHOLE-STEP fraction: the algorithmic 2*M*N*K of one step (full-tap convention; the extra
    #               products of the h3 / x3 forward are NOT counted) / ms_per_step / dense-tf32 peak
    #   alone     = the same launches (same arguments) replayed back to back as one CUDA graph, timed live with CUDA
    #               events: a diagnostic only -- inside the step launches overlap (2 CTAs/SM, 2-4 streams), so the step
    #               is shorter than that sum
    #   tensor_busy_frac = issued MMA work / peak: forward k-blocks cost 1.5x (h3: six kind::f16 MMAs = 1.5 tf32 k-blocks)
    #               or 3x (x3) the single pass -- how busy the tensor pipe is, as opposed to how much useful work it did
    peaks, which = read_peaks()
    roof = eng.gemm_only_roofline(flush)
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0      # dense tf32 = half the bf16 rate; sustained (long step)
    step_s = total_ms / args.steps / 1e3
    gflop = roof["gflop"]
    whole = gflop / 1e3 / step_s
    fwd_factor = {"h3": 1.5, "x3": 3.0, "tf32": 1.0}[eng.precision]
    issued = gflop * (2.0 + fwd_factor) / 3.0             # forward = 1/3 of the step's GEMM work, backward 2/3
    kernel = {"h3": "tc_gemm_tma_kernel<.., H3> forward (tcgen05 kind::f16 on fp16 hi/lo split operands, 3 products) + "
                    "tc_gemm_tma_kernel backward (kind::tf32)",
              "x3": "tc_gemm_x3_kernel forward (3xTF32) + tc_gemm_tma_kernel backward (kind::tf32)",
              "tf32": "tc_gemm_tma_kernel (tcgen05 kind::tf32)"}[eng.precision]
    roofline = {"bound": "tensor", "achieved": whole, "peak": tf32_peak, "unit": "TFLOP/s", "frac": whole / tf32_peak,
                "traffic": ncu_traffic_per_launch(),
                "traffic_source": "profiles/r2_launch_shares.txt: ncu dram__bytes_read+write of the GEMM launches of one "
                                  "step / launches (bytes per launch; ncu flushes caches between launches)",
                "kernel": kernel, "launches_per_step": roof["launches"], "launch_kinds": roof["kinds"],
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (%s)" % which,
                "algorithmic_gflop_per_step": gflop,
                "alone": {"ms": roof["ms"], "tflops": roof["tflops"], "frac": roof["tflops"] / tf32_peak,
                          "note": "the GEMM launches replayed serially on one stream; exceeds ms_per_step by design"},
                "tensor_busy_frac": issued / 1e3 / step_s / tf32_peak}

    precision = {"h3": "fp16 hi/lo split forward (3 products, fp32 accumulate: every forward output within 1e-3 of the "
                       "fp32 graph, measured <= 4e-5), single-pass tf32 backward",
                 "x3": "3xTF32 forward, single-pass tf32 backward",
                 "tf32": "single-pass tf32 (FAST mode: misses the 1e-3 parity bar on the decoder maps)"}[eng.precision]
    line = {
        "metric": "instance-crops/sec fwd+bwd", "value": value, "unit": "crops/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"h3": "f16x3+tf32", "x3": "tf32x3+tf32", "tf32": "tf32"}[eng.precision], "data": "synthetic",
        "config": {"workload": "cfg2: 32 crops 48x48x3 + 160x608 full image per GPU, monopsr_model_000 "
                               "fwd+bwd+train-op (clip+Adam+EMA)", "crops_per_gpu": CROPS_PER_SAMPLE,
                   "precision": precision, "precision_mode": eng.precision,
                   "points_per_instance": 2304, "parallelism": "dp%d" % world, "l2": "flushed between timed steps",
                   "weights": "random-init, seed 0"},
        "e2e": {"value": e2e_value, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(getattr(eng, "launches_per_step", 0)) * args.steps * 2,
        "clocks": clocks,
        "roofline": roofline,
    }
    eng.check_overflow()
    if not args.no_ops and world == 1:        # the op configurations are single-GPU measurements
        line["ops"] = tfops_micro(dev)
    if not args.no_cpu_baseline and world == 1:
        threads = cpu_threads()
        ts = cpu_reference_step(P, S, threads, 1)
        line["cpu_baseline"] = {"value": CROPS_PER_SAMPLE / ts[0], "unit": "crops/s", "cores": threads, "kind": "port",
                                "sample": "1 full step of 32 crops (restated TF1 graph, torch-CPU fp32)"}
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """multi-rank exit: tearing an NCCL communicator down while CUDA graphs that ran on it are still alive can block
    for minutes (seen on the 2-GPU box); the work is done and printed, so leave without the teardown"""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
