#!/usr/bin/env python
"""bench.py — SDXL-VAE training images/sec at 512^2 with tracking on (BASELINE.json metric).

One step = the reference's hot loop (src/train.py:283-330) on one synthetic batch:
  forward through SDXLVAEWrapper (encode -> sample -> decode) with the ActivityMonitor's three
  fonts_nudge target layers tracked every forward, loss (train.py:289-291), backward (DDP gradient
  all-reduce when N > 1), clip_grad_norm_(1.0), AdamW step, zero_grad, and at their configured cadence
  monitor.step / classifier.classify / handler.intervene (track_interval 20, intervention_interval 10).
Workload: configs[3] experiment_fonts_nudge adapted as BASELINE.json states — synthetic 512^2 glyph-like
images, random-init SDXL-VAE (seed 42), bf16 weights, B = 8 per GPU, dead channels planted by gamma=1e-3.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res 512] [--batch 8] [--impl reference]
N > 1: launched by torchrun (one rank per GPU, NCCL), weak scaling (per-GPU batch fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SDXL-VAE train images/sec at 512^2 (tracking on)"


def metric_name(res):
    return METRIC if res == 512 else f"SDXL-VAE train images/sec at {res}^2 (tracking on)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--track-interval", type=int, default=20)
    ap.add_argument("--graph", action="store_true", help="replay forward+loss+backward from one CUDA graph (GraphedVAEStep); "
                    "default is the eager per-op path that unchanged train.py gets (DDP for N>1)")
    ap.add_argument("--fused-adamw", action="store_true", help="torch.optim.AdamW(fused=True) instead of the constructor call "
                    "of train.py:184-187 (about 2.5 ms per step faster at 512^2; SURVEY 8f next-item 2)")
    ap.add_argument("--quick", action="store_true", help="profiling aid: warm-up as given, no e2e/roofline/cpu legs")
    ap.add_argument("--kernel-table", default="", help="profiling aid: after the timed run, trace 2 more steps with "
                    "torch.profiler (CUPTI) and write the per-kernel device-time table to this file")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ FLOP model
def conv_layers(R):
    """(Cin, Cout, k, out_h, kind) of every conv of the SDXL VAE for an RxR input.
    kind: "s1" stride-1 conv, "down" Downsample2D (pad (0,1,0,1), stride 2), "up" Upsample2D (nearest x2 + conv)."""
    L = []
    def res(ci, co, h):
        L.extend([(ci, co, 3, h, "s1"), (co, co, 3, h, "s1")])
        if ci != co:
            L.append((ci, co, 1, h, "s1"))
    L.append((3, 128, 3, R, "s1"))
    res(128, 128, R); res(128, 128, R); L.append((128, 128, 3, R // 2, "down"))
    res(128, 256, R // 2); res(256, 256, R // 2); L.append((256, 256, 3, R // 4, "down"))
    res(256, 512, R // 4); res(512, 512, R // 4); L.append((512, 512, 3, R // 8, "down"))
    for _ in range(4):
        res(512, 512, R // 8)                      # down3 x2, mid x2
    L.append((512, 8, 3, R // 8, "s1")); L.append((8, 8, 1, R // 8, "s1")); L.append((4, 4, 1, R // 8, "s1"))
    L.append((4, 512, 3, R // 8, "s1"))
    for _ in range(5):
        res(512, 512, R // 8)                      # mid x2, up0 x3
    L.append((512, 512, 3, R // 4, "up"))
    for _ in range(3):
        res(512, 512, R // 4)
    L.append((512, 512, 3, R // 2, "up"))
    res(512, 256, R // 2); res(256, 256, R // 2); res(256, 256, R // 2); L.append((256, 256, 3, R, "up"))
    res(256, 128, R); res(128, 128, R); res(128, 128, R)
    L.append((128, 3, 3, R, "s1"))
    return L


def train_flops_per_image(R):
    conv = sum(2.0 * h * h * co * ci * k * k for ci, co, k, h, _ in conv_layers(R))
    T = (R // 8) ** 2
    attn = 2 * (4 * 2.0 * T * 512 * 512 + 2 * 2.0 * T * T * 512)
    return 3.0 * (conv + attn), conv, attn


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ synthetic data
def glyph_batch(B, R, gen, torch):
    """'glyph-like' pixels: +/-1 blocks of 8x8 px, 10 % ink, fp32 NCHW in [-1, 1]."""
    coarse = (torch.rand(B, 3, R // 8, R // 8, generator=gen) < 0.1).float() * 2 - 1
    return coarse.repeat_interleave(8, 2).repeat_interleave(8, 3).contiguous()


TRACK_LAYERS = ["vae.encoder.conv_in", "vae.encoder.down_blocks.0.resnets.0.norm1",
                "vae.decoder.up_blocks.1.resnets.0.norm1"]
CLASSIFY = ["vae.encoder.down_blocks.0.resnets.0.norm1.output", "vae.decoder.up_blocks.1.resnets.0.norm1.output",
            "vae.decoder.conv_norm_out.output"]


def plant_dead_channels(vae, torch):
    with torch.no_grad():
        for n in ("encoder.down_blocks.0.resnets.0.norm1", "decoder.up_blocks.1.resnets.0.norm1"):
            vae.get_submodule(n).weight[::8] = 1e-3


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_step_fn(torch, R, B, track_interval):
    """The reference path on host cores: oracle AutoencoderKL + the reference's tracker formulas
    (oracle/components.py), loss train.py:289-291, clip, AdamW — fp32, all host threads."""
    from oracle.torch_vae import build_oracle, oracle_forward, oracle_losses
    from oracle import components as oc
    torch.set_num_threads(os.cpu_count() or 1)
    vae = build_oracle(42)
    plant_dead_channels(vae, torch)
    opt = torch.optim.AdamW(vae.parameters(), lr=5e-5, weight_decay=1e-2, eps=1e-8)
    buf = {n: [] for n in TRACK_LAYERS}
    hooks = [vae.get_submodule(n[len("vae."):]).register_forward_hook(
        lambda m, i, o, n=n: buf[n].append(oc.mean_abs_per_channel(o))) for n in TRACK_LAYERS]
    gen = torch.Generator().manual_seed(1234)
    state = {"step": 0}

    def step(res=R):
        x = glyph_batch(B, res, gen, torch)
        out = oracle_forward(vae, x, True)
        total, rec, kl = oracle_losses(out, x, 1e-6)
        total.backward()
        torch.nn.utils.clip_grad_norm_(vae.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        state["step"] += 1
        if state["step"] % track_interval == 0:
            for n in TRACK_LAYERS:
                vals = oc.aggregate_per_channel(buf[n])["value"]
                idx = oc.classify_indices(vals, 0.2)
                if n + ".output" in CLASSIFY and n.endswith("norm1"):
                    oc.nudge_gamma(vae.get_submodule(n[len("vae."):]).weight.data, idx.tolist(), 1.2, 1.5)
                buf[n].clear()
        return float(total)
    return step, hooks


def timed_cpu_sample(torch, R, budget_s, steps, warmup, track_interval):
    """Bounded CPU sample: B=1 at the largest resolution r <= R whose (steps+warmup) fit the budget;
    pixel-rate scaled to RxR images (conv work is linear in pixels; attention's quadratic 2.4 % share
    at 512^2 is ignored, which flatters the CPU)."""
    step, hooks = cpu_reference_step_fn(torch, R, 1, track_interval)
    t0 = time.perf_counter()
    step(64)
    t64 = time.perf_counter() - t0
    r = R
    while r > 64 and t64 * (r / 64) ** 2 * (steps + warmup) > budget_s:
        r //= 2
    for _ in range(warmup):
        step(r)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(r)
    dt = (time.perf_counter() - t0) / steps
    for h in hooks:
        h.remove()
    ips_r = 1.0 / dt
    return ips_r * (r / R) ** 2, dt, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    val, dt, r = timed_cpu_sample(torch, args.res, 150.0, args.steps, args.warmup, args.track_interval)
    cores = os.cpu_count() or 1
    sample = (f"{args.steps} timed + {args.warmup} warm-up full training steps (fwd+loss+bwd+clip+AdamW+tracker) of the "
              f"oracle at B=1, {r}x{r}, fp32, {cores} host threads; images/s scaled by ({r}/{args.res})^2 to {args.res}^2")
    line = {"impl": "reference", "metric": metric_name(args.res), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"experiment_fonts_nudge: synthetic {args.res}^2 glyph-like images, tracking + nudge, "
                                   f"random-init SDXL-VAE", "global_batch": args.batch * args.gpus, "resolution": args.res},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ kernel roofline pass
def conv_roofline(torch, vcd, R, B, peaks):
    """CUDA-event timing of the tcgen05 implicit-GEMM kernels on every distinct GEMM-path conv of the model (fprop +
    dgrad + wgrad through the C ABI, as the training step issues them), inputs rotated through > 126 MB so nothing is
    L2-resident.  achieved = EXECUTED FLOPs / measured time summed with the per-step multiplicity: 2*M*N*K per pass
    for ordinary convs; the three Upsample2D convs run as four 2x2 phase convolutions on the
    low-resolution tensor, i.e. 16/36 of the reference formulation's multiply-adds, and only those are counted."""
    ops = vcd.ops
    from collections import Counter
    shapes = Counter(l for l in conv_layers(R) if l[0] % 128 == 0 and l[1] % 128 == 0)
    tot_f = tot_t = ref_f = 0.0
    per_shape = []
    for (ci, co, k, h, kind), cnt in sorted(shapes.items()):
        hin = h // 2 if kind == "up" else (2 * h if kind == "down" else h)
        nbuf = min(8, max(2, int(200e6 // (B * hin * hin * ci * 2)) + 1))
        xs = [torch.randn(B, hin, hin, ci, device="cuda").to(torch.bfloat16).requires_grad_() for _ in range(nbuf)]
        w = (torch.randn(co, ci, k, k, device="cuda") * 0.02).to(torch.bfloat16).requires_grad_()
        bias = torch.zeros(co, device="cuda", dtype=torch.bfloat16).requires_grad_()
        pad = 1 if k == 3 else 0
        g = torch.randn(B, h, h, co, device="cuda").to(torch.bfloat16)
        if kind == "up":
            packs = ops.UpconvPackedWeights()
            run = lambda i: ops.upconv2d(xs[i % nbuf], w, bias, packs).backward(g)
        elif kind == "down":
            packs = ops.PackedWeights()
            run = lambda i: ops.conv2d(xs[i % nbuf], w, bias, packs, stride=2, pad_t=0, pad_l=0, out_hw=(h, h)).backward(g)
        else:
            packs = ops.PackedWeights()
            run = lambda i: ops.conv2d(xs[i % nbuf], w, bias, packs, stride=1, pad_t=pad, pad_l=pad).backward(g)
        for i in range(3):
            run(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        ref = 3 * 2.0 * B * h * h * co * ci * k * k
        fl = ref * 16.0 / 36.0 if kind == "up" else ref
        per_shape.append({"shape": f"{ci}->{co} k{k} {kind} out@{h}", "count": cnt, "ms_fwd_bwd": ms, "tflops": fl / ms / 1e9})
        tot_f += fl * cnt
        ref_f += ref * cnt
        tot_t += ms * cnt
        del xs, w, g
    achieved = tot_f / tot_t / 1e9
    peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r01_ncu_conv_summary.json")) as f:
            traffic = json.load(f).get("dominant_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic,
            "kernel": "umma_pair_kernel / umma_pair_wgrad_kernel / umma_gemm_kernel (fprop+dgrad+wgrad incl. weight packs, "
                      "bias-grad and finalize kernels; executed FLOPs)",
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback",
            "reference_formulation_tflops": ref_f / tot_t / 1e9,
            "per_shape": per_shape, "conv_ms_per_step": tot_t}


def gn_roofline(torch, vcd, R, B, peaks):
    """HBM roofline of the GroupNorm kernels on the largest tensor class of the model (128 channels at full
    resolution): algorithmic bytes (forward 4 B/element, backward reduce 4, backward apply 6 or 8 with the
    skip gradient) / CUDA-event time per kernel.  The forward statistics come from the producing conv's epilogue."""
    from vcd_b200.ops import _p, _st, call, dtype_code
    C, h, n = 128, R, 3
    xs = [torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16) for _ in range(n)]
    gs = [torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16) for _ in range(n)]
    out = torch.empty_like(xs[0])
    gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16)
    beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    sums = torch.empty(B * 32 * 2, dtype=torch.float64, device="cuda")
    dsdb = torch.empty(B * C * 2, dtype=torch.float32, device="cuda")
    colsum = torch.empty(C, dtype=torch.float32, device="cuda")
    slot = vcd.ops.TrackSlot(C, "cuda", 0.0)
    pdt, hw, ne = dtype_code(gamma), h * h, B * h * h * C
    call("vcd_gn_stats", _p(xs[0]), _p(sums), None, 0.0, B, hw, C, 32, _st())
    kernels = {
        "gn_apply_fwd(+SiLU, +per-channel statistics)": (4, lambda i: call(
            "vcd_gn_apply_fwd", _p(xs[i % n]), _p(sums), _p(gamma), _p(beta), pdt, _p(out), _p(slot.raw), 0.0, 1e-6, 1, B, hw, C,
            32, _st())),
        "gn_bwd_reduce": (4, lambda i: call(
            "vcd_gn_bwd_reduce", _p(xs[i % n]), _p(gs[i % n]), _p(sums), _p(gamma), _p(beta), pdt, _p(dsdb), 1e-6, 1, B, hw, C, 32,
            _st())),
        "gn_bwd_apply(+skip gradient, +bias-gradient column sums)": (8, lambda i: call(
            "vcd_gn_bwd_apply", _p(xs[i % n]), _p(gs[i % n]), _p(sums), _p(gamma), _p(beta), pdt, _p(dsdb), _p(out),
            _p(gs[(i + 1) % n]), _p(colsum), None, None, 1e-6, 1, B, hw, C, 32, _st())),
    }
    peak = peaks.get("hbm_gbs", 6650.0)
    res = []
    for name, (bpe, fn) in kernels.items():
        for i in range(2):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(6):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        ach = bpe * ne / ms / 1e6
        res.append({"kernel": name, "bytes_per_element": bpe, "ms": ms, "achieved": ach, "frac": ach / peak})
    worst = min(res, key=lambda r: r["frac"])
    return {"bound": "hbm", "unit": "GB/s", "peak": peak, "tensor": f"[{B},{h},{h},{C}] bf16", "kernels": res,
            "achieved": worst["achieved"], "frac": worst["frac"], "kernel": worst["kernel"]}


# ------------------------------------------------------------------------------------------ main (B200 arm)
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import vcd_b200
    vcd_b200.add_src_to_path()
    from models.sdxl_vae_wrapper import SDXLVAEWrapper
    from tracking.monitor import ActivityMonitor
    from classification.classifier import RegionClassifier
    from intervention.nudger import InterventionHandler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R, B = args.res, args.batch
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass

    wrapper = SDXLVAEWrapper("random-init:42", torch_dtype=torch.bfloat16).to(dev)
    plant_dead_channels(wrapper.vae, torch)
    model = wrapper
    if world > 1 and not args.graph:
        model = torch.nn.parallel.DistributedDataParallel(wrapper, device_ids=[local_rank], gradient_as_bucket_view=True)
    # exactly the constructor call of train.py:184-187 (no fused= flag: torch picks its foreach implementation on CUDA)
    opt = torch.optim.AdamW(wrapper.parameters(), lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8,
                            **({"fused": True} if args.fused_adamw else {}))
    tcfg = {"enabled": True, "track_interval": args.track_interval,
            "target_layers": [{"name": n, "capture_point": "output", "metrics": ["mean_abs_activation_per_channel"]}
                              for n in TRACK_LAYERS]}
    monitor = ActivityMonitor(model, tcfg)
    classifier = RegionClassifier(wrapper.vae, {"enabled": True, "method": "threshold_groupnorm_activity", "threshold": 0.2,
                                                "target_metric_key": "mean_abs_activation_per_channel",
                                                "layers_to_classify": CLASSIFY})
    handler = InterventionHandler(wrapper.vae, {"enabled": True, "strategy": "gentle_nudge_groupnorm_scale",
                                                "nudge_factor": 1.2, "max_scale_value": 1.5,
                                                "intervention_interval": 10}) if rank == 0 else None
    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host = [glyph_batch(B, R, gen, torch).pin_memory() for _ in range(n_host)]
    resident = [h.to(dev) for h in host]
    state = {"gs": 0, "nudged": 0, "inactive": 0}
    kl_weight = 1e-6
    launches_per_step = [0]

    graphed = None
    if args.graph:   # forward + loss + backward captured once in a CUDA graph (vcd_b200.GraphedVAEStep)
        graphed = vcd_b200.GraphedVAEStep(wrapper, kl_weight, resident[0])

    def train_step(x):
        if graphed is not None:
            total, rec, kl = graphed.step(x)
            torch.nn.utils.clip_grad_norm_(wrapper.parameters(), 1.0)
            opt.step()
        else:
            out = model(x, sample_posterior=True)
            total, rec, kl = vcd_b200.vae_loss(out, x, kl_weight)
            total.backward()
            torch.nn.utils.clip_grad_norm_(wrapper.parameters(), 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
        state["gs"] += 1
        gs = state["gs"]
        if gs % args.track_interval == 0:                 # train.py:308-319 cadence
            monitor.step(gs)
            if rank == 0:
                res = classifier.classify(monitor.get_data_for_step(gs), gs)
                if gs % 10 == 0 and res:
                    handler.intervene(res, gs)
                    state["nudged"] += handler.num_nudges_applied
                    state["inactive"] += sum(len(v["inactive_channel_indices"]) for v in res.values())
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(K, e2e):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if os.environ.get("VCD_BENCH_NOGC"):
            import gc
            gc.collect()
            gc.disable()
        barrier()
        l0 = vcd_b200._lib.launches
        e0.record()
        last = None
        trace = [] if os.environ.get("VCD_BENCH_TRACE") else None
        cpu_ms = []
        ends = []
        for i in range(K):
            # the host never runs more than one step ahead of the device (train.py reads three loss scalars with
            # .item() every step, train.py:295-297, so the real loop cannot either); without this bound an occasional
            # full launch queue at the monitor/nudge step produced 100-250 ms outliers
            ahead = int(os.environ.get("VCD_BENCH_AHEAD", "1"))
            if i >= ahead + 1:
                ends[i - ahead - 1].synchronize()
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                trace.append(ev)
            t_cpu0 = time.perf_counter()
            if e2e:
                x = host[i % n_host].to(dev, non_blocking=True)   # H2D from pinned memory inside the timed region
                last = float(train_step(x).detach().float().cpu())  # D2H read of the step's loss
            else:
                # device-resident inputs; the loss scalar is read every step exactly as train.py:295-297 does with
                # .item() (a free-running host made this phase noisy: 89-105 ms per step run to run, against a stable
                # 89-91 ms with the per-step read the real loop has anyway)
                last = float(train_step(resident[i % n_host]).detach().float().cpu())
            if trace is not None:
                cpu_ms.append((time.perf_counter() - t_cpu0) * 1e3)
            ends.append(torch.cuda.Event())
            ends[-1].record()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if trace is not None and rank == 0:
            per = [trace[i].elapsed_time(trace[i + 1]) for i in range(K - 1)] + [trace[-1].elapsed_time(e1)]
            print(f"[trace] {'e2e' if e2e else 'resident'} phase: {ms / K:.2f} ms/step over {K} steps: "
                  + " ".join(f"{t:.0f}" for t in per) + " | host ms/step: " + " ".join(f"{t:.0f}" for t in cpu_ms),
                  file=sys.stderr)
        launches_per_step[0] = (vcd_b200._lib.launches - l0) / K + (graphed.launches_per_replay if graphed is not None else 0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, int(launches_per_step[0] * K), float(last)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation over all GPUs of the node, 1-2 s) holds
    # driver locks and made the first timed phase 10-25 % slower in ~15 % of the runs when it was started right before it;
    # only the samples taken during the timed region are reported
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n_warm = args.warmup if args.quick else max(3, args.warmup)
    for i in range(n_warm):
        train_step(resident[i % n_host])
    # one untimed tracking event (monitor.step -> classify -> intervene): their first call pays one-time host costs
    # (lazy imports, the classifier's GroupNorm map, first D2H of the packed statistics) of 100-170 ms, which a real run
    # amortises over thousands of steps but which would inflate a 20-step timed region by ~5 %
    warm_gs = 1000 * args.track_interval
    monitor.step(warm_gs)
    if rank == 0:
        res = classifier.classify(monitor.get_data_for_step(warm_gs), warm_gs)
        if res:
            handler.intervene(res, warm_gs)
    # A fresh box pages libraries in, loads CUDA modules lazily and ramps clocks during its first process: keep
    # warming (untimed, counted in "warmup") until two consecutive steps agree within 10 %, at most 10 extra steps
    prev = None
    for _ in range(0 if args.quick else 10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        train_step(resident[n_warm % n_host])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_warm += 1
        steady = prev is not None and abs(dt - prev) < 0.10 * prev
        if world > 1:
            flag = torch.tensor([1.0 if steady else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            steady = bool(flag.item() > 0.5)
        prev = dt
        if steady:
            break
    clocks.rows.clear()   # keep only the samples of the timed region
    if os.environ.get("VCD_BENCH_ORDER") == "e2e_first" and not args.quick:
        ms_e2e, _, loss_e2e = timed(args.steps, e2e=True)
        ms, launches, loss = timed(args.steps, e2e=False)
    else:
        ms, launches, loss = timed(args.steps, e2e=False)
        ms_e2e, _, loss_e2e = (ms, 0, loss) if args.quick else timed(args.steps, e2e=True)
    clk = clocks.stop() if rank == 0 else None
    if os.environ.get("VCD_BENCH_TRACE") and rank == 0:
        print("[trace] sm clocks (200 ms samples): " + " ".join(r[0] for r in clocks.rows) + " | power W: "
              + " ".join(r[2].split(".")[0] for r in clocks.rows if len(r) > 2), file=sys.stderr)
    value = args.steps * B * world / (ms / 1e3)
    value_e2e = args.steps * B * world / (ms_e2e / 1e3)

    if args.kernel_table and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                float(train_step(resident[i % n_host]).detach().float().cpu())   # as in the timed region
            torch.cuda.synchronize()
        # idle time on the device between consecutive kernels, attributed to the kernel that follows the gap
        evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
                     key=lambda e: e.time_range.start)
        gaps = {}
        for a, b in zip(evs, evs[1:]):
            g = b.time_range.start - a.time_range.end
            if g > 0:
                k = (a.name[:48], b.name[:48])
                c = gaps.setdefault(k, [0, 0.0])
                c[0] += 1
                c[1] += g
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        with open(args.kernel_table, "w") as f:
            f.write(f"2 steps, {tot / 2e3:.2f} ms of device time per step (torch.profiler / CUPTI, warm, in-pipeline)\n")
            f.write(f"{'ms/step':>9} {'share':>6} {'n/step':>7}  kernel\n")
            for e in rows[:45]:
                f.write(f"{e.device_time_total / 2e3:9.3f} {100 * e.device_time_total / tot:5.1f}% {e.count / 2:7.1f}  {e.key[:110]}\n")
            gtot = sum(v[1] for v in gaps.values())
            f.write(f"\nidle gaps between consecutive device activities: {gtot / 2e3:.2f} ms per step\n")
            f.write(f"{'ms/step':>9} {'n/step':>7} {'us each':>8}  previous -> next\n")
            for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
                f.write(f"{v[1] / 2e3:9.3f} {v[0] / 2:7.1f} {v[1] / v[0]:8.1f}  {k[0]} -> {k[1]}\n")

    roof = roof_hbm = cpu = None
    if rank == 0 and not args.no_roofline and not args.quick:
        del host, resident
        torch.cuda.empty_cache()
        roof = conv_roofline(torch, vcd_b200, R, B, peaks)
        roof_hbm = gn_roofline(torch, vcd_b200, R, B, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        v, dt, r = timed_cpu_sample(torch, R, 25.0, 1, 0, args.track_interval)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 full training step of the oracle (plain-torch fp32 restatement + reference tracker formulas) at "
                         f"B=1, {r}x{r}, {os.cpu_count()} host threads, {dt:.1f} s; images/s scaled by ({r}/{R})^2 to {R}^2"}
    if rank == 0:
        fl_img, _, _ = train_flops_per_image(R)
        peak_t = peaks.get("bf16_tflops_sustained", 1414.5)
        line = {
            "metric": metric_name(R), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"experiment_fonts_nudge: synthetic {R}^2 glyph-like images, tracking (3 layers) every forward, "
                                   f"classify+nudge every {args.track_interval} steps, random-init SDXL-VAE seed 42, bf16 weights",
                       "resolution": R, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "optimizer": "clip_grad_norm 1.0 + torch.optim.AdamW " + ("(fused=True)" if args.fused_adamw else "constructed as in train.py:184-187") + ", stepped as in :301-304",
                       "execution": "eager per-op launches (DDP bucketed all-reduce for N>1)" if not args.graph else "forward+loss+backward replayed from one CUDA graph (GraphedVAEStep); clip/AdamW/tracker eager",
                       "l2": "4 distinct input batches rotated; every activation tensor exceeds the 126 MB L2 at this size",
                       "value_phase": "inputs resident in HBM; loss scalar read back every step (train.py:295-297)",
                       "train_tflop_per_image": fl_img / 1e12,
                       "step_mfu_of_sustained_peak": (value / world) * fl_img / 1e12 / peak_t,
                       "nudges_applied": state["nudged"], "inactive_flagged": state["inactive"], "final_loss": loss},
            "e2e": {"value": value_e2e, "unit": UNIT, "h2d_bytes_per_step": B * 3 * R * R * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
