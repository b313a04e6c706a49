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
    ap.add_argument("--optimizer", default="vcd-fused", choices=["vcd-fused", "torch", "torch-fused"],
                    help="vcd-fused (default): vcd_b200.FusedClipAdamW adopting the torch.optim.AdamW that train.py:184-187 constructs "
                         "(what accelerate.prepare does under VCD_FUSED_OPT=1, SURVEY 8f-2: clip + AdamW in two launches); torch: that "
                         "torch optimizer itself (foreach) + torch clip_grad_norm_; torch-fused: torch AdamW(fused=True)")
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


CLASSIFIER_CONFIG = {"enabled": True, "method": "threshold_groupnorm_activity", "threshold": 0.2,
                     "target_metric_key": "mean_abs_activation_per_channel", "layers_to_classify": CLASSIFY}
INTERVENTION_CONFIG = {"enabled": True, "strategy": "gentle_nudge_groupnorm_scale", "nudge_factor": 1.2, "max_scale_value": 1.5,
                       "intervention_interval": 10}
DNT_RAW = ["vae.encoder.conv_in.weight", "vae.decoder.conv_out.weight"]      # experiment_fonts_nudge.yaml dead_neuron_tracking


def DNT_CLASSES(torch):
    return (torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Conv3d, torch.nn.Linear, torch.nn.GroupNorm)     # train.py:38


def tracking_config(track_interval):
    return {"enabled": True, "track_interval": track_interval,
            "target_layers": [{"name": n, "capture_point": "output", "metrics": ["mean_abs_activation_per_channel"]}
                              for n in TRACK_LAYERS]}


def lr_lambda(step, warmup=100, max_steps=50 * 6250):
    """train.py:197-200 with experiment_fonts_nudge.yaml: lr_warmup_steps 100 (base_config), 50 epochs x 50 000 / 8 steps."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    return max(0.0, 1.0 - min(1.0, float(step - warmup) / float(max(1, max_steps - warmup))))


def plant_dead_channels(vae, torch):
    with torch.no_grad():
        for n in ("encoder.down_blocks.0.resnets.0.norm1", "decoder.up_blocks.1.resnets.0.norm1"):
            vae.get_submodule(n).weight[::8] = 1e-3


# ------------------------------------------------------------------------------------------ CPU reference arm
REF_SRC = os.path.join(ROOT, "baseline", "_ref", "src")     # unmodified copy of the reference (see __graft_entry__.stage_reference)


def _reference_modules():
    """The reference's OWN ActivityMonitor / RegionClassifier / InterventionHandler / DeadNeuronTracker (they need only
    torch + numpy), imported from baseline/_ref/src; None when that copy is not on the box."""
    if not os.path.isfile(os.path.join(REF_SRC, "tracking", "monitor.py")):
        return None
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_SRC, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    return {"monitor": load("_ref_monitor", "tracking/monitor.py"), "deadneuron": load("_ref_deadneuron", "tracking/deadneuron.py"),
            "classifier": load("_ref_classifier", "classification/classifier.py"), "nudger": load("_ref_nudger", "intervention/nudger.py")}


def cpu_reference_step_fn(torch, R, B, track_interval):
    """The reference path on host cores: oracle AutoencoderKL (plain-torch restatement of the diffusers arithmetic —
    diffusers is absent) behind the reference's own tracker / classifier / nudger / dead-weight modules when
    baseline/_ref is present (else their restatement oracle/components.py), loss train.py:289-291, clip, AdamW, LambdaLR
    — fp32, all host threads."""
    import logging
    from oracle.torch_vae import build_oracle, oracle_forward, oracle_losses
    from oracle import components as oc
    logging.disable(logging.WARNING)
    torch.set_num_threads(os.cpu_count() or 1)
    vae = build_oracle(42)
    plant_dead_channels(vae, torch)
    opt = torch.optim.AdamW(vae.parameters(), lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda)
    mods = _reference_modules()
    gen = torch.Generator().manual_seed(1234)
    state = {"step": 0}
    hooks = []
    if mods is not None:
        holder = torch.nn.Module()
        holder.vae = vae
        monitor = mods["monitor"].ActivityMonitor(holder, tracking_config(track_interval))
        classifier = mods["classifier"].RegionClassifier(vae, CLASSIFIER_CONFIG)
        handler = mods["nudger"].InterventionHandler(vae, INTERVENTION_CONFIG)
        dnt = mods["deadneuron"].DeadNeuronTracker(DNT_CLASSES(torch), DNT_RAW, threshold=1e-8, mean_percentage=0.01,
                                                   dead_type="threshold")
        kind = "oracle VAE + the reference's own src/tracking, src/classification, src/intervention modules"
    else:
        buf = {n: [] for n in TRACK_LAYERS}
        hooks = [vae.get_submodule(n[len("vae."):]).register_forward_hook(
            lambda m, i, o, n=n: buf[n].append(oc.mean_abs_per_channel(o))) for n in TRACK_LAYERS]
        kind = "oracle VAE + oracle/components.py restatement of the tracker formulas (baseline/_ref absent)"

    def step(res=R):
        x = glyph_batch(B, res, gen, torch)
        out = oracle_forward(vae, x, True)
        total, rec, kl = oracle_losses(out, x, 1e-6)
        items = [float(t.detach()) for t in (total, rec, kl)]          # train.py:292-297
        total.backward()
        torch.nn.utils.clip_grad_norm_(vae.parameters(), 1.0)
        opt.step()
        sched.step()
        opt.zero_grad(set_to_none=True)
        state["step"] += 1
        gs = state["step"]
        if gs % track_interval == 0:
            if mods is not None:
                monitor.step(gs)
                res_c = classifier.classify(monitor.get_data_for_step(gs), gs)
                if gs % INTERVENTION_CONFIG["intervention_interval"] == 0 and res_c:
                    handler.intervene(res_c, gs)
                dnt.track_dead_neurons(vae, gs)
            else:
                for n in TRACK_LAYERS:
                    vals = oc.aggregate_per_channel(buf[n])["value"]
                    idx = oc.classify_indices(vals, 0.2)
                    if n + ".output" in CLASSIFY and n.endswith("norm1"):
                        oc.nudge_gamma(vae.get_submodule(n[len("vae."):]).weight.data, idx.tolist(), 1.2, 1.5)
                    buf[n].clear()
        return items[0]
    return step, hooks, kind


def timed_cpu_sample(torch, R, budget_s, steps, warmup, track_interval):
    """Bounded CPU sample of the SAME workload: B = 1 (one image per step instead of the GPU arm's per-GPU batch) at the
    real resolution R when (steps + warmup) such steps fit the budget; otherwise the largest r = R/2^k that fits, with
    images/s scaled by (r/R)^2 and the result labelled 'extrapolated' (conv work is linear in pixels; the attention's
    quadratic share — 2.4 % at 512^2 — is then under-counted, which flatters the CPU)."""
    step, hooks, kind = cpu_reference_step_fn(torch, R, 1, track_interval)
    step(64)                                   # cold start (allocator, threads): not part of the size estimate
    t0 = time.perf_counter()
    step(128 if R >= 128 else R)
    t_probe = (time.perf_counter() - t0) * (R / min(R, 128)) ** 2      # estimate of one RxR step
    r = R
    while r > 64 and t_probe * (r / R) ** 2 * (steps + warmup) > budget_s:
        r //= 2
    for _ in range(warmup):
        step(r)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(r)
    dt = (time.perf_counter() - t0) / steps
    for h in hooks:
        h.remove()
    return (1.0 / dt) * (r / R) ** 2, dt, r, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    val, dt, r, kind = timed_cpu_sample(torch, args.res, 210.0, args.steps, args.warmup, args.track_interval)
    cores = os.cpu_count() or 1
    extr = "" if r == args.res else f" — EXTRAPOLATED: timed at {r}x{r}, images/s scaled by ({r}/{args.res})^2"
    sample = (f"{args.steps} timed + {args.warmup} warm-up full training steps (fwd + loss + 3 loss reads + bwd + clip + AdamW + "
              f"LambdaLR + tracker/classifier/nudger/dead-weight cadence) at B=1, {r}x{r}, fp32, {cores} host threads; {kind}{extr}")
    line = {"impl": "reference", "metric": metric_name(args.res), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"experiment_fonts_nudge: synthetic {args.res}^2 glyph-like images, tracking + nudge, "
                                   f"random-init SDXL-VAE" + (" (extrapolated from a smaller resolution)" if extr else ""),
                       "global_batch": args.batch * args.gpus, "resolution": args.res, "timed_resolution": r, "timed_batch": 1},
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
        def step(i, run=run, xs=xs, w=w, bias=bias):
            # as in the training step, no gradient is ACCUMULATED: activations are non-leaf there and zero_grad() sets the
            # parameter gradients to None, so the kernels' outputs become the gradients without a read-modify-write pass
            xs[i % nbuf].grad = None
            w.grad = None
            bias.grad = None
            run(i)

        for i in range(3):
            step(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            step(i)
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
    # these timings run back to back for ~2 s on a GPU that is already under the power cap: the sustained cuBLAS rate is the
    # matching denominator (B200_PROFILING.md); the fraction of the burst peak is reported next to it
    peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    burst = peaks.get("bf16_tflops") or peak
    traffic = traffic_info = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on a named shape (ncu --set full)
        with open(os.path.join(ROOT, "profiles", "r02_ncu_conv_summary.json")) as f:
            traffic_info = json.load(f).get("roofline_traffic")
            traffic = traffic_info.get("dram_bytes_per_launch")
    except Exception:
        pass
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "frac_of_burst_peak": achieved / burst, "burst_peak": burst,
            "traffic": traffic, "traffic_detail": traffic_info,
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
            "vcd_gn_apply_fwd", _p(xs[i % n]), _p(sums), _p(gamma), _p(beta), pdt, _p(out), None, _p(slot.raw), 0.0, 1e-6, 1, B, hw,
            C, 32, _st())),
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
# ------------------------------------------------------------------------------------------ in-step kernel timing
def _conv_flops(name, a):
    """Executed FLOPs of one GEMM-path C-ABI call from its integer arguments (include/vcd.h argument order)."""
    if name in ("vcd_conv2d_fprop", "vcd_conv2d_dgrad", "vcd_conv2d_wgrad"):
        if name == "vcd_conv2d_wgrad":
            a = a[1:]                      # leading dtype code
        N, H, W, Cin, Cout, KH, KW, stride = a[0:8]
        Ho, Wo = a[10], a[11]
        return 2.0 * N * Ho * Wo * Cout * Cin * KH * KW
    if name == "vcd_conv2d_dgrad_gn":
        N, H, W, Cin, Cout, KH, KW = a[0:7]
        return 2.0 * N * H * W * Cout * Cin * KH * KW
    if name in ("vcd_upconv2d_fprop", "vcd_upconv2d_dgrad", "vcd_upconv2d_wgrad"):
        N, H, W, Cin, Cout = a[-5:] if name != "vcd_upconv2d_fprop" else a[0:5]
        return 2.0 * N * (2 * H) * (2 * W) * Cout * Cin * 4       # four 2x2 phase convolutions: 16 taps per low-res pixel
    if name == "vcd_gemm_nt":
        b, M, Nn, K = a[0:4]
        return 2.0 * b * M * Nn * K
    if name == "vcd_gemm_tn":
        b, M, Nn, K = a[1:5]
        return 2.0 * b * M * Nn * K
    return None


def _gn_bytes(name, a, has_res):
    """Algorithmic HBM bytes of one GroupNorm C-ABI call: fwd 4 B/element, bwd reduce 4, bwd apply 6 (+2 with
    the skip-connection gradient)."""
    if name == "vcd_gn_apply_fwd":
        N, HW, C = a[-4], a[-3], a[-2]
        return 4.0 * N * HW * C
    if name == "vcd_gn_bwd_reduce":
        N, HW, C = a[-4], a[-3], a[-2]
        return 4.0 * N * HW * C
    if name == "vcd_gn_bwd_apply":
        N, HW, C = a[-4], a[-3], a[-2]
        return (8.0 if has_res else 6.0) * N * HW * C
    if name == "vcd_gn_stats":
        N, HW, C = a[-4], a[-3], a[-2]
        return 2.0 * N * HW * C
    return None


def instep_rooflines(torch, vcd, train_step, batches, steps, peaks):
    """Runs `steps` more training steps with EVERY C-ABI call bracketed by CUDA events on its launching stream
    (vcd_b200._lib.profile), i.e. each kernel timed in the pipeline, with the cache / clock / power state of the real
    step.  roofline (tensor): executed FLOPs of all implicit-GEMM entry points / their summed durations; roofline_hbm:
    algorithmic bytes of the GroupNorm entry points / their durations, per kernel family."""
    lib = vcd._lib
    vcd.ops.wgrad_side_stream_enabled = False   # two concurrently running kernels would each be charged the other's time
    float(train_step(batches[0]))          # one untimed step with the hook installed (event pool warm-up)
    torch.cuda.synchronize()
    lib.profile = []
    for i in range(steps):
        float(train_step(batches[i % len(batches)]))
    torch.cuda.synchronize()
    rec, lib.profile = lib.profile, None
    vcd.ops.wgrad_side_stream_enabled = True
    by = {}
    for name, args, e0, e1 in rec:
        # integer arguments without device pointers (>= 2^31) and without the trailing stream handle
        ints = [v for v in args[:-1] if isinstance(v, int) and not isinstance(v, bool) and abs(v) < (1 << 31)]
        ms = e0.elapsed_time(e1)
        fl = _conv_flops(name, ints)
        if fl is not None:
            d = by.setdefault(("tensor", name), [0, 0.0, 0.0])
            d[0] += 1; d[1] += ms; d[2] += fl
            continue
        has_res = name == "vcd_gn_bwd_apply" and args[8] is not None
        by_ = _gn_bytes(name, ints, has_res)
        if by_ is not None:
            key = name + ("+skip_grad" if has_res else "")
            if name == "vcd_gn_apply_fwd" and (args[6] is not None or args[7] is not None):
                key += "+channel_stats"
            d = by.setdefault(("hbm", key), [0, 0.0, 0.0])
            d[0] += 1; d[1] += ms; d[2] += by_
        else:
            d = by.setdefault(("other", name), [0, 0.0, 0.0])
            d[0] += 1; d[1] += ms
    tens = {k[1]: v for k, v in by.items() if k[0] == "tensor"}
    hbm = {k[1]: v for k, v in by.items() if k[0] == "hbm"}
    oth = {k[1]: v for k, v in by.items() if k[0] == "other"}
    tf = sum(v[2] for v in tens.values())
    tt = sum(v[1] for v in tens.values())
    peak_s = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    peak_b = peaks.get("bf16_tflops") or peak_s
    achieved = tf / tt / 1e9
    traffic = tinfo = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_conv_summary.json")) as f:
            tinfo = json.load(f).get("roofline_traffic")
            traffic = tinfo.get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak_s, "unit": "TFLOP/s", "frac": achieved / peak_s,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long power-capped step)"
                           if "bf16_tflops_sustained" in peaks else "fallback 1590 (B200_PROFILING.md)",
            "frac_of_burst_peak": achieved / peak_b, "burst_peak": peak_b,
            "traffic": traffic, "traffic_detail": tinfo,
            "kernel": "umma_pair_kernel / umma_pair_wgrad_kernel / umma_gemm_kernel behind the conv / upconv / gemm entry points "
                      "(entry-point durations include their memsets, finalize and colsum kernels); EXECUTED FLOPs (Upsample2D "
                      "convs run as four 2x2 phase convolutions = 16/36 of the reference formulation)",
            "how": f"CUDA events around every C-ABI call on its launching stream, {steps} training steps in the pipeline "
                   f"(weight-gradient kernels kept on the main stream for these steps so that no two timed kernels overlap)",
            "ms_per_step": tt / steps, "tflop_per_step": tf / steps / 1e12,
            "by_entry_point": {n: {"calls_per_step": v[0] / steps, "ms_per_step": v[1] / steps, "tflops": v[2] / v[1] / 1e9}
                               for n, v in sorted(tens.items())}}
    peak_h = peaks.get("hbm_gbs", 6650.0)
    kern = [{"kernel": n, "calls_per_step": v[0] / steps, "ms_per_step": v[1] / steps, "achieved": v[2] / v[1] / 1e6,
             "frac": v[2] / v[1] / 1e6 / peak_h} for n, v in sorted(hbm.items())]
    gb = sum(v[2] for v in hbm.values())
    gt = sum(v[1] for v in hbm.values())
    roof_hbm = {"bound": "hbm", "unit": "GB/s", "peak": peak_h, "achieved": gb / gt / 1e6, "frac": gb / gt / 1e6 / peak_h,
                "kernel": "GroupNorm(+SiLU, +statistics) forward / backward kernels (gn.cu), all launches of the step",
                "how": roof["how"], "ms_per_step": gt / steps, "algorithmic_gb_per_step": gb / steps / 1e9, "kernels": kern}
    other = {n: {"calls_per_step": v[0] / steps, "ms_per_step": v[1] / steps} for n, v in sorted(oth.items(), key=lambda kv: -kv[1][1])[:12]}
    return roof, roof_hbm, other


# ------------------------------------------------------------------------------------------ main (B200 arm)
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    # every CUDA module is loaded at context creation, not lazily inside the first steps (a fresh box otherwise spends its
    # first steps loading kernels, which the driver's 5 warm-up steps would not cover)
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

    import torch
    import torch.distributed as dist
    import vcd_b200
    vcd_b200.add_src_to_path()
    from models.sdxl_vae_wrapper import SDXLVAEWrapper
    from tracking.monitor import ActivityMonitor
    from tracking.deadneuron import DeadNeuronTracker
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
        bucket_mb = float(os.environ.get("VCD_BENCH_BUCKET_MB", "25"))       # torch's default bucket_cap_mb
        model = torch.nn.parallel.DistributedDataParallel(wrapper, device_ids=[local_rank], gradient_as_bucket_view=True,
                                                          bucket_cap_mb=bucket_mb)
    # exactly the constructor call of train.py:184-187 (no fused= flag: torch picks its foreach implementation on CUDA) ...
    torch_opt = torch.optim.AdamW(wrapper.parameters(), lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8,
                                  **({"fused": True} if args.optimizer == "torch-fused" else {}))
    sched = torch.optim.lr_scheduler.LambdaLR(torch_opt, lr_lambda)            # train.py:197-202
    # ... adopted by the fused clip+AdamW (shares param_groups with torch_opt, so the LambdaLR above keeps driving the lr):
    # what an accelerate-compatible prepare() does under VCD_FUSED_OPT=1 (tests/shims/accelerate)
    opt = vcd_b200.FusedClipAdamW.from_torch(torch_opt) if args.optimizer == "vcd-fused" else torch_opt
    monitor = ActivityMonitor(model, tracking_config(args.track_interval))
    classifier = RegionClassifier(wrapper.vae, CLASSIFIER_CONFIG)
    handler = InterventionHandler(wrapper.vae, INTERVENTION_CONFIG) if rank == 0 else None
    dnt = DeadNeuronTracker(target_layer_classes=DNT_CLASSES(torch), target_layer_names_for_raw_weights=DNT_RAW,
                            threshold=1e-8, mean_percentage=0.01, dead_type="threshold")          # train.py:104-106,219-223
    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host = [glyph_batch(B, R, gen, torch).pin_memory() for _ in range(n_host)]
    resident = [h.to(dev) for h in host]
    state = {"gs": 0, "nudged": 0, "inactive": 0}
    kl_weight = 1e-6

    graphed = None
    if args.graph:   # forward + loss + backward captured once in a CUDA graph (vcd_b200.GraphedVAEStep)
        graphed = vcd_b200.GraphedVAEStep(wrapper, kl_weight, resident[0])

    def gather_mean(t):                                    # accelerator.gather(x).mean(), train.py:292-294
        if world == 1:
            return t
        out = [torch.empty_like(t.reshape(1)) for _ in range(world)]
        dist.all_gather(out, t.reshape(1))
        return torch.cat(out).mean()

    def clip():
        if hasattr(opt, "clip_grad_norm_"):
            opt.clip_grad_norm_(None, 1.0)
        else:
            torch.nn.utils.clip_grad_norm_(wrapper.parameters(), 1.0)

    def train_step(x):
        """train.py:286-330,355-356 for one batch"""
        if graphed is not None:
            total, rec, kl = graphed.step(x)
            loss_value = float(total)
        else:
            out = model(x, sample_posterior=True)
            total, rec, kl = vcd_b200.vae_loss(out, x, kl_weight)                      # train.py:289-291
            g = [gather_mean(t.detach()) for t in (total, rec, kl)]                    # :292-294
            loss_value, _, _ = (v.item() for v in g)                                   # :295-297 — host sync BEFORE backward
            total.backward()                                                           # :299
        clip()                                                                         # :301
        opt.step()
        sched.step()
        if graphed is None:
            opt.zero_grad(set_to_none=True)                                            # :304
        state["gs"] += 1
        gs = state["gs"]
        if gs % args.track_interval == 0:                                              # :308-330
            monitor.step(gs)
            if rank == 0:
                res = classifier.classify(monitor.get_data_for_step(gs), gs)
                if gs % INTERVENTION_CONFIG["intervention_interval"] == 0 and res:
                    handler.intervene(res, gs)
                    state["nudged"] += handler.num_nudges_applied
                    state["inactive"] += sum(len(v["inactive_channel_indices"]) for v in res.values())
            dnt.track_dead_neurons(wrapper.vae, gs)                                    # :355-356 (dead_neuron_tracking.track_interval 20)
        return loss_value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(K, e2e):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if os.environ.get("VCD_BENCH_NOGC"):       # experiment knob: is a Python GC pause visible in the step time?
            import gc
            gc.collect()
            gc.disable()
        barrier()
        l0 = vcd_b200._lib.launches
        e0.record()
        last = None
        trace = [] if os.environ.get("VCD_BENCH_TRACE") else None
        # e2e: every step's batch is copied from pinned host memory inside the timed region, as train.py does it (batch.to(device)
        # at the start of the step).  VCD_BENCH_PREFETCH=1: through vcd_b200.data.DevicePrefetcher (batch i+1 on a copy stream
        # while step i runs) — measured on B200: 86.29 vs 85.83-86.64 ms per step, no gain (the 25 MB copy is 0.5 ms and mostly
        # hidden behind the host's launch lead already), so the plain copy stays the default
        if e2e and os.environ.get("VCD_BENCH_PREFETCH"):
            feed = iter(vcd_b200.data.DevicePrefetcher((host[i % n_host] for i in range(K)), dev))
        elif e2e:
            feed = (host[i % n_host].to(dev, non_blocking=True) for i in range(K))
        else:
            feed = (resident[i % n_host] for i in range(K))
        for i in range(K):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                trace.append(ev)
            last = train_step(next(feed))                          # the three loss scalars come back to the host every step
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if trace is not None and rank == 0:
            per = [trace[i].elapsed_time(trace[i + 1]) for i in range(K - 1)] + [trace[-1].elapsed_time(e1)]
            print(f"[trace] {'e2e' if e2e else 'resident'} phase: {ms / K:.2f} ms/step over {K} steps: "
                  + " ".join(f"{t:.0f}" for t in per), file=sys.stderr)
        n_launch = vcd_b200._lib.launches - l0 + (graphed.launches_per_replay * K if graphed is not None else 0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, int(n_launch), float(last)

    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation over all GPUs of the node, 1-2 s) holds
    # driver locks and made the first timed phase 10-25 % slower in ~15 % of the runs when it was started right before it;
    # only the samples taken during the timed region are reported
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n_warm = args.warmup if args.quick else max(3, args.warmup)       # exactly the W asked for (>= 3: timing rules)
    for i in range(n_warm):
        train_step(resident[i % n_host])
    # one untimed tracking event (monitor.step -> classify -> intervene -> dead-weight scan): their first call pays one-time
    # host costs (lazy imports, the classifier's GroupNorm map, first D2H of the packed statistics) of 100-170 ms, which a
    # real run amortises over thousands of steps but which would inflate a 20-step timed region by ~5 %.  Not a train step.
    warm_gs = 1000 * args.track_interval
    monitor.step(warm_gs)
    if rank == 0:
        res = classifier.classify(monitor.get_data_for_step(warm_gs), warm_gs)
        if res:
            handler.intervene(res, warm_gs)
    dnt.track_dead_neurons(wrapper.vae, warm_gs)
    if os.environ.get("VCD_BENCH_SETTLE"):     # opt-in: keep warming until two consecutive steps agree within 10 % (reported in "warmup")
        prev = None
        for _ in range(10):
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
    if os.environ.get("VCD_BENCH_AB") and world == 1:
        # profiling aid: VCD_BENCH_AB="VAR=a,b" times --steps steps under VAR=a and VAR=b alternately (3 rounds) inside this
        # one process — same box, same thermal state — for knobs the host side reads at run time (e.g. VCD_WGRAD_OVERLAP)
        var, vals = os.environ["VCD_BENCH_AB"].split("=")
        res_ab = {v: [] for v in vals.split(",")}
        for _ in range(3):
            for v in res_ab:
                os.environ[var] = v
                if var == "VCD_ZERO_POOL":      # read once at import by ops.py
                    vcd_b200.ops.zero_pool_enabled = v == "1"
                train_step(resident[0])
                res_ab[v].append(timed(args.steps, e2e=False)[0] / args.steps)
        print("[ab] " + var + ": " + "; ".join(f"{v}: " + " ".join(f"{t:.2f}" for t in ts) + f" (min {min(ts):.2f}) ms/step"
                                                 for v, ts in res_ab.items()), file=sys.stderr, flush=True)
        os.environ.pop(var, None)
        vcd_b200.ops.zero_pool_enabled = os.environ.get("VCD_ZERO_POOL", "1") == "1"
    clocks.rows.clear()   # keep only the samples of the timed region
    ms, launches, loss = timed(args.steps, e2e=False)
    ms_e2e, _, loss_e2e = (ms, 0, loss) if args.quick else timed(args.steps, e2e=True)
    clk = clocks.stop() if rank == 0 else None
    if os.environ.get("VCD_BENCH_TRACE") and rank == 0:
        print("[trace] sm clocks (200 ms samples): " + " ".join(r[0] for r in clocks.rows) + " | power W: "
              + " ".join(r[2].split(".")[0] for r in clocks.rows if len(r) > 2), file=sys.stderr)
    value = args.steps * B * world / (ms / 1e3)
    value_e2e = args.steps * B * world / (ms_e2e / 1e3)

    roof = roof_hbm = other = cpu = None
    if not args.no_roofline and not args.quick and graphed is None:
        # every rank runs the instrumented steps (DDP collectives must match); rank 0 reports
        roof, roof_hbm, other = instep_rooflines(torch, vcd_b200, train_step, resident, 3, peaks)

    if args.kernel_table and rank == 0:
        write_kernel_table(torch, args.kernel_table, train_step, resident)
    elif args.kernel_table and world > 1:
        for i in range(2):
            train_step(resident[i % n_host])     # keep the collectives of the profiled steps matched on the other ranks

    if rank == 0 and roof is not None and world == 1 and not os.environ.get("VCD_BENCH_NO_PER_SHAPE"):
        del host, resident
        torch.cuda.empty_cache()
        iso = conv_roofline(torch, vcd_b200, R, B, peaks)         # isolated per-shape rates (explains the in-step aggregate)
        roof["isolated_per_shape"] = {"achieved": iso["achieved"], "per_shape": iso["per_shape"],
                                      "conv_ms_per_step": iso["conv_ms_per_step"],
                                      "reference_formulation_tflops": iso["reference_formulation_tflops"]}
        roof_hbm["isolated_largest_tensor"] = gn_roofline(torch, vcd_b200, R, B, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        v, dt, r, kind = timed_cpu_sample(torch, R, 30.0, 1, 0, args.track_interval)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 full training step ({kind}) at B=1, {r}x{r}, fp32, {os.cpu_count()} host threads, {dt:.1f} s"
                         + ("" if r == R else f"; EXTRAPOLATED: images/s scaled by ({r}/{R})^2 to {R}^2")}
    if rank == 0:
        fl_img, _, _ = train_flops_per_image(R)
        peak_t = peaks.get("bf16_tflops_sustained", 1414.5)
        opt_desc = {"vcd-fused": "vcd_b200.FusedClipAdamW adopting the torch.optim.AdamW of train.py:184-187 (clip + AdamW in two launches; "
                                 "accelerate.prepare opt-in VCD_FUSED_OPT=1)",
                    "torch": "torch clip_grad_norm_ + torch.optim.AdamW constructed as in train.py:184-187 (foreach)",
                    "torch-fused": "torch clip_grad_norm_ + torch.optim.AdamW(fused=True)"}[args.optimizer]
        line = {
            "metric": metric_name(R), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"experiment_fonts_nudge: synthetic {R}^2 glyph-like images, tracking (3 layers) every forward, "
                                   f"classify+nudge+dead-weight scan every {args.track_interval} steps, random-init SDXL-VAE seed 42, "
                                   f"bf16 weights",
                       "resolution": R, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "loop": "train.py:286-330,355-356 — forward, loss, 3 gathers + 3 .item() before backward, backward, clip, "
                               "optimizer step, LambdaLR step, zero_grad, monitor/classifier/nudger/dead-weight cadence",
                       "optimizer": opt_desc,
                       "execution": "eager per-op launches (DDP bucketed all-reduce for N>1)" if not args.graph else "forward+loss+backward replayed from one CUDA graph (GraphedVAEStep); clip/AdamW/tracker eager",
                       "l2": "4 distinct input batches rotated; every activation tensor exceeds the 126 MB L2 at this size",
                       "value_phase": "inputs resident in HBM; the three loss scalars read back every step (train.py:295-297)",
                       "train_tflop_per_image": fl_img / 1e12,
                       "step_mfu_of_sustained_peak": (value / world) * fl_img / 1e12 / peak_t,
                       "nudges_applied": state["nudged"], "inactive_flagged": state["inactive"], "final_loss": loss},
            "e2e": {"value": value_e2e, "unit": UNIT, "h2d_bytes_per_step": B * 3 * R * R * 4, "d2h_bytes_per_step": 12,
                    "ms_per_step": ms_e2e / args.steps,
                    "h2d": ("every step's batch copied from pinned host memory inside the timed region, batch i+1 on a copy "
                            "stream while step i runs (vcd_b200.data.DevicePrefetcher)" if os.environ.get("VCD_BENCH_PREFETCH")
                            else "batch.to(device) on the compute stream at the start of every step")},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": roof, "roofline_hbm": roof_hbm, "other_entry_points_ms_per_step": other, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def write_kernel_table(torch, path, train_step, resident):
    from torch.profiler import ProfilerActivity, profile
    hostgap = bool(os.environ.get("VCD_BENCH_HOSTGAP"))      # profiling aid: what does the HOST do during the largest device gap?
    acts = [ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if hostgap else [])
    with profile(activities=acts) as prof:
        for i in range(2):
            train_step(resident[i % len(resident)])
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
    host_lines = []
    if hostgap:
        big = sorted(((b.time_range.start - a.time_range.end, a, b) for a, b in zip(evs, evs[1:])), key=lambda t: -t[0])[:3]
        cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
        for g, a, b in big:
            t0, t1 = a.time_range.end, b.time_range.start
            inside = sorted((e for e in cpu if e.time_range.end > t0 and e.time_range.start < t1),
                            key=lambda e: -(min(e.time_range.end, t1) - max(e.time_range.start, t0)))[:14]
            host_lines.append(f"gap {g:.0f} us between {a.name[:40]} and {b.name[:40]}: host events overlapping it")
            for e in inside:
                ov = min(e.time_range.end, t1) - max(e.time_range.start, t0)
                host_lines.append(f"    {ov:8.0f} us of {e.time_range.end - e.time_range.start:8.0f} us  {e.name[:90]}")
    rows = sorted((e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA),
                  key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    with open(path, "w") as f:
        for l in host_lines:
            f.write(l + "\n")
        f.write(f"2 steps, {tot / 2e3:.2f} ms of device time per step (torch.profiler / CUPTI, warm, in-pipeline)\n")
        f.write(f"{'ms/step':>9} {'share':>6} {'n/step':>7}  kernel\n")
        for e in rows[:45]:
            f.write(f"{e.device_time_total / 2e3:9.3f} {100 * e.device_time_total / tot:5.1f}% {e.count / 2:7.1f}  {e.key[:110]}\n")
        gtot = sum(v[1] for v in gaps.values())
        f.write(f"\nidle gaps between consecutive device activities: {gtot / 2e3:.2f} ms per step\n")
        f.write(f"{'ms/step':>9} {'n/step':>7} {'us each':>8}  previous -> next\n")
        for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write(f"{v[1] / 2e3:9.3f} {v[0] / 2:7.1f} {v[1] / v[0]:8.1f}  {k[0]} -> {k[1]}\n")


if __name__ == "__main__":
    main()
