"""CPU: the C-ABI library loads and exports every symbol include/vcd.h declares (no compute calls without a
GPU), the ctypes table matches the header's arity, the host modules keep the reference's interface, and the
product path fails loudly — never falls back — when there is no CUDA device."""
import inspect
import os
import re

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    src = open(os.path.join(ROOT, "include", "vcd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(vcd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


def test_library_exports_every_declared_symbol(vcd):
    decls = _header_decls()
    assert len(decls) >= 35
    lib = vcd._lib.lib()
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/vcd.h but not exported by libvcd_b200.so"
    assert lib.vcd_version() >= 100
    assert lib.vcd_last_error() is not None


def test_ctypes_table_matches_header(vcd):
    decls = _header_decls()
    sigs = vcd._lib.SIGNATURES
    assert set(sigs) == set(decls), set(sigs) ^ set(decls)
    for name, (_, argtypes) in sigs.items():
        assert len(argtypes) == decls[name], (name, len(argtypes), decls[name])


def test_library_is_sm100a_tcgen05_code():
    import subprocess
    lib = os.path.join(ROOT, "vae-channel-dynamics_b200", "libvcd_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STG.E.128", "LDG.E.128"):
        assert mnemonic in sass, mnemonic


def test_module_tree_is_diffusers_compatible(vcd):
    from oracle.torch_vae import build_oracle
    model = vcd.B200AutoencoderKL()
    oracle = build_oracle(1)
    assert [n for n, _ in model.named_parameters()] == [n for n, _ in oracle.named_parameters()]
    assert [tuple(p.shape) for p in model.parameters()] == [tuple(p.shape) for p in oracle.parameters()]
    model.load_state_dict(oracle.state_dict())
    assert model.config.scaling_factor == 0.13025 and model.config["latent_channels"] == 4
    gns = [m for m in model.modules() if isinstance(m, nn.GroupNorm)]          # classifier.py:56
    assert len(gns) == 52 and all(isinstance(m.weight, nn.Parameter) for m in gns)
    assert sum(isinstance(m, (nn.Conv2d, nn.Linear)) for m in model.modules()) == 64 + 8   # train.py:38


def test_save_and_from_pretrained_roundtrip(vcd, tmp_path):
    torch.manual_seed(3)
    m = vcd.B200AutoencoderKL()
    m.save_pretrained(str(tmp_path / "vae"))
    assert sorted(os.listdir(tmp_path / "vae")) == ["config.json", "diffusion_pytorch_model.safetensors"]
    m2 = vcd.B200AutoencoderKL.from_pretrained(str(tmp_path / "vae"), torch_dtype=torch.bfloat16)
    assert m2.dtype == torch.bfloat16
    for (n, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a.to(torch.bfloat16), b), n
    m3 = vcd.B200AutoencoderKL.from_pretrained("random-init:7")
    m4 = vcd.B200AutoencoderKL.from_pretrained("random-init:7")
    assert torch.equal(m3.encoder.conv_in.weight, m4.encoder.conv_in.weight)
    with pytest.raises(vcd.VcdError):
        vcd.B200AutoencoderKL.from_pretrained("stabilityai/sdxl-vae")       # no network, no silent random weights


def test_host_modules_keep_reference_interface(vcd):
    vcd.add_src_to_path()
    from models.sdxl_vae_wrapper import SDXLVAEWrapper
    from tracking.monitor import ActivityMonitor
    from tracking.deadneuron import DeadNeuronTracker
    from classification.classifier import RegionClassifier
    from intervention.nudger import InterventionHandler
    assert list(inspect.signature(SDXLVAEWrapper.__init__).parameters) == ["self", "pretrained_model_name_or_path", "torch_dtype"]
    assert list(inspect.signature(SDXLVAEWrapper.forward).parameters) == ["self", "pixel_values", "sample_posterior"]
    for meth in ("add_hooks", "remove_hooks", "get_captured_activations", "clear_captured_activations", "encode", "decode"):
        assert hasattr(SDXLVAEWrapper, meth)
    for meth in ("step", "get_data_for_step", "export_all_processed_data_to_records", "remove_hooks", "_get_layer"):
        assert hasattr(ActivityMonitor, meth)
    assert list(inspect.signature(DeadNeuronTracker.__init__).parameters) == [
        "self", "target_layer_classes", "target_layer_names_for_raw_weights", "threshold", "mean_percentage", "dead_type"]
    assert list(inspect.signature(RegionClassifier.classify).parameters) == ["self", "tracked_data_for_step", "global_step"]
    assert list(inspect.signature(InterventionHandler.intervene).parameters) == ["self", "classification_results", "global_step"]
    w = SDXLVAEWrapper("random-init:42")
    assert w.scaling_factor == 0.13025 and isinstance(w.vae, vcd.B200AutoencoderKL)
    # classifier map: 52 GroupNorms -> plain + 'vae.'-prefixed keys (classifier.py:43-81)
    clf = RegionClassifier(w.vae, {"enabled": True, "threshold": 0.2})
    assert len(clf._layer_to_param_map) == 104
    assert clf._lookup_param_info("vae.encoder.down_blocks.0.resnets.0.norm1.output") == \
        ("encoder.down_blocks.0.resnets.0.norm1.weight", 128)
    assert clf._lookup_param_info("vae.encoder.conv_in.output") is None
    assert clf.classify({}, 1) == {} and RegionClassifier(w.vae, {"enabled": False}).classify({"x": {}}, 1) == {}
    # nudger guards (nudger.py:89-103) need no device
    ih = InterventionHandler(w.vae, {"enabled": True, "strategy": "gentle_nudge_groupnorm_scale", "intervention_interval": 10})
    ih.intervene({}, 10)
    ih.intervene({"k": {"param_name_scale": "nope.weight", "inactive_channel_indices": [0]}}, 10)
    assert ih.num_nudges_applied == 0 and ih._get_parameter("decoder.conv_norm_out.weight") is not None
    # monitor registration resolves the shipped config names through a DDP-style '.module' wrapper (monitor.py:41-54)
    class Wrap(nn.Module):
        def __init__(self, m):
            super().__init__()
            self.module = m
    mon = ActivityMonitor(Wrap(w), {"enabled": True, "track_interval": 20, "target_layers": [
        {"name": "vae.encoder.conv_in", "capture_point": "output", "metrics": ["mean_abs_activation_per_channel"]},
        {"name": "vae.decoder.up_blocks.1.resnets.0.norm1", "capture_point": "input", "metrics": ["full_activation_map"]},
        {"name": "vae.does.not.exist", "capture_point": "output"}]})
    assert len(mon.hooks) == 2 and mon.step(7) == {} and mon.get_data_for_step(20) == {}
    mon.remove_hooks()
    assert mon.hooks == [] and not w.vae.encoder.conv_in._forward_hooks
    assert ActivityMonitor(w, {"enabled": False}).step(20) == {}


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(vcd):
    m = vcd.B200AutoencoderKL()
    with pytest.raises(vcd.VcdError, match="CUDA"):
        m.encode(torch.zeros(1, 3, 32, 32))
    vcd.add_src_to_path()
    from classification.classifier import RegionClassifier
    from tracking.deadneuron import DeadNeuronTracker
    clf = RegionClassifier(m, {"enabled": True, "threshold": 0.2})
    with pytest.raises(vcd.VcdError):
        clf.classify({"encoder.conv_norm_out.output": {"mean_abs_activation_per_channel": np.zeros(512, np.float32)}}, 1)
    trk = DeadNeuronTracker((nn.Conv2d,), [], 1e-3, 0.1, "both")
    with pytest.raises(vcd.VcdError):
        trk.get_percentage(torch.zeros(8))


def test_bench_flop_model_matches_survey():
    import bench
    assert len(bench.conv_layers(512)) == 64
    kinds = [l[4] for l in bench.conv_layers(512)]
    assert kinds.count("up") == 3 and kinds.count("down") == 3 and kinds.count("s1") == 58
    total, conv, attn = bench.train_flops_per_image(512)
    assert abs(conv / 1e9 - 3545.3) < 0.1 and abs(attn / 1e9 - 85.9) < 0.1 and abs(total / 1e12 - 10.89) < 0.01
    total, conv, attn = bench.train_flops_per_image(256)
    assert abs(total / 1e12 - 2.685) < 0.001


def test_groupnorm_producers_are_marked(vcd):
    """Every conv / linear whose output is the input of a GroupNorm emits that GroupNorm's sums from its GEMM epilogue
    (vae.py _mark_groupnorm_producers): conv_in, every resnet conv1, block outputs that feed a norm1 / attention /
    conv_norm_out, the sampler convs and the attention output projection — but NOT the last resnet of a block that ends
    in a Down/Upsample2D conv, the 1x1 shortcuts, conv_out or the quant convs."""
    m = vcd.B200AutoencoderKL()
    g = m.config["norm_num_groups"]
    marked = {n for n, mod in m.named_modules() if getattr(mod, "_gn_groups", 0) == g}
    assert "encoder.conv_in" in marked and "decoder.conv_in" in marked
    for n, mod in m.named_modules():
        if n.endswith(".conv1"):
            assert n in marked, n
        if n.endswith(".conv_shortcut") or n.endswith("conv_out") or n in ("quant_conv", "post_quant_conv"):
            assert n not in marked, n
    assert "encoder.down_blocks.0.resnets.0.conv2" in marked          # feeds resnets.1.norm1
    assert "encoder.down_blocks.0.resnets.1.conv2" not in marked      # feeds the Downsample2D conv
    assert "encoder.down_blocks.0.downsamplers.0.conv" in marked      # feeds down_blocks.1.resnets.0.norm1
    assert "encoder.down_blocks.3.resnets.1.conv2" in marked          # no downsampler: feeds mid_block.resnets.0.norm1
    assert "decoder.up_blocks.2.resnets.2.conv2" not in marked        # feeds the Upsample2D conv
    assert "decoder.up_blocks.2.upsamplers.0.conv" in marked
    assert "decoder.up_blocks.3.resnets.2.conv2" in marked            # feeds decoder.conv_norm_out
    assert "encoder.mid_block.attentions.0.to_out.0" in marked and "encoder.mid_block.attentions.0.to_q" not in marked
    # one producer per GroupNorm: 52 GroupNorms, 52 marked producers
    assert len(marked) == sum(1 for mod in m.modules() if isinstance(mod, nn.GroupNorm))


def test_from_pretrained_reads_legacy_sdxl_vae_checkpoints(tmp_path):
    """SURVEY 8f-3 / train.py:412, evaluate.py:99: the published stabilityai/sdxl-vae file predates diffusers' attention
    refactor — its mid-block attention parameters are named query / key / value / proj_attn ([upstream] diffusers converts
    them on load), and older exports store them as 1x1 convolutions [C, C, 1, 1].  A directory in that layout must load
    into the same module tree, and save_pretrained must write the modern diffusers names back."""
    import json

    import torch
    from safetensors.torch import load_file, save_file
    import vcd_b200
    torch.manual_seed(3)
    vae = vcd_b200.B200AutoencoderKL()
    sd = {k: v.detach().clone() for k, v in vae.state_dict().items()}
    legacy = {}
    ren = {".to_q.": ".query.", ".to_k.": ".key.", ".to_v.": ".value.", ".to_out.0.": ".proj_attn."}
    n_renamed = 0
    for k, v in sd.items():
        nk = k
        for new, old in ren.items():
            if new in k:
                nk = k.replace(new, old)
                n_renamed += 1
                if k.endswith(".weight"):
                    v = v[:, :, None, None].contiguous()          # Linear [C, C] stored as a 1x1 conv
        legacy[nk] = v
    assert n_renamed == 16 and not any(".to_q." in k for k in legacy)
    d = tmp_path / "legacy_vae"
    d.mkdir()
    save_file(legacy, str(d / "diffusion_pytorch_model.safetensors"), metadata={"format": "pt"})
    cfg = dict(vcd_b200.vae.SDXL_VAE_CONFIG, _diffusers_version="0.18.0.dev0", _name_or_path="stabilityai/sdxl-vae")
    (d / "config.json").write_text(json.dumps(cfg))
    loaded = vcd_b200.B200AutoencoderKL.from_pretrained(str(d))
    for k, v in loaded.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert loaded.config.scaling_factor == 0.13025 and loaded.config["_class_name"] == "AutoencoderKL"
    half = vcd_b200.B200AutoencoderKL.from_pretrained(str(d), torch_dtype=torch.bfloat16)
    assert half.dtype == torch.bfloat16
    out = tmp_path / "resaved"
    loaded.save_pretrained(str(out))
    again = load_file(str(out / "diffusion_pytorch_model.safetensors"))
    assert set(again) == set(sd) and all(torch.equal(again[k], sd[k]) for k in sd)
    assert json.loads((out / "config.json").read_text())["block_out_channels"] == [128, 256, 512, 512]


def test_zero_pool_hands_out_disjoint_zero_slices_once(vcd, monkeypatch):
    """Host logic of the zero-filled accumulator arenas (ops._ZeroPool): every slice is zero, 256-byte aligned inside its
    chunk, handed out exactly once (no two slices overlap), typed as asked; a request that does not fit the current chunk
    opens a new one while the old chunk stays alive through its slices; oversize requests, a disabled pool and a CUDA-graph
    capture fall back to an uninitialised tensor WITHOUT the VCD_ACC_PREZEROED flag.  (CPU tensors stand in for device
    memory; the stream key and the capture query are the two CUDA calls of take().)"""
    import torch
    ops, lib = vcd.ops, vcd._lib
    capturing = {"v": False}
    monkeypatch.setattr(ops, "_st", lambda: 7)
    monkeypatch.setattr(torch._C, "_cuda_getDevice", lambda: 0)
    monkeypatch.setattr(torch._C, "_cuda_isCurrentStreamCapturing", lambda: capturing["v"])
    monkeypatch.setattr(ops, "zero_pool_enabled", True)
    pool = ops._ZeroPool(4096)
    dev = torch.device("cpu")
    taken = []
    for numel, dt in ((10, torch.float64), (3, torch.float32), (100, torch.float32), (1, torch.float64)):
        t, flag = pool.take(numel, dt, dev)
        assert flag == lib.ACC_PREZEROED and t.dtype == dt and t.numel() == numel and float(t.abs().sum()) == 0.0
        taken.append(t)
    chunk0 = pool.cur[(0, 7)][0]
    base = chunk0.data_ptr()
    spans = sorted((t.data_ptr() - base, t.data_ptr() - base + t.numel() * t.element_size()) for t in taken)
    assert all(a % 256 == 0 for a, _ in spans)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    for t in taken:
        t.fill_(1.0)                                   # the kernels add into their slices ...
    t, flag = pool.take(16, torch.float32, dev)        # ... and a later slice is still zero
    assert flag and float(t.abs().sum()) == 0.0
    # chunk rollover: 4096-byte chunk, 1280 bytes used so far -> a 3000-byte request opens a fresh chunk
    big, flag = pool.take(750, torch.float32, dev)
    assert flag and pool.cur[(0, 7)][0] is not chunk0 and float(big.abs().sum()) == 0.0
    assert float(taken[0].sum()) == 10.0               # the old chunk lives on through its slices
    # per-stream chunks
    monkeypatch.setattr(ops, "_st", lambda: 9)
    other, flag = pool.take(4, torch.float32, dev)
    assert flag and (0, 9) in pool.cur and pool.cur[(0, 9)][0] is not pool.cur[(0, 7)][0]
    # fallbacks: too large for a chunk, capture in progress, pool disabled
    t, flag = pool.take(2000, torch.float32, dev)
    assert flag == 0 and t.numel() == 2000
    capturing["v"] = True
    t, flag = pool.take(4, torch.float32, dev)
    assert flag == 0
    capturing["v"] = False
    monkeypatch.setattr(ops, "zero_pool_enabled", False)
    t, flag = pool.take(4, torch.float32, dev)
    assert flag == 0
    pool.reset()
    assert not pool.cur


def test_wgrad_overlap_mode_selection(vcd, monkeypatch):
    """VCD_WGRAD_OVERLAP / VCD_WGRAD_STREAM / the bench's profiling switch select how a conv's weight gradient overlaps its
    data gradient (ops._wgrad_overlap_mode): side stream by default, programmatic dependent launch or none on request, and
    none at all while bench.py brackets every call with CUDA events."""
    ops = vcd.ops
    monkeypatch.delenv("VCD_WGRAD_OVERLAP", raising=False)
    monkeypatch.delenv("VCD_WGRAD_STREAM", raising=False)
    monkeypatch.setattr(ops, "wgrad_side_stream_enabled", True)
    assert ops._wgrad_overlap_mode() == "stream"
    monkeypatch.setenv("VCD_WGRAD_OVERLAP", "pdl")
    assert ops._wgrad_overlap_mode() == "pdl"
    monkeypatch.setenv("VCD_WGRAD_OVERLAP", "off")
    assert ops._wgrad_overlap_mode() == "off"
    monkeypatch.setenv("VCD_WGRAD_OVERLAP", "pdl")
    monkeypatch.setenv("VCD_WGRAD_STREAM", "0")
    assert ops._wgrad_overlap_mode() == "off"
    monkeypatch.delenv("VCD_WGRAD_STREAM")
    monkeypatch.setattr(ops, "wgrad_side_stream_enabled", False)
    assert ops._wgrad_overlap_mode() == "off"
    assert vcd._lib.WGRAD_OVERLAP_PREV == 0x100 and vcd._lib.ACC_PREZEROED == 0x200      # include/vcd.h
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "vcd.h")).read()
    assert "#define VCD_WGRAD_OVERLAP_PREV 0x100" in hdr and "#define VCD_ACC_PREZEROED 0x200" in hdr


def test_pack_plan_tiles_cover_every_weight_exactly_once(vcd):
    """Host side of vcd_multi_pack_weights (ops.PackPlan): the work list covers every (Cout, Cin) position of every layer
    exactly once with the library's tile size, including the ragged small-channel layers (3 -> 128, 512 -> 8, 4 -> 4) and
    Linear weights (taps = 1); descriptors carry the shapes the kernel indexes with.  Built on CPU tensors (no launch)."""
    import ctypes as C
    ops, lib = vcd.ops, vcd._lib.lib()
    tco, tci = lib.vcd_pack_tile_co(), lib.vcd_pack_tile_ci()
    assert tco >= 8 and tci >= 8 and tco % 8 == 0 and tci % 8 == 0       # 16-byte stores of eight channels
    shapes = [(128, 3, 3, 3), (8, 512, 3, 3), (4, 4, 1, 1), (512, 512), (256, 128, 1, 1), (128, 128, 3, 3), (512, 4, 3, 3)]
    layers = []
    for sh in shapes:
        w = torch.zeros(sh)
        b = torch.zeros(sh[0])
        layers.append((w, b, ops.PackedWeights(), 0))
    layers.append((torch.zeros(256, 256, 3, 3), torch.zeros(256), ops.UpconvPackedWeights(), 1))     # Upsample2D conv
    plan = ops.PackPlan(layers)
    assert plan.n_tiles == len(plan.tile_layer) == len(plan.tile_co) == len(plan.tile_ci)
    seen = {}
    for l, co, ci in zip(plan.tile_layer.tolist(), plan.tile_co.tolist(), plan.tile_ci.tolist()):
        cout, cin = layers[l][0].shape[0], layers[l][0].shape[1]
        assert co % tco == 0 and ci % tci == 0 and co < cout and ci < cin
        assert (l, co, ci) not in seen
        seen[(l, co, ci)] = True
    for l, (w, _, packs, mode) in enumerate(layers):
        cout, cin = w.shape[0], w.shape[1]
        assert sum(1 for k in seen if k[0] == l) == -(-cout // tco) * -(-cin // tci)
        taps = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
        assert packs.wf.numel() == (16 if mode == 1 else taps) * cout * cin == packs.wd.numel()
        assert packs.bias is not None and packs.bias.numel() == cout and packs.bias.dtype == torch.float32
    # descriptor bytes: 5 pointers + 6 int32 per layer, shapes in the order the kernel reads them
    raw = bytes(plan.descs.tolist())
    rec = 5 * 8 + 6 * 4
    assert len(raw) == rec * len(layers)
    for l, (w, _, _, mode) in enumerate(layers):
        dtype, cout, cin, taps, m, _pad = (C.c_int32 * 6).from_buffer_copy(raw[l * rec + 40:(l + 1) * rec])
        assert (cout, cin, m) == (w.shape[0], w.shape[1], mode)
        assert taps == (w.shape[2] * w.shape[3] if w.dim() == 4 else 1) and dtype == vcd._lib.F32
