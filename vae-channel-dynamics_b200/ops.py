"""Autograd bindings of the libvcd_b200 kernels.

Internal activation layout: bf16, NHWC contiguous tensors of shape [N, H, W, C].  Parameters stay
ordinary ``nn.Parameter``s (fp32 or bf16, diffusers OIHW layout) so the optimizer, DDP, the nudger
and the dead-weight tracker all see the storage they see in the reference; GEMM operand packs are
rebuilt on every forward (fused optimizers do not bump ``param._version``, so no host-side key can prove a pack current).
"""
from __future__ import annotations

import functools
import math
from typing import Optional

import torch

from . import _lib
from ._lib import BF16, F32, IMPL_AUTO, IMPL_SIMT, IMPL_UMMA, call

_conv_impl = IMPL_AUTO


def set_conv_impl(impl: int) -> None:
    """IMPL_AUTO (default), IMPL_SIMT (CUDA-core cross-check) or IMPL_UMMA."""
    global _conv_impl
    _conv_impl = impl


def get_conv_impl() -> int:
    return _conv_impl


def _st() -> int:
    # raw cudaStream_t of the current stream of the current device (torch.cuda.current_stream() builds a Stream object and
    # costs ~16 us per call — 6 ms of host time per training step)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _on_tensor_device(fwd):
    """Decorator of every autograd Function.forward below: kernels go to the CURRENT stream of the CURRENT device (_st), so a
    tensor that lives on another GPU (a model on cuda:1 in a process whose current device is cuda:0) makes its own device
    current for the call.  Backward nodes already run under autograd's device guard."""
    @functools.wraps(fwd)
    def guarded(ctx, *args, **kw):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fwd(ctx, *args, **kw)
                break
        return fwd(ctx, *args, **kw)
    return guarded


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise _lib.VcdError(f"unsupported parameter dtype {t.dtype} (fp32 and bf16 only; fp16 is not a meaningful "
                        "target, SURVEY appendix A.10)")


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.VcdError(f"{what} must live on a CUDA device: the hot path has no CPU implementation")


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.bfloat16:
        raise _lib.VcdError(f"internal activations are bf16, got {x.dtype}")
    return x if x.is_contiguous() else x.contiguous()


# ------------------------------------------------------------------------------------------
# zero-filled accumulator arenas
# ------------------------------------------------------------------------------------------
class _ZeroPool:
    """Buffers that kernels ADD into (GroupNorm sums from the GEMM epilogues, backward sums, bias-gradient column sums,
    split-K weight-gradient workspaces) are carved out of zero-filled chunks and handed to the entry points with
    VCD_ACC_PREZEROED, instead of being zeroed one by one inside every call: a training step at 512^2 issued 246 memsets —
    2 us of device time plus one more launch boundary each (measured: ~1 ms per step) — and now issues ~25 memsets and a
    handful of chunk fills.

    A chunk is an ordinary caching-allocator tensor, filled once (`torch.zeros`) on the stream that uses it; every slice
    is handed out exactly once and keeps the chunk alive, so lifetimes need no bookkeeping (forward sums live until their
    backward, a retained graph keeps its chunk).  Chunks are per (device, stream): the weight-gradient side stream owns
    its own.  Inside a CUDA-graph capture the pool is bypassed (a chunk must not straddle the capture boundary)."""

    def __init__(self, chunk_bytes: int):
        self.chunk_bytes = chunk_bytes
        self.cur = {}     # (device, raw stream) -> [chunk, offset]

    def take(self, numel: int, dtype: torch.dtype, device):
        """-> (tensor, VCD_ACC_PREZEROED) out of the arena, or (uninitialised tensor, 0) when it does not apply"""
        nbytes = numel * _ITEMSIZE[dtype]
        need = (nbytes + 255) & ~255
        if not zero_pool_enabled or need > self.chunk_bytes or torch._C._cuda_isCurrentStreamCapturing():
            return torch.empty(numel, dtype=dtype, device=device), 0
        key = (torch._C._cuda_getDevice(), _st())
        e = self.cur.get(key)
        if e is None or e[1] + need > self.chunk_bytes:
            e = self.cur[key] = [torch.zeros(self.chunk_bytes, dtype=torch.uint8, device=device), 0]
        off = e[1]
        e[1] = off + need
        return e[0][off:off + nbytes].view(dtype), _lib.ACC_PREZEROED

    def reset(self) -> None:
        self.cur.clear()


_ITEMSIZE = {torch.float32: 4, torch.float64: 8}
zero_pool_enabled = __import__("os").environ.get("VCD_ZERO_POOL", "1") == "1"
_acc_pool = _ZeroPool(2 << 20)       # sums [N][G][2] fp64, dsdb [N][C][2] fp32, column sums [C] fp32: ~1.5 MB per step at B=8
_ws_pool = _ZeroPool(96 << 20)       # fp32 [taps][Cout][Cin] split-K accumulators (<= 9.4 MB per layer, 336 MB per step)


def zero_pool_reset() -> None:
    _acc_pool.reset()
    _ws_pool.reset()


# ------------------------------------------------------------------------------------------
# statistics slots (device side of ActivityMonitor)
# ------------------------------------------------------------------------------------------
class TrackSlot:
    """Per (layer, capture point) device accumulators.

    raw  fp32 [5][C] : sums of ONE forward (sum, sum^2, sum|x|, max|x|, #near-zero), zeroed by finalize
    run  fp32 [5][C] : sum over forwards of per-forward (mean|x|, mean, var, max, near-zero fraction)
    scal fp64 [3]    : sum of per-forward mean_activation, std_activation, forward count
    """

    def __init__(self, channels: int, device, near_zero: float = 0.0):
        self.C = channels
        self.near_zero = float(near_zero)
        self.raw = torch.zeros(5 * channels, dtype=torch.float32, device=device)
        self.run = torch.zeros(5 * channels, dtype=torch.float32, device=device)
        self.scal = torch.zeros(3, dtype=torch.float64, device=device)
        self.on_finalize = None  # monitor callback: records the order in which capture points fire

    def finalize(self, n_per_channel: int) -> None:
        if self.on_finalize is not None:
            self.on_finalize()
        call("vcd_stats_finalize", _p(self.raw), _p(self.run), _p(self.scal), int(n_per_channel), self.C, _st())

    def reset(self) -> None:
        self.run.zero_()
        self.scal.zero_()
        self.raw.zero_()


def chan_stats(x: torch.Tensor, slot: TrackSlot) -> None:
    """Stand-alone statistics of a logically-[N, C, *] tensor (any strides handled via layout flag)."""
    _require_cuda(x, "tracked tensor")
    N, Cc = x.shape[0], x.shape[1]
    hw = x.numel() // max(1, N * Cc)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    dt = dtype_code(x)
    if x.dim() >= 3 and x.permute(0, *range(2, x.dim()), 1).is_contiguous():
        call("vcd_chan_stats", _p(x), dt, _p(slot.raw), slot.near_zero, N, hw, Cc, 1, _st())
    else:
        x = x.contiguous()
        call("vcd_chan_stats", _p(x), dt, _p(slot.raw), slot.near_zero, N, hw, Cc, 0, _st())
    slot.finalize(N * hw)


# ------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------
class PackedWeights:
    """bf16 GEMM operand packs of one conv / linear layer, refreshed when the parameters change."""

    def __init__(self):
        self.key = None
        self.frozen = False
        self.valid = False   # set by pack_model_weights(): the packs were just rebuilt by the model-wide launch
        self.wf = self.wd = self.bias = None

    def ensure(self, weight: torch.Tensor, bias: Optional[torch.Tensor], taps_out: int):
        cout, cin = weight.shape[0], weight.shape[1]
        n = taps_out * cout * cin
        if self.wf is None or self.wf.numel() != n or self.wf.device != weight.device:
            self.wf = torch.empty(n, dtype=torch.bfloat16, device=weight.device)
            self.wd = torch.empty(n, dtype=torch.bfloat16, device=weight.device)
            self.bias = None if bias is None else torch.empty(cout, dtype=torch.float32, device=weight.device)
            self.key = None

    def get(self, weight: torch.Tensor, bias: Optional[torch.Tensor], training: bool = False):
        """training: this forward will be differentiated w.r.t. the weights (ctx.needs_input_grad inside the autograd
        Function — grad mode itself is off there)."""
        if self.valid:          # rebuilt at the start of this encode()/decode() by ONE launch for all layers
            self.valid = False
            return self.wf, self.wd, self.bias
        key = (weight.data_ptr(), weight._version, weight.dtype, None if bias is None else (bias.data_ptr(), bias._version))
        # EVERY forward repacks (75 small kernels, ~0.4 ms per 512^2 step): fused optimizers (torch.optim.AdamW(fused=True))
        # and `.data` writes update a parameter without touching `_version`, so no host-side key can prove that the packs
        # are current — a version-keyed cache left the GEMMs on the weights of the first step.  `frozen` (set by
        # freeze_weight_packs for inference loops on fixed weights) is the only way to skip the repack.
        if key != self.key or not self.frozen or training or torch.cuda.is_current_stream_capturing():
            _require_cuda(weight, "conv weight")
            w = weight.detach()
            if not w.is_contiguous():
                w = w.contiguous()
            cout, cin = w.shape[0], w.shape[1]
            kh, kw = (w.shape[2], w.shape[3]) if w.dim() == 4 else (1, 1)
            n = kh * kw * cout * cin
            if self.wf is None or self.wf.numel() != n or self.wf.device != w.device:
                self.wf = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                self.wd = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                self.bias = None if bias is None else torch.empty(cout, dtype=torch.float32, device=w.device)
            call("vcd_pack_conv_weight", _p(w), _p(None if bias is None else bias.detach()), dtype_code(w), cout, cin,
                 kh, kw, _p(self.wf), _p(self.wd), _p(self.bias), _st())
            self.key = key
        return self.wf, self.wd, self.bias

    def current(self):
        """the packs of the forward pass (backward never repacks)"""
        return self.wf, self.wd, self.bias


def freeze_weight_packs(model: torch.nn.Module, frozen: bool = True) -> None:
    """Inference on fixed weights: keep the GEMM operand packs of every layer between forwards (they are still rebuilt
    when a parameter's `_version` changes).  Training code never needs this."""
    for m in model.modules():
        for attr in ("_packs", "_up_packs"):
            pk = getattr(m, attr, None)
            if pk is not None:
                pk.frozen = frozen


class PackPlan:
    """Device-side work list of vcd_multi_pack_weights for a fixed set of layers: (weight, bias, packs, mode)."""

    def __init__(self, layers):
        import ctypes as C
        self.layers = layers
        self.key = self.make_key(layers)
        dev = layers[0][0].device
        tco, tci = _lib.lib().vcd_pack_tile_co(), _lib.lib().vcd_pack_tile_ci()

        class Desc(C.Structure):
            _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p), ("wf", C.c_void_p), ("wd", C.c_void_p), ("bias_f32", C.c_void_p),
                        ("dtype", C.c_int32), ("cout", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32), ("mode", C.c_int32),
                        ("pad_", C.c_int32)]
        descs = (Desc * len(layers))()
        tl, tc, ti = [], [], []
        for i, (w, b, packs, mode) in enumerate(layers):
            cout, cin = w.shape[0], w.shape[1]
            taps = (w.shape[2] * w.shape[3]) if w.dim() == 4 else 1
            packs.ensure(w, b, 16 if mode == 1 else taps)
            if not w.is_contiguous():
                raise _lib.VcdError("pack plan: weights must be contiguous")
            descs[i] = Desc(w.data_ptr(), None if b is None else b.data_ptr(), packs.wf.data_ptr(), packs.wd.data_ptr(),
                            None if packs.bias is None else packs.bias.data_ptr(), dtype_code(w), cout, cin, taps, mode, 0)
            for co in range(0, cout, tco):
                for ci in range(0, cin, tci):
                    tl.append(i); tc.append(co); ti.append(ci)
        raw = bytes(descs)
        self.descs = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        self.tile_layer, self.tile_co, self.tile_ci = i32(tl), i32(tc), i32(ti)
        self.n_tiles = len(tl)

    @staticmethod
    def make_key(layers):
        return tuple((w.data_ptr(), w.dtype, None if b is None else b.data_ptr(), p.wf.data_ptr() if p.wf is not None else 0)
                     for w, b, p, _ in layers)

    def run(self):
        call("vcd_multi_pack_weights", _p(self.descs), _p(self.tile_layer), _p(self.tile_co), _p(self.tile_ci), self.n_tiles, _st())
        for _, _, packs, _ in self.layers:
            packs.valid = True

    def expire(self):
        """end of the encode() / decode() call the packs were built for: a later direct call of a layer packs itself"""
        for _, _, packs, _ in self.layers:
            packs.valid = False


def _workspace(fn: str, shape, impl: int, device):
    if impl == IMPL_SIMT:
        return None
    nbytes = getattr(_lib.lib(), fn)(*shape)
    return torch.empty(nbytes, dtype=torch.uint8, device=device) if nbytes > 0 else None


# column sums of a gradient tensor, produced for free by the kernel that wrote it (vcd_gn_bwd_apply) and
# consumed as the bias gradient of the conv that receives that tensor as dy
_COLSUMS = {}


def push_colsum(t: torch.Tensor, colsum: torch.Tensor) -> None:
    # the entry keeps `t` alive, so its address cannot be recycled for another tensor while the entry exists; the version
    # counter detects an in-place change of `t` after the hand-off (autograd accumulating a second gradient into it)
    _COLSUMS[t.data_ptr()] = (t, colsum, t._version)


def pop_colsum(t: torch.Tensor):
    e = _COLSUMS.pop(t.data_ptr(), None)
    if e is None or e[0].numel() != t.numel() or e[0].shape[-1] != t.shape[-1] or e[0]._version != e[2]:
        return None
    return e[1]


def clear_colsums() -> None:
    _COLSUMS.clear()
    _GNSUMS.clear()
    _GN_FWD.clear()
    _GN_BWD.clear()


# GroupNorm -> conv backward fusion (vcd_conv2d_dgrad_gn).  _GN_FWD: output tensor of a GroupNorm(+SiLU) whose ONLY
# consumer is a 3x3 conv -> what that conv's dgrad epilogue needs (x, sums, gamma, beta, eps, act, groups).
# _GN_BWD: the tensor such a dgrad returned (g = dL/d(pre-activation)) -> the per-channel sums it already reduced.
_GN_FWD = {}
_GN_BWD = {}


# GroupNorm sums (sum, sum of squares per (image, group)) of a tensor, produced by the epilogue of the GEMM that wrote
# it (vcd_conv2d_fprop gn_sums) and consumed by the GroupNorm that normalises it: no statistics pass over the tensor
_GNSUMS = {}


def push_gn_sums(t: torch.Tensor, sums: torch.Tensor, groups: int) -> None:
    _GNSUMS[t.data_ptr()] = (t, sums, groups, t._version)


def pop_gn_sums(t: torch.Tensor, groups: int):
    e = _GNSUMS.pop(t.data_ptr(), None)
    if e is None or e[0].shape != t.shape or e[2] != groups or e[0]._version != e[3]:   # modified in place since
        return None
    return e[1]


# Weight-gradient kernels on a side stream (default; VCD_WGRAD_STREAM=0 disables): dgrad and wgrad of one layer are independent, both
# are persistent kernels over all SMs, and in one stream the second cannot start before the LAST cluster of the first has
# finished.  Forked onto a second stream and joined before backward() returns, the clusters of the second kernel fill the
# SMs the first one's tail leaves idle (wave quantisation: 512->512 at 64^2 has 256 tiles for 74 clusters = 3.46 waves).
# Measured on B200 (bench.py, 512^2 B=8): 88.77 -> 88.16 ms per step.
_WGRAD_STREAMS = {}
wgrad_side_stream_enabled = True      # bench.py switches it off while it times every call with CUDA events


def _wgrad_overlap_mode() -> str:
    """How the weight-gradient GEMM of a conv overlaps the data-gradient GEMM of the same layer (VCD_WGRAD_OVERLAP):
    "stream" (default) the fork / join onto a side stream described above;
    "pdl"    one stream: workspace zeroed first, then dgrad, then wgrad as a programmatic dependent launch
             (VCD_WGRAD_OVERLAP_PREV, include/vcd.h) whose clusters fill the SMs dgrad's last wave leaves idle — no
             event record / wait per layer;
    "off"    strictly one after the other.
    Same-process A/B on B200 (bench.py VCD_BENCH_AB, 512^2 B=8, 3 x 15 steps each): stream 86.98-87.79 ms per step, pdl
    87.90-88.61, off 88.31-88.71 — the side stream also hides the HBM-bound bias-gradient column sums and the finalize
    kernels behind the tensor-bound dgrad, which a dependent launch on one stream cannot, so it stays the default.
    bench.py switches any overlap off (wgrad_side_stream_enabled) while it times every call with CUDA events."""
    import os
    if not wgrad_side_stream_enabled or os.environ.get("VCD_WGRAD_STREAM", "1") != "1":
        return "off"
    return os.environ.get("VCD_WGRAD_OVERLAP", "stream")


def _wgrad_side_stream(device):
    import os
    if not wgrad_side_stream_enabled or os.environ.get("VCD_WGRAD_STREAM", "1") != "1":
        return None
    if _wgrad_overlap_mode() != "stream":
        return None
    st = _WGRAD_STREAMS.get(device)
    if st is None:
        st = _WGRAD_STREAMS[device] = torch.cuda.Stream(device=device)
    return st


class _ConvFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, weight, bias, residual, packs: PackedWeights, stride: int, pad_t: int, pad_l: int,
                out_hw, impl: int, gn_groups: int = 0):
        x = _nhwc(x)
        N, H, W, Cin = x.shape
        Cout = weight.shape[0]
        KH, KW = (weight.shape[2], weight.shape[3]) if weight.dim() == 4 else (1, 1)
        Ho, Wo = out_hw
        wf, wd, b32 = packs.get(weight, bias, training=ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        umma = impl != IMPL_SIMT and _lib.lib().vcd_conv_umma_supported(Cin, Cout, KH, KW, stride) == 1
        planes = 0
        xs = x
        if residual is not None:
            residual = _nhwc(residual)
        y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
        ws = _workspace("vcd_conv2d_fprop_ws_bytes", (N, H, W, Cin, Cout, KH, KW, stride), impl, x.device)
        sums, zf = _acc_pool.take(N * gn_groups * 2, torch.float64, x.device) if gn_groups else (None, 0)
        call("vcd_conv2d_fprop", _p(xs), _p(wf), _p(b32), _p(residual), _p(y), _p(ws), N, H, W, Cin, Cout, KH, KW, stride,
             pad_t, pad_l, Ho, Wo, planes, impl | zf, _p(sums), gn_groups, _st())
        if sums is not None:
            push_gn_sums(y, sums, gn_groups)
        # x = act(GroupNorm(.)) with this conv as its only consumer: the dgrad epilogue will do the GroupNorm's reduction
        gi = _GN_FWD.pop(x.data_ptr(), None)
        ctx.gn_info = None
        if (gi is not None and gi[0].shape == x.shape and gi[0]._version == gi[-1] and umma and
                _lib.lib().vcd_conv2d_dgrad_gn_supported(N, H, W, Cin, Cout, KH, KW, stride) == 1):
            ctx.gn_info = gi[1:-1]
        ctx.save_for_backward(xs, weight, bias)
        ctx.packs = packs
        ctx.cfg = (N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, planes, impl)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, weight, bias = ctx.saved_tensors
        N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, planes, impl = ctx.cfg
        dy = _nhwc(dy)
        wf, wd, _ = ctx.packs.current()
        dx = dw = db = None
        need_w = ctx.needs_input_grad[1] or (bias is not None and ctx.needs_input_grad[2])
        side = _wgrad_side_stream(dy.device) if (need_w and ctx.needs_input_grad[0]) else None
        keep = None
        if side is not None:      # fork: the weight gradient runs concurrently with the data gradient below
            main = torch.cuda.current_stream(dy.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                # `keep`: tensors allocated on the MAIN stream that the side-stream kernels read (the bias-gradient column
                # sums).  They must stay referenced until the join below is enqueued — freed earlier, the caching allocator
                # would hand their memory to the main-stream allocations of this very backward (dx, dsdb), racing the read.
                dw, db, keep = _ConvFn._wgrad(ctx, xs, dy, weight, bias)
        pdl = None
        if side is None and need_w and ctx.needs_input_grad[0] and _wgrad_overlap_mode() == "pdl":
            pdl = _ConvFn._wgrad_alloc(ctx, dy, weight, bias)     # zeroes the workspace BEFORE the dgrad kernel
        if ctx.needs_input_grad[0]:
            if ctx.gn_info is not None:
                gx, gsums, ggamma, gbeta, geps, gact, ggroups = ctx.gn_info
                dx = torch.empty((N, H, W, Cin), dtype=torch.bfloat16, device=dy.device)
                dsdb = torch.empty(N * Cin * 2, dtype=torch.float32, device=dy.device)
                ab = torch.empty(N * Cin * 2, dtype=torch.float32, device=dy.device)
                gg, gb = ggamma.detach(), gbeta.detach()
                call("vcd_conv2d_dgrad_gn", _p(dy), _p(wd), _p(dx), N, H, W, Cin, Cout, KH, KW, pad_t, pad_l, _p(gx), _p(gsums),
                     _p(gg), _p(gb), dtype_code(gg), ggroups, geps, gact, _p(dsdb), _p(ab), _st())
                _GN_BWD[dx.data_ptr()] = (dx, dsdb, dx._version)
            else:
                dx = torch.empty((N, H, W, Cin), dtype=torch.bfloat16, device=dy.device)
                ws = _workspace("vcd_conv2d_dgrad_ws_bytes", (N, H, W, Cin, Cout, KH, KW, stride), impl, dy.device)
                call("vcd_conv2d_dgrad", _p(dy), _p(wf), _p(wd), _p(dx), _p(ws), N, H, W, Cin, Cout, KH, KW, stride, pad_t,
                     pad_l, Ho, Wo, 0, impl, _st())
        if side is not None:      # join: everything after this backward node is ordered after the weight gradient
            torch.cuda.current_stream(dy.device).wait_stream(side)
        elif need_w:
            dw, db, _ = _ConvFn._wgrad(ctx, xs, dy, weight, bias, pdl)
        del keep
        dres = dy if ctx.has_res and ctx.needs_input_grad[3] else None
        return dx, dw, db, dres, None, None, None, None, None, None, None

    @staticmethod
    def _wgrad_alloc(ctx, dy, weight, bias, prepare=True):
        """-> (dw, db, ws, flags): ws out of the zero-filled arena (flags = VCD_ACC_PREZEROED) when it fits a chunk;
        prepare: otherwise zero it now (the programmatic-dependent-launch path needs it zero BEFORE the dgrad kernel)"""
        N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, planes, impl = ctx.cfg
        dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        db = None if bias is None else torch.empty_like(bias)
        nbytes = _lib.lib().vcd_conv2d_wgrad_ws_bytes(N, H, W, Cin, Cout, KH, KW, stride)
        ws, zf = _ws_pool.take(nbytes // 4, torch.float32, dy.device)
        if prepare and not zf:
            call("vcd_conv2d_wgrad_prepare", _p(ws), Cin, Cout, KH, KW, _st())
        return dw, db, ws, zf

    @staticmethod
    def _wgrad(ctx, xs, dy, weight, bias, prepared=None):
        """prepared = (dw, db, ws) of _wgrad_alloc issued BEFORE the dgrad kernel that directly precedes this call in the
        stream: the GEMM is launched with VCD_WGRAD_OVERLAP_PREV and starts while that kernel's last wave drains."""
        N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, planes, impl = ctx.cfg
        dw, db, ws, zf = prepared if prepared is not None else _ConvFn._wgrad_alloc(ctx, dy, weight, bias, prepare=False)
        colsum = pop_colsum(dy) if db is not None else None
        call("vcd_conv2d_wgrad", _p(xs), _p(dy), _p(dw), _p(db), _p(colsum), dtype_code(weight), _p(ws), N, H, W, Cin, Cout,
             KH, KW, stride, pad_t, pad_l, Ho, Wo, planes,
             impl | zf | (_lib.WGRAD_OVERLAP_PREV if prepared is not None else 0), _st())
        return dw, db, colsum


def conv2d(x, weight, bias, packs, stride=1, pad_t=1, pad_l=1, out_hw=None, residual=None, impl=None, gn_groups=0):
    """gn_groups > 0: the output feeds a GroupNorm of that many groups — its sums come out of the GEMM epilogue"""
    if out_hw is None:
        out_hw = (x.shape[1], x.shape[2])
    return _ConvFn.apply(x, weight, bias, residual, packs, stride, pad_t, pad_l, tuple(out_hw),
                         _conv_impl if impl is None else impl, gn_groups)


class UpconvPackedWeights:
    """bf16 operand packs of an Upsample2D conv in its four-phase form (vcd_pack_upconv_weight):
    wf [16][Cout][Cin] for fprop, wd [16][Cin][Cout] for dgrad, fp32 bias."""

    def __init__(self):
        self.key = None
        self.frozen = False
        self.valid = False
        self.wf = self.wd = self.bias = None

    ensure = PackedWeights.ensure

    def get(self, weight: torch.Tensor, bias: Optional[torch.Tensor], training: bool = False):
        if self.valid:
            self.valid = False
            return self.wf, self.wd, self.bias
        key = (weight.data_ptr(), weight._version, weight.dtype, None if bias is None else (bias.data_ptr(), bias._version))
        if key != self.key or not self.frozen or training or torch.cuda.is_current_stream_capturing():
            _require_cuda(weight, "conv weight")
            w = weight.detach()
            if not w.is_contiguous():
                w = w.contiguous()
            cout, cin = w.shape[0], w.shape[1]
            n = 16 * cout * cin
            if self.wf is None or self.wf.numel() != n or self.wf.device != w.device:
                self.wf = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                self.wd = torch.empty(n, dtype=torch.bfloat16, device=w.device)
                self.bias = None if bias is None else torch.empty(cout, dtype=torch.float32, device=w.device)
            call("vcd_pack_upconv_weight", _p(w), _p(None if bias is None else bias.detach()), dtype_code(w), cout, cin,
                 _p(self.wf), _p(self.wd), _p(self.bias), _st())
            self.key = key
        return self.wf, self.wd, self.bias

    def current(self):
        return self.wf, self.wd, self.bias


def upconv_supported(cin: int, cout: int) -> bool:
    return cin % 128 == 0 and cout % 128 == 0 and _conv_impl != IMPL_SIMT


class _UpConvFn(torch.autograd.Function):
    """[upstream] Upsample2D: nearest x2 followed by conv3x3(pad 1), computed as four 2x2 phase convolutions on
    the low-resolution tensor (include/vcd.h vcd_upconv2d_*): the upsampled tensor is never materialised and the
    GEMMs do 16/36 of the multiply-adds."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, weight, bias, packs: UpconvPackedWeights, gn_groups: int = 0):
        x = _nhwc(x)
        N, H, W, Cin = x.shape
        Cout = weight.shape[0]
        wf, wd, b32 = packs.get(weight, bias, training=ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        y = torch.empty((N, 2 * H, 2 * W, Cout), dtype=torch.bfloat16, device=x.device)
        sums = torch.empty(N * gn_groups * 2, dtype=torch.float64, device=x.device) if gn_groups else None
        call("vcd_upconv2d_fprop", _p(x), _p(wf), _p(b32), _p(y), N, H, W, Cin, Cout, _p(sums), gn_groups, _st())
        if sums is not None:
            push_gn_sums(y, sums, gn_groups)
        ctx.save_for_backward(x, weight, bias)
        ctx.packs = packs
        ctx.cfg = (N, H, W, Cin, Cout)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        N, H, W, Cin, Cout = ctx.cfg
        dy = _nhwc(dy)
        wf, wd, _ = ctx.packs.current()
        colsum = pop_colsum(dy) if bias is not None else None
        dyp = dy   # the phase kernels read the parity planes of dy in place (element-strided TMA maps)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((N, H, W, Cin), dtype=torch.bfloat16, device=dy.device)
            call("vcd_upconv2d_dgrad", _p(dyp), _p(wd), _p(dx), N, H, W, Cin, Cout, _st())
        if ctx.needs_input_grad[1] or (bias is not None and ctx.needs_input_grad[2]):
            dw = torch.empty_like(weight, memory_format=torch.contiguous_format)
            db = None if bias is None else torch.empty_like(bias)
            nbytes = _lib.lib().vcd_upconv2d_wgrad_ws_bytes(Cin, Cout)
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dy.device)
            call("vcd_upconv2d_wgrad", _p(x), _p(dyp), _p(dw), _p(db), _p(colsum), dtype_code(weight), _p(ws), N, H, W,
                 Cin, Cout, _st())
        return dx, dw, db, None, None


def upconv2d(x, weight, bias, packs, gn_groups=0):
    return _UpConvFn.apply(x, weight, bias, packs, gn_groups)


# ------------------------------------------------------------------------------------------
# GroupNorm (+SiLU) with fused statistics
# ------------------------------------------------------------------------------------------
class _GroupNormFn(torch.autograd.Function):
    """GroupNorm [+SiLU].  With split=True the function also returns its input unchanged as a second output:
    the block routes its skip connection through it, so backward receives the skip gradient explicitly and
    adds it inside the dx kernel (no separate add pass) while emitting the column sums of dx (= the bias
    gradient of the conv that produced x)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, gamma, beta, groups: int, eps: float, act: bool, slot_in: Optional[TrackSlot],
                slot_out: Optional[TrackSlot], split: bool, sole_consumer_is_conv: bool = False,
                slot_in_extra: Optional[TrackSlot] = None):
        x = _nhwc(x)
        N, C = x.shape[0], x.shape[-1]
        hw = x.numel() // (N * C)
        sums = pop_gn_sums(x, groups)   # always popped: an entry left behind would pin x until the next encode()
        stats_in_apply = slot_in is not None and sums is not None
        if sums is None:   # not produced by the epilogue of the GEMM that wrote x: statistics pass (+ input statistics)
            sums = torch.empty(N * groups * 2, dtype=torch.float64, device=x.device)
            call("vcd_gn_stats", _p(x), _p(sums), _p(None if slot_in is None else slot_in.raw),
                 0.0 if slot_in is None else slot_in.near_zero, N, hw, C, groups, _st())
        out = torch.empty_like(x)
        g, b = gamma.detach(), beta.detach()
        # group sums from the producer's epilogue AND the input tracked: its per-channel statistics ride in the apply pass,
        # which reads x anyway (no extra pass over the tensor)
        nz = slot_out.near_zero if slot_out is not None else (slot_in.near_zero if slot_in is not None else 0.0)
        call("vcd_gn_apply_fwd", _p(x), _p(sums), _p(g), _p(b), dtype_code(g), _p(out),
             _p(slot_in.raw if stats_in_apply else None), _p(None if slot_out is None else slot_out.raw), nz,
             float(eps), 1 if act else 0, N, hw, C, groups, _st())
        if slot_in is not None:
            if slot_in_extra is not None:     # a second subscription to the same tensor (e.g. conv_in.output + norm1.input)
                slot_in_extra.raw.copy_(slot_in.raw)
                slot_in_extra.finalize(N * hw)
            slot_in.finalize(N * hw)
        if slot_out is not None:
            slot_out.finalize(N * hw)
        ctx.save_for_backward(x, sums, gamma, beta)
        ctx.cfg = (N, hw, C, groups, float(eps), 1 if act else 0)
        if sole_consumer_is_conv and x.requires_grad:
            _GN_FWD[out.data_ptr()] = (out, x, sums, gamma, beta, float(eps), 1 if act else 0, groups, out._version)
        if split:
            return out, x.view(x.shape)
        return out

    @staticmethod
    def backward(ctx, dout, dres=None):
        x, sums, gamma, beta = ctx.saved_tensors
        N, hw, C, G, eps, act = ctx.cfg
        dout = _nhwc(dout)
        if dres is not None:
            dres = _nhwc(dres)
        g, b = gamma.detach(), beta.detach()
        pdt = dtype_code(g)
        fused = _GN_BWD.pop(dout.data_ptr(), None)
        if fused is not None and fused[0].shape == dout.shape and fused[0]._version == fused[2]:
            # dout is already g = dL/d(pre-activation) and its channel sums were reduced by the conv's dgrad epilogue
            dsdb, act = fused[1], 0
        else:
            dsdb, zf = _acc_pool.take(N * C * 2, torch.float32, x.device)
            call("vcd_gn_bwd_reduce", _p(x), _p(dout), _p(sums), _p(g), _p(b), pdt, _p(dsdb), eps, act | zf, N, hw, C, G, _st())
        dx = None
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(beta)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            colsum, zf = _acc_pool.take(C, torch.float32, x.device)
            call("vcd_gn_bwd_apply", _p(x), _p(dout), _p(sums), _p(g), _p(b), pdt, _p(dsdb), _p(dx), _p(dres),
                 _p(colsum), _p(dgamma), _p(dbeta), eps, act | zf, N, hw, C, G, _st())   # also writes dgamma / dbeta
            push_colsum(dx, colsum)
        else:
            call("vcd_gn_param_grad", _p(sums), _p(dsdb), _p(dgamma), _p(dbeta), pdt, eps, N, hw, C, G, _st())
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None


def group_norm(x, gamma, beta, groups, eps, act, slot_in=None, slot_out=None, split=False, sole_consumer_is_conv=False,
               slot_in_extra=None):
    """sole_consumer_is_conv: the caller guarantees that the (first) output goes into exactly one ops.conv2d and nowhere
    else — that conv's dgrad then performs SiLU' and the backward reduction of this GroupNorm in its epilogue.
    slot_in / slot_in_extra: statistics slots of the INPUT tensor (slot_in_extra receives a copy of slot_in's sums)."""
    if slot_in is None and slot_in_extra is not None:
        slot_in, slot_in_extra = slot_in_extra, None
    return _GroupNormFn.apply(x, gamma, beta, groups, eps, act, slot_in, slot_out, split, sole_consumer_is_conv,
                              slot_in_extra)


# ------------------------------------------------------------------------------------------
# small element-wise ops
# ------------------------------------------------------------------------------------------
class _SiluFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, x):
        x = _nhwc(x)
        y = torch.empty_like(x)
        call("vcd_silu_fwd", _p(x), _p(y), x.numel(), _st())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _nhwc(dy)
        dx = torch.empty_like(x)
        call("vcd_silu_bwd", _p(x), _p(dy), _p(dx), x.numel(), _st())
        return dx


class _AddFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, a, b):
        a, b = _nhwc(a), _nhwc(b)
        o = torch.empty_like(a)
        call("vcd_add", _p(a), _p(b), _p(o), a.numel(), _st())
        return o

    @staticmethod
    def backward(ctx, d):
        return d, d


class _Upsample2xFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, x):
        x = _nhwc(x)
        N, H, W, C = x.shape
        y = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
        call("vcd_upsample2x_fwd", _p(x), _p(y), N, H, W, C, _st())
        ctx.shape = (N, H, W, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C = ctx.shape
        dy = _nhwc(dy)
        dx = torch.empty((N, H, W, C), dtype=dy.dtype, device=dy.device)
        call("vcd_upsample2x_bwd", _p(dy), _p(dx), N, H, W, C, _st())
        return dx


class _ToNHWC(torch.autograd.Function):
    """[N, C, H, W] (fp32 | bf16, any strides) -> bf16 NHWC."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x):
        _require_cuda(x, "model input")
        N, C, H, W = x.shape
        ctx.meta = (N, C, H, W, x.dtype)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        y = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=x.device)
        call("vcd_nchw_to_nhwc", _p(x), dtype_code(x), _p(y), N, C, H, W, _st())
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W, dt = ctx.meta
        dy = _nhwc(dy)
        odt = dt if dt in (torch.float32, torch.bfloat16) else torch.float32
        dx = torch.empty((N, C, H, W), dtype=odt, device=dy.device)
        call("vcd_nhwc_to_nchw", _p(dy), _p(dx), dtype_code(dx), N, C, H, W, _st())
        return dx.to(dt)


class _ToNCHW(torch.autograd.Function):
    """bf16 NHWC -> contiguous [N, C, H, W] in `dtype` (what train.py / evaluate.py consume)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, dtype):
        x = _nhwc(x)
        N, H, W, C = x.shape
        y = torch.empty((N, C, H, W), dtype=dtype, device=x.device)
        call("vcd_nhwc_to_nchw", _p(x), _p(y), dtype_code(y), N, C, H, W, _st())
        ctx.meta = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.meta
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        dy = dy.contiguous()
        dx = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=dy.device)
        call("vcd_nchw_to_nhwc", _p(dy), dtype_code(dy), _p(dx), N, C, H, W, _st())
        return dx, None


def silu(x):
    return _SiluFn.apply(x)


def add(a, b):
    return _AddFn.apply(a, b)


def upsample2x(x):
    return _Upsample2xFn.apply(x)


def to_nhwc(x):
    """logical [N, C, H, W] -> physical bf16 [N, H, W, C] (zero-copy when already bf16 channels-last)."""
    if x.dtype == torch.bfloat16 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1)
    return _ToNHWC.apply(x)


def to_nchw(x, dtype=torch.float32):
    return _ToNCHW.apply(x, dtype)


# ------------------------------------------------------------------------------------------
# attention core: softmax(Q K^T / sqrt(C)) V, one head (mid_block.attentions.0)
# ------------------------------------------------------------------------------------------
def _gemm_nt(A, B, D, batch, M, Nn, K, b_batched, alpha=1.0, bias=None, residual=None):
    call("vcd_gemm_nt", _p(A), _p(B), _p(bias), _p(residual), _p(D), batch, M, Nn, K, 1 if b_batched else 0,
         float(alpha), _st())


def _gemm_tn(A, B, D, batch, M, Nn, K, reduce_batch):
    nb = 1 if reduce_batch else batch
    ws = torch.empty(nb * M * Nn, dtype=torch.float32, device=A.device)
    call("vcd_gemm_tn", _p(A), _p(B), _p(D), dtype_code(D), _p(ws), batch, M, Nn, K, 1 if reduce_batch else 0, _st())


def _transpose(x, batch, rows, cols):
    y = torch.empty((batch, cols, rows), dtype=torch.bfloat16, device=x.device)
    call("vcd_transpose_bf16", _p(x), _p(y), batch, rows, cols, _st())
    return y


def _pad_tokens(x, N, T, Tp, C):
    """[N, T, C] -> zero-padded [N, Tp, C] (token counts that are not a multiple of the GEMM K step, e.g. 11 x 9 = 99)."""
    y = x.new_zeros((N, Tp, C))
    y[:, :T] = x.reshape(N, T, C)
    return y


class _AttnCoreFn(torch.autograd.Function):
    """softmax(Q K^T / sqrt(C)) V for ONE head of C channels over T = h*w tokens (diffusers Attention in the VAE mid blocks).
    T is padded to a multiple of 64 when it is not one (the P.V product contracts over tokens in 64-wide K steps): padded
    keys get a score of -inf, so their probabilities — and every gradient that flows through them — are exactly zero."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, q, k, v):
        q, k, v = _nhwc(q), _nhwc(k), _nhwc(v)
        shape = q.shape
        N, C = q.shape[0], q.shape[-1]
        T = q.numel() // (N * C)
        Tp = (T + 63) // 64 * 64
        if Tp != T:
            q, k, v = (_pad_tokens(t, N, T, Tp, C) for t in (q, k, v))
        scale = 1.0 / math.sqrt(C)
        dev = q.device
        s = torch.empty((N, Tp, Tp), dtype=torch.bfloat16, device=dev)
        _gemm_nt(q, k, s, N, Tp, Tp, C, True, alpha=scale)
        if Tp != T:
            s[:, :, T:] = float("-inf")
        p = torch.empty_like(s)
        call("vcd_softmax_fwd", _p(s), _p(p), N * Tp, Tp, _st())
        del s
        vt = _transpose(v, N, Tp, C)  # [N][C][Tp]
        o = torch.empty_like(q)
        _gemm_nt(p, vt, o, N, Tp, C, Tp, True)
        ctx.save_for_backward(q, k, v, p)
        ctx.meta = (N, T, Tp, C, scale, shape)
        return o if Tp == T else o[:, :T].reshape(shape)

    @staticmethod
    def backward(ctx, do):
        q, k, v, p = ctx.saved_tensors
        N, T, Tp, C, scale, shape = ctx.meta
        do = _nhwc(do)
        if Tp != T:
            do = _pad_tokens(do, N, T, Tp, C)
        dev = q.device
        dv = torch.empty_like(v)
        _gemm_tn(p, do, dv, N, Tp, C, Tp, False)           # dV[j][c] = sum_i P[i][j] dO[i][c]
        dp = torch.empty((N, Tp, Tp), dtype=torch.bfloat16, device=dev)
        _gemm_nt(do, v, dp, N, Tp, Tp, C, True)            # dP[i][j] = sum_c dO[i][c] V[j][c]
        ds = torch.empty_like(dp)
        call("vcd_softmax_bwd", _p(p), _p(dp), _p(ds), scale, N * Tp, Tp, _st())
        del dp
        kt = _transpose(k, N, Tp, C)                       # [N][C][Tp]
        dq = torch.empty_like(q)
        _gemm_nt(ds, kt, dq, N, Tp, C, Tp, True)           # dQ[i][c] = sum_j dS[i][j] K[j][c]
        dk = torch.empty_like(k)
        _gemm_tn(ds, q, dk, N, Tp, C, Tp, False)           # dK[j][c] = sum_i dS[i][j] Q[i][c]
        if Tp != T:
            dq, dk, dv = (t[:, :T].reshape(shape) for t in (dq, dk, dv))
        return dq, dk, dv


def attention_core(q, k, v):
    return _AttnCoreFn.apply(q, k, v)


# ------------------------------------------------------------------------------------------
# latent distribution and losses
# ------------------------------------------------------------------------------------------
class _GaussFn(torch.autograd.Function):
    """moments [N,h,w,8] (+ noise [N,4,h,w] fp32) -> z [N,h,w,4] bf16, kl [N] fp32, mean/logvar [N,4,h,w] fp32."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, moments, noise):
        moments = _nhwc(moments)
        N, h, w, C2 = moments.shape
        L = C2 // 2
        dev = moments.device
        z = torch.empty((N, h, w, L), dtype=torch.bfloat16, device=dev)
        kl = torch.empty(N, dtype=torch.float32, device=dev)
        mean = torch.empty((N, L, h, w), dtype=torch.float32, device=dev)
        logvar = torch.empty((N, L, h, w), dtype=torch.float32, device=dev)
        if noise is not None:
            noise = noise.to(torch.float32).contiguous()
        call("vcd_gauss_sample_kl_fwd", _p(moments), _p(noise), _p(z), _p(mean), _p(logvar), _p(kl), N, h * w, L, _st())
        ctx.save_for_backward(moments, noise)
        ctx.meta = (N, h * w, L)
        ctx.mark_non_differentiable(mean, logvar)
        return z, kl, mean, logvar

    @staticmethod
    def backward(ctx, dz, dkl, _dm, _dl):
        moments, noise = ctx.saved_tensors
        N, hw, L = ctx.meta
        if dz is not None:
            dz = _nhwc(dz)
        if dkl is not None:
            dkl = dkl.to(torch.float32).contiguous()
        dm = torch.empty_like(moments)
        call("vcd_gauss_sample_kl_bwd", _p(moments), _p(noise), _p(dz), _p(dkl), _p(dm), N, hw, L, _st())
        return dm, None


def gauss_sample_kl(moments, noise):
    return _GaussFn.apply(moments, noise)


class _MseFn(torch.autograd.Function):
    """mean((rec - x)^2) with rec bf16 NHWC and x the fp32 NCHW loader tensor (train.py:289)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, rec, x):
        rec = _nhwc(rec)
        N, H, W, C = rec.shape
        x = x.to(torch.float32).contiguous()
        loss = torch.empty(1, dtype=torch.float64, device=rec.device)
        drec = torch.empty_like(rec)
        numel = rec.numel()
        call("vcd_mse_fwd_bwd", _p(rec), _p(x), _p(loss), _p(drec), 1.0 / numel, N, C, H, W, _st())
        ctx.save_for_backward(drec)
        return (loss / numel).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        (drec,) = ctx.saved_tensors
        return drec * g.to(drec.dtype), None


def mse_loss(rec_nhwc, x_nchw):
    return _MseFn.apply(rec_nhwc, x_nchw)
