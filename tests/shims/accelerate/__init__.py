"""TEST SHIM (not product code): the slice of huggingface `accelerate` that the reference's src/train.py and
src/evaluate.py use, so those files run UNCHANGED in an image where accelerate is not installed.

Surface (call sites train.py:120-123,145,205-212,286-304,334,361,387,401,464 and evaluate.py:82,163,222-227):
Accelerator(gradient_accumulation_steps, mixed_precision, log_with, project_config); .state .device .num_processes
.process_index .is_main_process .is_local_main_process .mixed_precision .sync_gradients; prepare (model -> device + DDP +
bf16 autocast with fp32 outputs; optimizer / scheduler wrappers that step only on gradient-sync steps; dataloader ->
device placement + per-rank batch sharding), unwrap_model, accumulate, gather, backward, clip_grad_norm_, log,
save_state, wait_for_everyone, end_training.  Semantics follow accelerate's documented behaviour.

Opt-in used by this repo: VCD_FUSED_OPT=1 makes prepare() swap torch.optim.AdamW for the fused
multi-tensor clip+AdamW of libvcd_b200 with identical hyper-parameters; train.py itself stays untouched.
"""
from __future__ import annotations

import contextlib
import os
import pickle
import random

import numpy as np
import torch
import torch.distributed as dist

from . import logging as _logging  # noqa: F401
from .utils import ProjectConfiguration, set_seed  # noqa: F401

__version__ = "0.0-vcd-test-shim"


class _State:
    def __init__(self, acc):
        self._a = acc

    def __repr__(self):
        a = self._a
        return (f"Distributed environment: {'MULTI_GPU' if a.num_processes > 1 else 'NO'}\nNum processes: {a.num_processes}\n"
                f"Process index: {a.process_index}\nLocal process index: {a.local_process_index}\nDevice: {a.device}\n"
                f"Mixed precision type: {a.mixed_precision}\n")


def _to_fp32(obj):
    if isinstance(obj, torch.Tensor):
        return obj.float() if obj.is_floating_point() and obj.dtype != torch.float32 else obj
    if isinstance(obj, dict):
        return type(obj)((k, _to_fp32(v)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_fp32(v) for v in obj)
    return obj


def _to_device(obj, device, non_blocking):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, dict):
        return type(obj)((k, _to_device(v, device, non_blocking)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, device, non_blocking) for v in obj)
    return obj


class _ShardedLoader:
    """accelerate's DataLoaderShard + BatchSamplerShard (split_batches=False, even_batches=True): rank r of N takes
    batches r, r+N, ... of the underlying loader (the tail wraps around so every rank sees the same number of
    batches), tensors are moved to the accelerator's device, `end_of_dataloader` is raised on the last batch."""

    def __init__(self, loader, acc):
        self.loader, self.acc = loader, acc
        self.end_of_dataloader = False
        self.dataset = loader.dataset
        self.batch_size = loader.batch_size

    def __len__(self):
        n = len(self.loader)
        return (n + self.acc.num_processes - 1) // self.acc.num_processes

    def __iter__(self):
        a = self.acc
        self.end_of_dataloader = False
        a._active_loader = self
        nb = getattr(self.loader, "pin_memory", False) and a.device.type == "cuda"
        world, rank = a.num_processes, a.process_index

        def mine():
            first = []
            k = -1
            for k, batch in enumerate(self.loader):
                if len(first) < world:
                    first.append(batch)
                if k % world == rank:
                    yield batch
            rem = (k + 1) % world
            if world > 1 and rem and rank >= rem:            # uneven tail: wrap around to the first batches
                yield first[(rank - rem) % len(first)]

        prev = None
        for batch in mine():
            if prev is not None:
                yield _to_device(prev, a.device, nb)
            prev = batch
        self.end_of_dataloader = True
        if prev is not None:
            yield _to_device(prev, a.device, nb)


class _Optimizer:
    """accelerate's AcceleratedOptimizer: step / zero_grad only on gradient-synchronisation steps."""

    def __init__(self, opt, acc):
        self.optimizer, self._acc = opt, acc
        self._step_was_skipped = False

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @property
    def state(self):
        return self.optimizer.state

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd)

    def zero_grad(self, set_to_none=None):
        if self._acc.sync_gradients:
            self.optimizer.zero_grad(**({} if set_to_none is None else {"set_to_none": set_to_none}))

    def step(self, closure=None):
        if self._acc.sync_gradients:
            self.optimizer.step() if closure is None else self.optimizer.step(closure)

    @property
    def step_was_skipped(self):
        return self._step_was_skipped


class _Scheduler:
    """accelerate's AcceleratedScheduler (step_with_optimizer=True, split_batches=False): one underlying step per
    process on gradient-synchronisation steps."""

    def __init__(self, sched, acc):
        self.scheduler, self._acc = sched, acc

    def step(self, *a, **k):
        if not self._acc.sync_gradients:
            return
        for _ in range(self._acc.num_processes):
            self.scheduler.step(*a, **k)

    def get_last_lr(self):
        return self.scheduler.get_last_lr()

    def state_dict(self):
        return self.scheduler.state_dict()

    def load_state_dict(self, sd):
        self.scheduler.load_state_dict(sd)


class Accelerator:
    def __init__(self, gradient_accumulation_steps: int = 1, mixed_precision=None, log_with=None, project_config=None,
                 cpu: bool = False, **_):
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.mixed_precision = str(mixed_precision or os.environ.get("ACCELERATE_MIXED_PRECISION", "no"))
        self.project_configuration = project_config
        self.log_with = log_with
        self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
        self.process_index = int(os.environ.get("RANK", "0"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        use_cuda = torch.cuda.is_available() and not cpu
        if use_cuda:
            torch.cuda.set_device(self.local_process_index)
            self.device = torch.device("cuda", self.local_process_index)
        else:
            self.device = torch.device("cpu")
        if self.num_processes > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl" if use_cuda else "gloo", rank=self.process_index, world_size=self.num_processes,
                                    **({"device_id": self.device} if use_cuda else {}))
        self.state = _State(self)
        self.sync_gradients = True
        self.step = 0
        self._models, self._optimizers, self._schedulers = [], [], []
        self._active_loader = None
        self.trackers = []
        _logging._STATE["main"] = self.is_main_process
        _logging._STATE["ready"] = True

    # ---- process topology -------------------------------------------------------------------------
    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    def wait_for_everyone(self):
        if self.num_processes > 1:
            dist.barrier()

    # ---- prepare ----------------------------------------------------------------------------------
    def _prepare_model(self, model):
        model = model.to(self.device)
        model._original_forward = model.forward
        if self.mixed_precision in ("bf16", "fp16"):
            dt = torch.bfloat16 if self.mixed_precision == "bf16" else torch.float16
            inner = model.forward
            dev_type = self.device.type

            def forward(*a, **k):
                with torch.autocast(device_type=dev_type, dtype=dt):
                    out = inner(*a, **k)
                return _to_fp32(out)                         # accelerate: convert_outputs_to_fp32
            model.forward = forward
        if self.num_processes > 1:
            kw = {"device_ids": [self.local_process_index], "output_device": self.local_process_index} \
                if self.device.type == "cuda" else {}
            model = torch.nn.parallel.DistributedDataParallel(model, **kw)
        self._models.append(model)
        return model

    def _prepare_optimizer(self, opt):
        if os.environ.get("VCD_FUSED_OPT") == "1" and type(opt) is torch.optim.AdamW:
            import vcd_b200
            # shares opt.param_groups, so the LambdaLR train.py built on `opt` (train.py:202) still drives the lr
            opt = vcd_b200.FusedClipAdamW.from_torch(opt)
        w = _Optimizer(opt, self)
        self._optimizers.append(w)
        return w

    def prepare(self, *args):
        out = []
        for obj in args:
            if isinstance(obj, torch.nn.Module):
                out.append(self._prepare_model(obj))
            elif isinstance(obj, torch.optim.Optimizer):
                out.append(self._prepare_optimizer(obj))
            elif isinstance(obj, torch.utils.data.DataLoader):
                out.append(_ShardedLoader(obj, self))
            elif isinstance(obj, torch.optim.lr_scheduler.LRScheduler):
                s = _Scheduler(obj, self)
                self._schedulers.append(s)
                out.append(s)
            else:
                out.append(obj)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model, keep_fp32_wrapper: bool = True):
        while isinstance(model, torch.nn.parallel.DistributedDataParallel):
            model = model.module
        if not keep_fp32_wrapper and hasattr(model, "_original_forward"):
            model.forward = model._original_forward
        return model

    # ---- the training step ------------------------------------------------------------------------
    @contextlib.contextmanager
    def accumulate(self, *models):
        self.step += 1
        end = self._active_loader is not None and self._active_loader.end_of_dataloader
        self.sync_gradients = (self.step % self.gradient_accumulation_steps == 0) or end
        with contextlib.ExitStack() as stack:
            if not self.sync_gradients:
                for m in models:
                    if isinstance(m, torch.nn.parallel.DistributedDataParallel):
                        stack.enter_context(m.no_sync())
            yield

    def gather(self, tensor):
        if self.num_processes == 1:
            return tensor
        t = tensor.reshape(1) if tensor.dim() == 0 else tensor.contiguous()
        outs = [torch.empty_like(t) for _ in range(self.num_processes)]
        dist.all_gather(outs, t)
        return torch.cat(outs, dim=0)

    def backward(self, loss, **kw):
        if self.gradient_accumulation_steps > 1:
            loss = loss / self.gradient_accumulation_steps
        loss.backward(**kw)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        for w in self._optimizers:
            if hasattr(w.optimizer, "clip_grad_norm_"):      # fused clip+AdamW: the norm pass belongs to the optimizer
                return w.optimizer.clip_grad_norm_(parameters, max_norm)
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)

    # ---- logging / checkpoints --------------------------------------------------------------------
    def log(self, values, step=None, **_):
        pass

    def init_trackers(self, *a, **k):
        pass

    def save_state(self, output_dir=None, **_):
        """accelerate layout: model.safetensors, optimizer.bin, scheduler.bin, random_states_<rank>.pkl"""
        if output_dir is None:
            output_dir = os.path.join(self.project_configuration.project_dir, "checkpoints", "checkpoint_0")
        os.makedirs(output_dir, exist_ok=True)
        from safetensors.torch import save_file
        for i, m in enumerate(self._models):
            sd = {k: v.detach().contiguous().cpu() for k, v in self.unwrap_model(m).state_dict().items()}
            save_file(sd, os.path.join(output_dir, "model.safetensors" if i == 0 else f"model_{i}.safetensors"),
                      metadata={"format": "pt"})
        for i, o in enumerate(self._optimizers):
            torch.save(o.state_dict(), os.path.join(output_dir, "optimizer.bin" if i == 0 else f"optimizer_{i}.bin"))
        for i, s in enumerate(self._schedulers):
            torch.save(s.state_dict(), os.path.join(output_dir, "scheduler.bin" if i == 0 else f"scheduler_{i}.bin"))
        states = {"step": self.step, "random_state": random.getstate(), "numpy_random_seed": np.random.get_state(),
                  "torch_manual_seed": torch.get_rng_state()}
        if torch.cuda.is_available():
            states["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
        with open(os.path.join(output_dir, f"random_states_{self.process_index}.pkl"), "wb") as f:
            pickle.dump(states, f)
        return output_dir

    def end_training(self):
        if self.num_processes > 1 and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
