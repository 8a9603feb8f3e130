"""Data-parallel training step of the fusion network: flat parameter / gradient buckets, ONE
NCCL all-reduce per optimizer step, fused clip + AdamW + EMA kernel.

Mirrors the reference's cached training loop (``train.py:300-357``: forward_with_precomputed ->
clamp -> criterion -> backward -> clip_grad_norm_(1.0) -> AdamW.step -> EMA.update) with the
per-tensor ATen loops replaced by two kernels over a flat bucket (``csrc/optim.cu``) and, under
``torchrun``, one ``ncclAllReduce`` of the 1,433,217-float gradient bucket over NVLink/NVSwitch
(the reference has no DDP; an N-rank step equals its own ``accumulation_steps=N`` run over
the same N micro-batches, including per-micro-batch BatchNorm statistics).

``FusedAdamW`` is a ``torch.optim.Optimizer`` so LR schedulers and ``optimizer.state_dict()`` keep
working in an unchanged ``train.py``; parameters stay ordinary leaf ``nn.Parameter``s whose storage
is re-pointed into the flat bucket.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _cabi as K


class FlatBucket:
    """Flattens tensors into one contiguous fp32 buffer and re-points them at views of it."""

    def __init__(self, tensors: List[torch.Tensor], align: int = 4):
        self.shapes = [tuple(t.shape) for t in tensors]
        self.offsets, off = [], 0
        for t in tensors:
            self.offsets.append(off)
            off += (t.numel() + align - 1) // align * align
        self.numel = off
        dev = tensors[0].device if tensors else torch.device("cpu")
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        for t, o in zip(tensors, self.offsets):
            self.flat[o:o + t.numel()].view(t.shape).copy_(t.detach())

    def view(self, i: int) -> torch.Tensor:
        o, s = self.offsets[i], self.shapes[i]
        n = 1
        for d in s:
            n *= d
        return self.flat[o:o + n].view(s)

    def new_like(self) -> torch.Tensor:
        return torch.zeros_like(self.flat)


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_mean_(flat: torch.Tensor) -> None:
    """In-place mean over ranks (sum + scale; gloo has no AVG)."""
    w = _world()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / w)


def sync_bn_buffers(model: torch.nn.Module) -> None:
    """Per-rank BatchNorm statistics (no SyncBN in the reference); the running buffers are
    mean-all-reduced so every rank checkpoints the same state."""
    if _world() == 1:
        return
    bufs = [b for n, b in model.named_buffers() if n.endswith("running_mean") or n.endswith("running_var")]
    if not bufs:
        return
    flat = torch.cat([b.reshape(-1) for b in bufs])
    allreduce_mean_(flat)
    o = 0
    for b in bufs:
        b.copy_(flat[o:o + b.numel()].view_as(b))
        o += b.numel()


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over a flat bucket with optional fused global-norm clipping, data-parallel gradient
    all-reduce and EMA shadow update (one kernel pass: csrc/optim.cu).

    One documented difference from ``torch.optim.AdamW``: a parameter that received no gradient in a step has a
    zero gradient in the bucket (not ``None``), so it still gets weight decay and moment decay.  In the cached
    training loop every parameter receives a gradient every step (198/198, scripts/test_cached_training.py:212-218);
    the difference only shows when ``expert_feats`` is omitted and Phase 4 is skipped."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-4, max_grad_norm: float = 0.0, ema_decay: Optional[float] = None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdamW got no trainable parameters")
        if not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise RuntimeError("FusedAdamW (sm_100a build) needs fp32 CUDA parameters: there is no CPU path")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.lib = K.load()
        self.max_grad_norm = float(max_grad_norm)
        self.ema_decay = ema_decay
        self._params = params
        self.bucket = FlatBucket(params)
        self.grads = self.bucket.new_like()
        self.exp_avg = self.bucket.new_like()
        self.exp_avg_sq = self.bucket.new_like()
        self.ema = self.bucket.flat.clone() if ema_decay is not None else None
        dev = self.bucket.flat.device
        self._gsq = torch.zeros(1, device=dev, dtype=torch.float64)
        # step counter and learning rate live on the device so that a CUDA-graph replay sees fresh values
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self._lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self._lr_host = float(lr)
        self.steps = 0
        for i, p in enumerate(params):
            p.data = self.bucket.view(i)                       # same values, storage now inside the bucket
            o = self.bucket.offsets[i]
            p.grad = self.grads[o:o + p.numel()].view(p.shape)  # autograd accumulates in place into the bucket
        self.sync_from_rank0()

    @torch.no_grad()
    def sync_from_rank0(self) -> None:
        """Data-parallel replicas must start from (and resume with) ONE state: parameters, Adam moments, EMA shadow and
        step count are broadcast from rank 0 (what DistributedDataParallel does for parameters at construction).  Only
        gradients and BatchNorm statistics are reduced afterwards, so ranks that were seeded differently, or where only
        rank 0 loaded a checkpoint, would otherwise train diverging replicas silently.  No-op without a process group."""
        if _world() == 1:
            return
        for t in (self.bucket.flat, self.exp_avg, self.exp_avg_sq, self.ema):
            if t is not None:
                dist.broadcast(t, src=0)
        dist.broadcast(self._step_dev, src=0)
        self.steps = int(self._step_dev.item())
        self.bump_versions()

    @torch.no_grad()
    def reset_ema(self) -> None:
        """Re-initialise the EMA shadow from the current parameters (call after loading weights into the model once the
        optimizer exists: the shadow was taken from the construction-time weights)."""
        if self.ema is not None:
            self.ema.copy_(self.bucket.flat)

    # -- torch.optim API -------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        self.grads.zero_()
        for i, p in enumerate(self._params):                   # re-attach if a caller set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * self.bucket.offsets[i]:
                o = self.bucket.offsets[i]
                p.grad = self.grads[o:o + p.numel()].view(p.shape)

    def _gather_stray_grads(self):
        for i, p in enumerate(self._params):
            o = self.bucket.offsets[i]
            if p.grad is not None and p.grad.data_ptr() != self.grads.data_ptr() + 4 * o:
                self.grads[o:o + p.numel()].view(p.shape).copy_(p.grad)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._gather_stray_grads()
        g = self.param_groups[0]
        world = _world()
        if world > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)   # ONE collective: the flat 5.7 MB gradient bucket
        dev = self.bucket.flat.device
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:                    # a capture records the step without running it (the replay loop counts it)
            self.steps += 1
        if not capturing and float(g["lr"]) != self._lr_host:   # scheduler changed the LR: refresh the device scalar
            self._lr_host = float(g["lr"])
            self._lr_dev.fill_(self._lr_host)
        self._step_dev.add_(1)
        with torch.cuda.device(dev):
            S = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            n = self.bucket.numel
            gsq = None
            if self.max_grad_norm > 0:
                self._gsq.zero_()
                K.check(self.lib.ffsr_sumsq(self.grads.data_ptr(), n, self._gsq.data_ptr(), S), "sumsq")
                gsq = self._gsq.data_ptr()
            K.check(self.lib.ffsr_adamw_ema_step(
                self.bucket.flat.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                self.ema.data_ptr() if self.ema is not None else None, n, float(g["lr"]), float(g["betas"][0]),
                float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), 0, self._step_dev.data_ptr(),
                self._lr_dev.data_ptr(), gsq, 1.0 / world, self.max_grad_norm, float(self.ema_decay or 0.0), S),
                "adamw_ema_step")
        self.bump_versions()      # the kernel wrote through raw pointers: tell the version-keyed weight caches (pipeline.py)
        return loss

    # -- checkpointing (train.py saves optimizer.state_dict(): src/utils/checkpoint_manager.py:120-160) --------
    def state_dict(self):
        """param_groups as torch.optim does, plus the flat Adam moments, EMA shadow and step count."""
        sd = super().state_dict()
        sd["fused"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                       "ema": self.ema.detach().clone() if self.ema is not None else None, "steps": int(self.steps),
                       "offsets": list(self.bucket.offsets), "numel": int(self.bucket.numel)}
        return sd

    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        rest = {k: v for k, v in state_dict.items() if k != "fused"}
        per_param = rest.get("state") or {}
        if fused is None and per_param:
            # a torch.optim.AdamW checkpoint as the reference writes it (src/utils/checkpoint_manager.py:120-160):
            # state[i] = {step, exp_avg, exp_avg_sq} per parameter index -> scatter into the flat buckets
            if len(per_param) != len(self._params):
                raise ValueError(f"FusedAdamW.load_state_dict: checkpoint has state for {len(per_param)} parameters, "
                                 f"this optimizer owns {len(self._params)}")
            steps = 0
            for i, p in enumerate(self._params):
                st = per_param.get(i, per_param.get(str(i)))
                if st is None or tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"FusedAdamW.load_state_dict: no matching AdamW state for parameter {i}")
                o = self.bucket.offsets[i]
                self.exp_avg[o:o + p.numel()].view(p.shape).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
                steps = max(steps, int(st["step"]))
            rest = dict(rest)
            rest["state"] = {}
            super().load_state_dict(rest)
            self.steps = steps
            self._step_dev.fill_(steps)
            self.reset_ema()                 # a plain AdamW checkpoint carries no shadow: restart it from the weights
            self.set_lr(float(self.param_groups[0]["lr"]))
            self.sync_from_rank0()
            return
        super().load_state_dict(rest)
        if fused is not None:
            if fused["numel"] != self.bucket.numel or list(fused["offsets"]) != list(self.bucket.offsets):
                raise ValueError("FusedAdamW.load_state_dict: bucket layout differs from the checkpoint's")
            self.exp_avg.copy_(fused["exp_avg"])
            self.exp_avg_sq.copy_(fused["exp_avg_sq"])
            if self.ema is not None and fused.get("ema") is not None:
                self.ema.copy_(fused["ema"])
            elif self.ema is not None:
                self.reset_ema()             # checkpoint written with EMA disabled
            self.steps = int(fused["steps"])
            self._step_dev.fill_(self.steps)
        self.set_lr(float(self.param_groups[0]["lr"]))
        self.sync_from_rank0()

    # -- extras ----------------------------------------------------------------------------
    def set_lr(self, lr: float) -> None:
        """Scheduler hook usable between CUDA-graph replays (also picked up from param_groups[0]['lr'])."""
        self.param_groups[0]["lr"] = float(lr)
        self._lr_host = float(lr)
        self._lr_dev.fill_(float(lr))

    def bump_versions(self) -> None:
        for p in self._params:
            torch.autograd.graph.increment_version(p)

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm seen by the last step (device scalar, no sync)."""
        return self._gsq.sqrt() / _world()

    @torch.no_grad()
    def swap_ema(self) -> None:
        """Exchange the parameters with their EMA shadow in place (call again to undo).  ``EMAModel.save`` + ``apply``
        + ``restore`` (checkpoint_manager.py:359-377, train.py:450-452, 515) clone every parameter twice and copy 198
        tensors three times; with flat buckets the swap is three copies of one 5.7 MB buffer and no allocation beyond
        the temporary."""
        if self.ema is None:
            raise RuntimeError("EMA is disabled (ema_decay=None)")
        tmp = self.bucket.flat.clone()
        self.bucket.flat.copy_(self.ema)
        self.ema.copy_(tmp)
        self.bump_versions()                                   # packed / bf16 weight caches are keyed on _version

    def ema_shadow(self, names: List[str]) -> Dict[str, torch.Tensor]:
        """EMAModel.shadow-compatible dict (checkpoint_manager.py:344-350): name -> view of the flat shadow."""
        if self.ema is None:
            raise RuntimeError("EMA is disabled (ema_decay=None)")
        out = {}
        for i, nme in enumerate(names):
            o, s = self.bucket.offsets[i], self.bucket.shapes[i]
            out[nme] = self.ema[o:o + self._params[i].numel()].view(s)
        return out


class FusionTrainer:
    """One data-parallel training step (BASELINE configs[1] / configs[3]).

    ``cuda_graph=True`` (default): after ``graph_warmup`` eager steps the whole step -- forward, fused losses,
    backward, gradient all-reduce, clip + AdamW + EMA, BatchNorm buffer sync -- is captured once into a CUDA
    graph and replayed from static input buffers: the ~2,000 kernel launches of a step cost one graph launch
    instead of ~80 us of Python / driver time each.  The values that change between steps (optimizer step
    count, learning rate, dropout seed) are device-side scalars.  A change of input shapes re-captures."""

    def __init__(self, model, criterion, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-4, max_grad_norm: float = 1.0, ema_decay: Optional[float] = 0.999,
                 cuda_graph: bool = True, graph_warmup: int = 3):
        self.model, self.criterion = model, criterion
        self.names = [n for n, p in model.named_parameters() if p.requires_grad]
        self.optimizer = FusedAdamW(model.parameters(), lr, betas, eps, weight_decay, max_grad_norm, ema_decay)
        if _world() > 1:                      # BatchNorm statistics / counters start from rank 0's too (parameters: FusedAdamW)
            for n, b in model.named_buffers():
                if n.endswith(("running_mean", "running_var", "num_batches_tracked")):
                    dist.broadcast(b, src=0)
        self.optimizer.zero_grad()
        self.cuda_graph = cuda_graph
        self.graph_warmup = graph_warmup
        self._eager_steps = 0
        self._graph = None
        self._static = None
        self._sig = None

    def _step_body(self, lr_img, expert_imgs, expert_feats, hr_img):
        from .training import seed_counter
        sr = self.model.forward_with_precomputed(lr_img, expert_imgs, expert_feats).clamp(0, 1)   # train.py:326
        loss, comps = self.criterion(sr, hr_img, return_components=True)
        loss.backward()
        self.optimizer.step()
        sync_bn_buffers(self.model)
        self.optimizer.zero_grad()
        seed_counter(lr_img.device).add_(1)
        return loss.detach(), {k: v.detach() for k, v in comps.items()}

    def _signature(self, lr_img, expert_imgs, expert_feats, hr_img):
        """Everything a capture bakes in: input shapes / dtypes, and the host scalars and Python control flow of the step --
        the loss weights and component set (the curriculum calls ``criterion.set_weights`` every epoch, train.py:296), the
        optimizer hyper-parameters passed by value, the precision mode.  A change re-warms and re-captures."""
        f = expert_feats or {}
        crit, g = self.criterion, self.optimizer.param_groups[0]
        cw = tuple(sorted((k, float(v)) for k, v in getattr(crit, "weights", {}).items()))
        return (tuple(lr_img.shape), lr_img.dtype, tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in expert_imgs.items())),
                tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in f.items())), tuple(hr_img.shape),
                cw, bool(getattr(crit, "use_swt", True)), bool(getattr(crit, "use_fft", True)),
                float(g["weight_decay"]), tuple(float(b) for b in g["betas"]), float(g["eps"]),
                float(self.optimizer.max_grad_norm), self.optimizer.ema_decay, getattr(self.model, "precision", None))

    def _capture(self, lr_img, expert_imgs, expert_feats, hr_img):
        self._static = (lr_img.clone(), {k: v.clone() for k, v in expert_imgs.items()},
                        {k: v.clone() for k, v in expert_feats.items()} if expert_feats else None, hr_img.clone())
        self._graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(self._graph):
            self._out = self._step_body(*self._static)
        self._sig = self._signature(lr_img, expert_imgs, expert_feats, hr_img)

    def step(self, lr_img, expert_imgs, expert_feats, hr_img):
        """forward -> clamp -> loss -> backward -> all-reduce -> clip + AdamW + EMA.  Returns the
        (device) loss and its components; nothing here synchronises the host."""
        self.model.train()
        opt = self.optimizer
        if not self.cuda_graph:
            return self._step_body(lr_img, expert_imgs, expert_feats, hr_img)
        sig = self._signature(lr_img, expert_imgs, expert_feats, hr_img)
        if self._graph is not None and sig != self._sig:
            self._graph, self._static, self._eager_steps = None, None, 0       # new shapes: warm up and re-capture
        if self._graph is None:
            if self._eager_steps < self.graph_warmup:
                self._eager_steps += 1
                return self._step_body(lr_img, expert_imgs, expert_feats, hr_img)
            self._capture(lr_img, expert_imgs, expert_feats, hr_img)     # capture does not execute the step
        lr_now = float(opt.param_groups[0]["lr"])
        if lr_now != opt._lr_host:
            opt.set_lr(lr_now)
        s_lr, s_imgs, s_feats, s_hr = self._static
        s_lr.copy_(lr_img, non_blocking=True)
        for k, v in expert_imgs.items():
            s_imgs[k].copy_(v, non_blocking=True)
        if s_feats:
            for k, v in expert_feats.items():
                s_feats[k].copy_(v, non_blocking=True)
        s_hr.copy_(hr_img, non_blocking=True)
        self._graph.replay()
        opt.steps += 1
        opt.bump_versions()
        return self._out

    # ---- validation -----------------------------------------------------------------------------
    def ema_weights(self):
        """``with trainer.ema_weights():`` -- the model runs on its EMA weights inside the block."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            self.optimizer.swap_ema()
            try:
                yield self.model
            finally:
                self.optimizer.swap_ema()
        return cm()

    def validate(self, val_loader, crop_border: int = 4, test_y_channel: bool = True, use_ema: bool = True) -> Dict[str, float]:
        """``validate_epoch`` of the reference's cached mode (train.py:415-515) on this trainer's model."""
        return validate_epoch(self.model, val_loader, self.model.residual_scale.device, crop_border, test_y_channel,
                              self.optimizer if (use_ema and self.optimizer.ema is not None) else None)


@torch.no_grad()
def validate_epoch(model, val_loader, device, crop_border: int = 4, test_y_channel: bool = True,
                   ema: Optional[FusedAdamW] = None) -> Dict[str, float]:
    """Cached-mode ``validate_epoch`` (train.py:415-515): eval mode, EMA weights if given, full-image
    ``forward_with_precomputed`` per batch, clamp, Y-channel PSNR / SSIM with border crop -- with the metrics kept on
    the device (ONE host sync per epoch instead of two per image) and the EMA exchange done on the flat bucket.
    ``val_loader`` yields the reference's batch dict (host tensors from a DataLoader, fp16 allowed) or device batches
    (``cache.DeviceBatchLoader``).  The model's train / eval mode is restored on exit."""
    from .metrics import MetricCalculator
    was_training = model.training
    model.eval()
    if ema is not None:
        ema.swap_ema()
    calc = MetricCalculator(crop_border, test_y_channel)
    try:
        for batch in val_loader:
            lr_img = batch["lr"].to(device, non_blocking=True).float()
            hr_img = batch["hr"].to(device, non_blocking=True).float()
            imgs = {k: v.to(device, non_blocking=True).float() for k, v in batch["expert_imgs"].items()}
            feats = None
            if batch.get("expert_feats") is not None:
                feats = {k: v.to(device, non_blocking=True).float() for k, v in batch["expert_feats"].items()}
            sr = model.forward_with_precomputed(lr_img, imgs, feats).clamp(0, 1)
            calc.update(sr, hr_img)
    finally:
        if ema is not None:
            ema.swap_ema()
        model.train(was_training)
    return calc.get_metrics()
