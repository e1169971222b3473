"""Data-parallel MAE pre-training step on B200: the step recipe of the reference's
``pretrain_one_epoch`` (cinema/mae/pretrain.py:242-272) with its runtime pieces re-designed:

  reference                                         here
  ------------------------------------------------  -------------------------------------------------
  batch.to(device) per view                         pinned host buffers -> static device buffers (async)
  autocast forward + autograd backward (~4k nodes)  one fused autograd node, replayed as a CUDA graph
  DDP reducer, 25 MB fp32 buckets, every micro-step flat gradient arena, 3 buckets (decoder / upper encoder / rest)
                                                    all-reduced by NCCL while the backward still runs
  GradScaler.unscale_ + clip_grad_norm_ + AdamW     sum-of-squares kernel + one fused AdamW kernel per
  (~10 passes over ~600 tensors)                    arena region (clip folded in, bf16 shadow refreshed)
  metrics .item() x 14 (host syncs)                 loss copied to pinned memory, read one step later

``cosine_lr`` restates ``adjust_learning_rate`` (cinema/optim.py:21-52).
"""

from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn

from cinema_b200 import _C
from cinema_b200.arena import ensure_arena


def cosine_lr(step: float, warmup_steps: float, max_n_steps: float, lr: float, min_lr: float) -> float:
    """Linear warm-up then half-cycle cosine decay (cinema/optim.py:21-52)."""
    if step < warmup_steps:
        return lr * step / warmup_steps
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (step - warmup_steps) / (max_n_steps - warmup_steps)))


def adjust_learning_rate(optimizer: torch.optim.Optimizer, step: float, warmup_steps: float, max_n_steps: float, lr: float,
                         min_lr: float) -> float:
    """Set the schedule's learning rate on a torch optimiser whose groups may carry ``lr_scale`` (layer-wise decay groups
    of ``convvit.param_groups_lr_decay``), as the fine-tuning scripts do (cinema/optim.py:21-52).  Returns the base lr."""
    cur = cosine_lr(step, warmup_steps, max_n_steps, lr, min_lr)
    for group in optimizer.param_groups:
        group["lr"] = cur * group.get("lr_scale", 1.0)
    return cur


def get_n_accum_steps(batch_size: int, batch_size_per_device: int, world_size: int) -> int:
    """Gradient-accumulation factor for an effective batch size (cinema/optim.py:122-143), same checks and errors."""
    per_step = batch_size_per_device * world_size
    if per_step > batch_size:
        raise ValueError(f"batch_size_per_step {per_step} should be less than batch_size {batch_size}.")
    if batch_size % per_step != 0:
        raise ValueError(f"batch_size {batch_size} should be divisible by batch_size_per_step {per_step}.")
    return batch_size // per_step


def allreduce_gradients(model: nn.Module, process_group=None, average: bool = True) -> None:
    """Data-parallel gradient exchange for the EAGER loops (fine-tuning ``ConvViT`` / ``ConvUNetR``, eager ``CineMA``):
    the replacement for wrapping the model in ``DistributedDataParallel`` (cinema/device.py:86-104).

    The fused autograd nodes of this package write parameter gradients straight into the flat gradient arena (every
    ``p.grad`` is a view of it), so ``AccumulateGrad`` hooks never fire for the backbone and DDP's reducer would never
    see those gradients: **do not wrap these models in DDP**.  Call this after ``loss.backward()`` and before
    ``optimizer.step()`` instead -- ONE sum all-reduce over the whole arena (heads included), divided by the world size
    like DDP's mean.  With gradient accumulation call it on the last micro-step only.  No-op on one rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(process_group)
    if world == 1:
        return
    arena = ensure_arena(model)
    arena.prepare_grads()
    dist.all_reduce(arena.gflat, op=dist.ReduceOp.SUM, group=process_group)
    if average:
        arena.gflat.div_(world)


class FlatAdamW:
    """AdamW over the flat arena with global-norm clipping (cinema/mae/pretrain.py:365-367, cinema/optim.py:204-212)."""

    def __init__(self, arena, lr: float, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.05,
                 clip_grad: float | None = 5.0) -> None:
        self.arena = arena
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.clip_grad = clip_grad if clip_grad and clip_grad > 0 else 0.0
        dev = arena.device
        self.m = torch.zeros_like(arena.flat32)
        self.v = torch.zeros_like(arena.flat32)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        # step() never synchronises with the host, which may therefore run several steps ahead of the device: every
        # step's scalars get their own pinned slot, and a slot is rewritten only after the async copy that read it has
        # completed (CUDA event per slot), so an in-flight H2D copy can never pick up a later step's values
        self._hyper_ring = [torch.zeros(4, dtype=torch.float32) for _ in range(self.HYPER_SLOTS)]
        self._hyper_done: list = [None] * self.HYPER_SLOTS
        if dev.type == "cuda":
            self._hyper_ring = [t.pin_memory() for t in self._hyper_ring]
        self.t = 0
        self.segments = [s for s in arena.segments() if s[2] != 2]

    HYPER_SLOTS = 8

    def set_step_scalars(self) -> None:
        """Host -> device copy of {lr, 1 - beta1^t, 1 - beta2^t} for the NEXT ``apply`` (outside any CUDA graph).
        Note: ``t`` counts updates issued by the host; a step the device skips for a non-finite gradient norm
        (GradScaler semantics, csrc/optim.cu) still advances the bias corrections, unlike torch's GradScaler + AdamW."""
        self.t += 1
        slot = self.t % self.HYPER_SLOTS
        if self._hyper_done[slot] is not None:
            self._hyper_done[slot].synchronize()  # the copy issued HYPER_SLOTS steps ago has read this slot
        host = self._hyper_ring[slot]
        host[0] = self.lr
        host[1] = 1.0 - self.betas[0] ** self.t
        host[2] = 1.0 - self.betas[1] ** self.t
        self.hyper.copy_(host, non_blocking=True)
        if self.hyper.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.hyper.device))
            self._hyper_done[slot] = ev

    def apply(self, grad_scale: float = 1.0) -> None:
        """Norm + clip + AdamW + bf16 shadow refresh; graph-capturable (no host-dependent values)."""
        a = self.arena
        self.gnorm_sq.zero_()
        for s, e, _ in self.segments:
            _C.sumsq(a.gflat[s:e], self.gnorm_sq)
        for s, e, cat in self.segments:
            _C.adamw_flat(a.flat32[s:e], a.gflat[s:e], self.m[s:e], self.v[s:e], a.flat16[s:e], self.hyper, self.betas[0],
                          self.betas[1], self.eps, self.weight_decay if cat == 0 else 0.0, self.gnorm_sq, self.clip_grad,
                          grad_scale)

    def grad_norm(self, grad_scale: float = 1.0) -> torch.Tensor:
        return self.gnorm_sq.sqrt() * grad_scale

    def state_dict(self, names: dict[int, str] | None = None) -> dict:
        """Optimiser state per PARAMETER NAME (layout-independent: the arena order may differ between runs / versions):
        {"step", "exp_avg": {name: tensor}, "exp_avg_sq": {name: tensor}} -- the content of torch AdamW's state
        (cinema/optim.py:229-260 saves ``optimizer.state_dict()``)."""
        a = self.arena
        names = names or a._names
        out = {"step": self.t, "exp_avg": {}, "exp_avg_sq": {}}
        for p in a.params:
            if not p.requires_grad:
                continue
            off, n = a.offset(p), p.numel()
            out["exp_avg"][names[id(p)]] = self.m[off:off + n].view(p.shape).clone()
            out["exp_avg_sq"][names[id(p)]] = self.v[off:off + n].view(p.shape).clone()
        return out

    def load_state_dict(self, state: dict) -> None:
        a = self.arena
        by_name = {a._names[id(p)]: p for p in a.params}
        unknown = set(state["exp_avg"]) - set(by_name)
        if unknown:
            raise KeyError(f"optimizer state for unknown parameters: {sorted(unknown)[:5]}")
        self.t = int(state["step"])
        for name, t in state["exp_avg"].items():
            p = by_name[name]
            off, n = a.offset(p), p.numel()
            self.m[off:off + n].copy_(t.reshape(-1))
            self.v[off:off + n].copy_(state["exp_avg_sq"][name].reshape(-1))


class MAETrainer:
    """One object = one rank.  ``step(batch)`` runs H2D -> forward -> backward -> all-reduce -> clip -> AdamW."""

    def __init__(self, model: nn.Module, *, lr: float = 1e-3, betas=(0.9, 0.95), weight_decay: float = 0.05,
                 clip_grad: float | None = 5.0, enc_mask_ratio: float = 0.75, use_cuda_graph: bool = True,
                 process_group=None, graph_warmup: int = 2, overlap_allreduce: bool | None = None,
                 n_accum_steps: int = 1) -> None:
        """``n_accum_steps`` > 1: gradients of that many consecutive ``step`` calls are summed in the arena and the
        all-reduce + clip + AdamW run on the last one with the mean (cinema/mae/pretrain.py:258-267; unlike DDP, which
        all-reduces on every micro-step, the exchange happens once per update)."""
        if n_accum_steps < 1:
            raise ValueError(f"n_accum_steps must be >= 1, got {n_accum_steps}")
        self.n_accum = n_accum_steps
        self._micro = 0
        self.updated = False  # did the last ``step`` call apply an optimiser update?
        self.model = model
        self.ratio = enc_mask_ratio
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.arena = ensure_arena(model)
        if self.world > 1:  # rank 0's initial weights everywhere (cinema/mae/pretrain.py:345-361)
            dist.broadcast(self.arena.flat32, src=0, group=self.pg)
        self.arena.shadow_managed = False
        self.arena.refresh_shadow()
        self.arena.shadow_managed = True
        model.register_load_state_dict_post_hook(lambda *_: self.sync_shadow())
        self.arena.gflat.zero_()
        self.arena.prepare_grads()
        self.opt = FlatAdamW(self.arena, lr, betas, 1e-8, weight_decay, clip_grad)
        # Bucketed gradient all-reduce overlapped with the backward pass: the model reports when the decoder subtree and
        # the upper half of the encoder are final (mae._grad_stage_done); their flat ranges go to NCCL asynchronously
        # while the rest of the backward runs, the remainder follows after the backward.  All of it is inside the
        # forward/backward region, so it is captured in the CUDA graph together with the kernels.
        self._stage_ranges: dict[str, list[tuple[int, int]]] = {}
        self._rest_ranges: list[tuple[int, int]] = [(0, self.arena.numel)]
        self._pending: list = []
        import os

        # forward + backward without the autograd engine (CineMA.train_step) when the model supports it
        self._direct = bool(getattr(model, "direct_step_supported", lambda: False)()) and os.environ.get("CB_DIRECT_STEP", "1") == "1"
        if overlap_allreduce is None:  # CB_OVERLAP_ALLREDUCE=0 turns the bucketed overlap off
            overlap_allreduce = os.environ.get("CB_OVERLAP_ALLREDUCE", "1") == "1" and self._direct
        # CB_FORCE_SPLIT=1: split the step's graph at the gradient stages even on one rank (tests of the capture logic)
        self._force_split = os.environ.get("CB_FORCE_SPLIT", "0") == "1"
        self.overlap = ((self.world > 1 or self._force_split) and overlap_allreduce and hasattr(model, "dec_linear")
                        and self.n_accum == 1)
        if self.overlap:
            from cinema_b200.mae import grad_stages

            covered: list[nn.Parameter] = []
            for name, params in grad_stages(model).items():
                self._stage_ranges[name] = self.arena.ranges_of(params)
                covered += [p for p in params if p.requires_grad]
            ids = {id(p) for p in covered}
            self._rest_ranges = self.arena.ranges_of([p for p in self.arena.params if id(p) not in ids])
            model._grad_stage_hook = self._on_stage
        # Optional experiment (CB_NCCL_BG_CTAS=n > 0): overlapped buckets go through a SECOND communicator limited to n
        # CTAs, the tail bucket stays on the default one.  Measured at 2 x B200 (profiles/r02_scaling.md): 26.69 ms with
        # n = 4 and 29.21 ms with n = 2 against 26.09 ms on the default communicator -- the throttled all-reduce of the
        # last overlapped bucket no longer finishes under the backward -- so it is OFF by default.
        self._bg_pg = None
        bg_ctas = int(os.environ.get("CB_NCCL_BG_CTAS", "0"))
        if self.overlap and self.world > 1 and bg_ctas > 0 and self.arena.device.type == "cuda" and self.pg is None:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = bg_ctas
                opts.config.min_ctas = 1
                self._bg_pg = dist.new_group(backend="nccl", pg_options=opts)
            except Exception:  # noqa: BLE001 -- older torch / non-NCCL backends: fall back to the default communicator
                self._bg_pg = None
        self.use_graph = use_cuda_graph and self.arena.device.type == "cuda"
        self.graph_warmup = graph_warmup
        self._n_calls = 0
        self._inputs: dict[str, torch.Tensor] | None = None
        self._prefetched = None
        self._copy_stream = None
        self._staged = None
        self._staged_free = None
        self.prefetch_misses = 0  # step() calls that found a prefetched batch other than the one they were given
        self._g_fb = self._g_opt = None
        self._fb_segments: list = []  # [(graph, flat ranges to all-reduce once it has run)] when the capture is split
        self._capturing = False
        self._cur_graph = None
        self._loss = None
        self._loss_host = None
        self.launches_per_step = 0

    # ------------------------------------------------------------------ pieces
    def sync_shadow(self) -> None:
        self.arena.shadow_managed = False
        self.arena.refresh_shadow()
        self.arena.shadow_managed = True

    def set_lr(self, lr: float) -> None:
        self.opt.lr = lr

    def _stage(self, batch: dict[str, torch.Tensor]) -> None:
        dev = self.arena.device
        if self._inputs is None:
            self._inputs = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
        if self._prefetched is not None and self._prefetched[0] is batch:
            # the batch was uploaded ahead of time (prefetch): one device-to-device copy into the graph's input buffers
            _, staged, ready = self._prefetched
            self._prefetched = None
            torch.cuda.current_stream(dev).wait_event(ready)
            for k, v in staged.items():
                self._inputs[k].copy_(v, non_blocking=True)
            self._staged_free.record(torch.cuda.current_stream(dev))
            return
        if self._prefetched is not None:
            # a DIFFERENT batch than the prefetched one: the staged copy is dropped (the hand-off is by object identity, as a
            # loader that yields each pinned batch once would use it) and this batch takes the in-line upload below
            self._prefetched = None
            self.prefetch_misses += 1
        for k, v in batch.items():
            self._inputs[k].copy_(v, non_blocking=True)

    def prefetch(self, batch: dict[str, torch.Tensor]) -> None:
        """Start the host-to-device upload of a (pinned) batch on a copy stream, so that it overlaps the step in flight;
        the next ``step(batch)`` called with the SAME dict object consumes the uploaded copy.  The device-side role of the
        reference's ``DataLoader(pin_memory=True)`` + ``.to(device, non_blocking=True)`` (cinema/mae/pretrain.py:226-249)."""
        dev = self.arena.device
        if dev.type != "cuda":
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
            self._staged_free = torch.cuda.Event()
            self._staged_free.record(torch.cuda.current_stream(dev))
        ready = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._staged_free)  # the previous staged batch has been consumed
            for k, v in batch.items():
                self._staged[k].copy_(v, non_blocking=True)
            ready.record(self._copy_stream)
        self._prefetched = (batch, self._staged, ready)

    def _reduce_async(self, ranges, tail: bool = False) -> None:
        """Hand flat gradient ranges to NCCL on its own stream (ordered after everything queued on the current stream so
        far); the caller joins with ``_join_reduces`` before the optimiser reads the gradients.  ``tail``: the bucket
        that follows the backward (full-speed communicator); the others use the CTA-limited background communicator."""
        if self.world > 1:
            pg = self.pg if (tail or self._bg_pg is None) else self._bg_pg
            for s, e in ranges:
                self._pending.append(dist.all_reduce(self.arena.gflat[s:e], op=dist.ReduceOp.SUM, group=pg, async_op=True))

    def _join_reduces(self) -> None:
        for w in self._pending:
            w.wait()
        self._pending = []

    def _on_stage(self, stage: str) -> None:
        """Called by the model in the middle of the backward when a parameter subtree's gradients are final."""
        ranges = self._stage_ranges.get(stage, ())
        if self._capturing:
            # end the current graph here and continue the capture in a new one: at replay time the all-reduce of this
            # stage is launched between the two graphs and overlaps the second one -- no NCCL call is ever captured
            g = self._cur_graph
            g.capture_end()
            self._fb_segments.append((g, ranges))
            g2 = torch.cuda.CUDAGraph()
            g2.capture_begin(pool=self._fb_segments[0][0].pool())
            self._cur_graph = g2
        else:
            self._reduce_async(ranges)

    def _compute(self) -> None:
        """Forward + backward into the gradient arena; the stage hooks fire inside."""
        if self.n_accum == 1:
            self.arena.gflat.zero_()  # (with accumulation the first micro-step of an update clears it, outside the graph)
        if self._direct:
            self._loss = self.model.train_step(self._inputs, self.ratio)
        else:
            loss, _, _, _ = self.model(self._inputs, self.ratio)
            loss.backward()
            self._loss = loss.detach()

    def _fwd_bwd(self) -> None:
        self._pending = []
        self._compute()
        if self.overlap:  # the rest of the arena, then join every bucket before the optimiser reads the gradients
            self._reduce_async(self._rest_ranges, tail=True)
            self._join_reduces()

    def _replay_fwd_bwd(self) -> None:
        if not self._fb_segments:
            self._g_fb.replay()
            return
        self._pending = []
        last = len(self._fb_segments) - 1
        for j, (g, ranges) in enumerate(self._fb_segments):
            g.replay()
            self._reduce_async(ranges, tail=j == last)
        self._join_reduces()

    def _reduce(self) -> None:
        if self.world > 1 and not self.overlap:
            dist.all_reduce(self.arena.gflat, op=dist.ReduceOp.SUM, group=self.pg)

    def _opt_apply(self) -> None:
        self.opt.apply(grad_scale=1.0 / (self.world * self.n_accum))

    def _begin_micro(self) -> bool:
        """Bookkeeping of gradient accumulation; returns True when this call ends with an optimiser update."""
        if self.n_accum > 1 and self._micro == 0:
            self.arena.gflat.zero_()
        self._micro = (self._micro + 1) % self.n_accum
        self.updated = self._micro == 0
        if self.updated:
            self.opt.set_step_scalars()
        return self.updated

    # ------------------------------------------------------------------ the step
    def step(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        """batch: {view: (B, C, *spatial)} host (pinned) or device tensors.  Returns the loss as a 0-d DEVICE tensor
        (valid until the next call); nothing here synchronises with the host."""
        self._stage(batch)
        update = self._begin_micro()
        if not self.use_graph or self._n_calls < self.graph_warmup:
            n0 = _C.launches
            self._fwd_bwd()
            if update:
                self._reduce()
                self._opt_apply()
            self.launches_per_step = _C.launches - n0
        else:
            if self._g_fb is None:
                self._capture()
            self._replay_fwd_bwd()
            if update:
                self._reduce()
                self._g_opt.replay()
        self._n_calls += 1
        return self._loss

    def eager_step(self, batch: dict[str, torch.Tensor]) -> torch.Tensor:
        """The same step without CUDA-graph replay (used for per-kernel instrumentation)."""
        self._stage(batch)
        if self._begin_micro():
            self._fwd_bwd()
            self._reduce()
            self._opt_apply()
        else:
            self._fwd_bwd()
        return self._loss

    def _capture_split(self) -> None:
        """Capture forward + backward as a chain of CUDA graphs cut at the gradient stages (decoder subtree final, upper
        encoder half final, rest).  ``train_step`` runs on this thread, so ``_on_stage`` can end one capture and begin
        the next on the same stream; all graphs share one memory pool and are always replayed in capture order."""
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        dev = self.arena.device
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        self._fb_segments = []
        with torch.cuda.stream(stream):
            g = torch.cuda.CUDAGraph()
            g.capture_begin()
            self._cur_graph, self._capturing = g, True
            try:
                self._compute()
            finally:
                self._capturing = False
                self._cur_graph.capture_end()
            self._fb_segments.append((self._cur_graph, self._rest_ranges))
        torch.cuda.current_stream(dev).wait_stream(stream)
        self._cur_graph = None
        self._g_fb = self._fb_segments[0][0]

    # ------------------------------------------------------------------ checkpoint (cinema/optim.py:229-294)
    def state_dict(self) -> dict:
        """{"model": reference-schema state dict, "optimizer": per-name AdamW moments + step}."""
        return {"model": {k: v.detach().clone() for k, v in self.model.state_dict().items()}, "optimizer": self.opt.state_dict()}

    def load_state_dict(self, state: dict) -> None:
        self.model.load_state_dict(state["model"])  # the post-hook refreshes the bf16 shadow
        self.opt.load_state_dict(state["optimizer"])
        self._micro = 0

    def _capture(self) -> None:
        torch.cuda.synchronize()
        if self.overlap and self._direct:
            self._capture_split()
        else:
            self._g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_fb):
                self._fwd_bwd()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt, pool=self._g_fb.pool()):
            self._opt_apply()
        # the capture pass recorded, it did not run: play the step that was asked for
        # (callers replay right after)
