"""Flat parameter arena: the HBM layout of the model state.

The reference keeps ~600 separate fp32 ``nn.Parameter`` tensors, lets autocast re-cast every
weight to bf16 on every use, lets autograd allocate ~600 gradient tensors and lets DDP copy
them into 25 MB buckets (cinema/device.py:101-103).  Here every parameter of the model is a
*view* into three flat device buffers:

    flat32  fp32 master weights   (what ``state_dict()`` / the optimiser see -- schema unchanged)
    flat16  bf16 shadow           (what the tcgen05 GEMMs read; refreshed by ONE cast kernel)
    gflat   fp32 gradients        (wgrad GEMMs ``red.add`` straight into it; ONE NCCL all-reduce)

Parameters that a kernel wants to read as one matrix (q / kv projection weights of a block, the
kv projections of all cross-attention decoder blocks) are placed back to back ("fusion groups"),
so a single GEMM with N = 3D (or N = depth * 2D) serves them without changing the reference's
state-dict layout (``attn.q.weight`` and ``attn.kv.weight`` stay separate keys).
"""

from __future__ import annotations

import weakref
from typing import Iterable

import torch
from torch import nn

from cinema_b200 import _C

ALIGN = 64  # elements; 256 B in fp32, 128 B in bf16


class ParamArena:
    def __init__(self, root: nn.Module, groups: Iterable[list[nn.Parameter]] = ()) -> None:
        params: list[nn.Parameter] = []
        seen: set[int] = set()
        for p in root.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise ValueError("module has no parameters")
        device = params[0].device
        if any(p.device != device for p in params):
            raise ValueError("all parameters of a model must live on one device")
        if any(p.dtype != torch.float32 for p in params):
            raise ValueError("cinema_b200 keeps fp32 master parameters (reference state-dict contract)")

        names = {id(p): n for n, p in root.named_parameters()}
        placed: set[int] = set()
        group_of: dict[int, list[nn.Parameter]] = {}  # first member -> group
        for g in groups:
            g = [p for p in g if id(p) in seen]
            if len(g) < 2 or any(id(p) in placed for p in g) or any(p.numel() % 8 for p in g[:-1]):
                continue  # cannot be laid out back to back with 16-byte aligned members; callers re-check adjacency
            if len({self.category(p, names[id(p)]) for p in g}) != 1:
                continue
            group_of[id(g[0])] = g
            placed.update(id(p) for p in g)
        # registration order, a fusion group sitting where its first member is registered: the parameters of a module
        # subtree then occupy ONE flat range per optimiser category, which is what the bucketed gradient all-reduce
        # (train.py) hands to NCCL as soon as that subtree's backward is done
        order: list[list[nn.Parameter]] = []
        for p in params:
            if id(p) in group_of:
                order.append(group_of[id(p)])
            elif id(p) not in placed:
                order.append([p])
        # optimiser regions: [weight-decayed | not decayed | frozen] so that AdamW is a few long flat launches
        order.sort(key=lambda g: self.category(g[0], names[id(g[0])]))
        self._names = names

        offsets: dict[int, int] = {}
        cur = 0
        for g in order:
            cur = (cur + ALIGN - 1) // ALIGN * ALIGN
            for p in g:
                offsets[id(p)] = cur
                cur += p.numel()
        total = (cur + ALIGN - 1) // ALIGN * ALIGN

        self.device = device
        self.numel = total
        self.flat32 = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat16 = torch.zeros(total, dtype=torch.bfloat16, device=device)
        self.gflat = torch.zeros(total, dtype=torch.float32, device=device)
        self.params = params
        self._w16: dict[int, torch.Tensor] = {}
        self._g32: dict[int, torch.Tensor] = {}
        self._off = offsets
        self.shadow_managed = False  # True while an optimiser kernel keeps flat16 in sync with flat32
        self._group_starts = []
        for g in order:
            self._group_starts.append((offsets[id(g[0])], g))
        with torch.no_grad():
            for p in params:
                off, n = offsets[id(p)], p.numel()
                view = self.flat32[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self._w16[id(p)] = self.flat16[off:off + n].view(p.shape)
                self._g32[id(p)] = self.gflat[off:off + n].view(p.shape)
                if p.grad is not None:
                    self._g32[id(p)].copy_(p.grad)
                    p.grad = self._g32[id(p)]
                p._cb_arena = weakref.ref(self)  # type: ignore[attr-defined]

    @staticmethod
    def category(p: nn.Parameter, name: str) -> int:
        """0: trainable with weight decay, 1: trainable without (1-D tensors and biases, the rule of timm's
        ``param_groups_weight_decay`` used at cinema/mae/pretrain.py:365), 2: frozen."""
        if not p.requires_grad:
            return 2
        return 1 if (p.ndim <= 1 or name.endswith(".bias")) else 0

    def ranges_of(self, params: Iterable[nn.Parameter]) -> list[tuple[int, int]]:
        """Flat [start, end) ranges (alignment padding included) covering exactly the given trainable parameters:
        consecutive parameters of the arena are merged, any other parameter in between splits the range."""
        want = {id(p) for p in params if p.requires_grad}
        layout = sorted(((self._off[id(p)], p) for p in self.params), key=lambda t: t[0])
        out: list[list[int]] = []
        open_run = False
        for off, p in layout:
            if id(p) in want:
                if open_run:
                    out[-1][1] = off + p.numel()
                else:
                    out.append([off, off + p.numel()])
                    open_run = True
            else:
                open_run = False
        return [(a, b) for a, b in out]

    def segments(self) -> list[tuple[int, int, int]]:
        """Maximal runs (start, end, category) of the arena under the CURRENT requires_grad flags."""
        runs: list[list[int]] = []
        for start, g in self._group_starts:
            cat = max(self.category(p, self._names[id(p)]) for p in g) if len(
                {self.category(p, self._names[id(p)]) for p in g}) > 1 else self.category(g[0], self._names[id(g[0])])
            if runs and runs[-1][2] == cat:
                continue
            if runs:
                runs[-1][1] = start
            runs.append([start, self.numel, cat])
        return [(a, b, c) for a, b, c in runs]

    # ------------------------------------------------------------------ queries
    def owns(self, p: nn.Parameter) -> bool:
        off = self._off.get(id(p))
        return off is not None and p.data_ptr() == self.flat32.data_ptr() + 4 * off

    def valid(self) -> bool:
        """True while the parameters still alias the arena (``.to()`` / ``assign=True`` loads break it)."""
        ps = self.params
        return self.owns(ps[0]) and self.owns(ps[-1]) and self.owns(ps[len(ps) // 2])

    def w16(self, p: nn.Parameter) -> torch.Tensor:
        return self._w16[id(p)]

    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        return self._g32[id(p)]

    def offset(self, p: nn.Parameter) -> int:
        return self._off[id(p)]

    def adjacent(self, ps: list[nn.Parameter]) -> bool:
        """Are the given parameters laid out back to back (so they can be read as one matrix)?"""
        for a, b in zip(ps[:-1], ps[1:]):
            if self._off[id(a)] + a.numel() != self._off[id(b)]:
                return False
        return True

    def fused16(self, ps: list[nn.Parameter], shape: tuple[int, ...]) -> torch.Tensor:
        off = self._off[id(ps[0])]
        n = sum(p.numel() for p in ps)
        return self.flat16[off:off + n].view(shape)

    def fused32(self, ps: list[nn.Parameter], shape: tuple[int, ...]) -> torch.Tensor:
        off = self._off[id(ps[0])]
        n = sum(p.numel() for p in ps)
        return self.flat32[off:off + n].view(shape)

    def fused_grad(self, ps: list[nn.Parameter], shape: tuple[int, ...]) -> torch.Tensor:
        off = self._off[id(ps[0])]
        n = sum(p.numel() for p in ps)
        return self.gflat[off:off + n].view(shape)

    # ------------------------------------------------------------------ per-step work
    def refresh_shadow(self) -> None:
        """fp32 master -> bf16 shadow: one HBM-bound kernel over the whole arena (6 B / parameter)."""
        if self.shadow_managed:
            return
        _C.cast_bf16(self.flat32, self.flat16)

    def prepare_grads(self) -> None:
        """Make every trainable parameter's ``.grad`` the arena view (zero-filled when it was ``None``),
        so the wgrad kernels can accumulate in place with the usual ``.grad +=`` semantics."""
        missing = [p for p in self.params if p.requires_grad and p.grad is None]
        n_train = sum(1 for p in self.params if p.requires_grad)
        if missing:
            if len(missing) == n_train:
                self.gflat.zero_()
            else:
                for p in missing:
                    self._g32[id(p)].zero_()
            for p in missing:
                p.grad = self._g32[id(p)]
        for p in self.params:
            if p.requires_grad and p.grad is not self._g32[id(p)] and p.grad.data_ptr() != self._g32[id(p)].data_ptr():
                self._g32[id(p)].copy_(p.grad)
                p.grad = self._g32[id(p)]


def arena_of(p: nn.Parameter) -> ParamArena | None:
    ref = getattr(p, "_cb_arena", None)
    a = ref() if ref is not None else None
    return a if a is not None and a.owns(p) else None


def ensure_arena(root: nn.Module) -> ParamArena:
    """Arena that owns the parameters of ``root``: the cached one, the enclosing model's, or a new one."""
    a = getattr(root, "_cb_arena_obj", None)
    if a is not None and a.valid():
        return a
    first = next(iter(root.parameters()), None)
    if first is None:
        raise ValueError("module has no parameters")
    a = arena_of(first)
    if a is not None and all(a.owns(p) for p in root.parameters()):
        return a
    groups = []
    for m in root.modules():
        fn = getattr(m, "_arena_groups", None)
        if fn is not None:
            groups.extend(fn())
    a = ParamArena(root, groups)
    object.__setattr__(root, "_cb_arena_obj", a)
    return a
