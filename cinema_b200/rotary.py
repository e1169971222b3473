"""Rotary position embedding with the standalone contract of the reference (cinema/rotary.py).

``RotaryEmbedding.forward(q, k, offset)`` takes (batch, n_tokens, n_heads, head_dim) tensors, caches
cos / sin tables in the activation dtype and rotates the first ``dim`` channels NeoX-style; the
rotation itself is one elementwise CUDA kernel (``cb_rope_apply``), its backward the same kernel
with the sine negated.

Note: inside the reference ``Attention`` the module is handed (B, H, N, d) tensors, so
the angle is indexed by the head and the rotation cancels in q.k^T; ``cinema_b200.vit.Attention``
therefore reproduces the observable behaviour (no rotation) and does not call this module.
"""

from __future__ import annotations

import torch

from cinema_b200 import _C


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """(x1, x2) -> (-x2, x1) over the last axis (cinema/rotary.py:12-22).  Host-side helper for callers/tests."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin):
        ctx.save_for_backward(cos, sin)
        ctx.dt = x.dtype
        xin = x.detach()
        if xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        return _C.rope_apply(xin.contiguous(), cos, sin).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        cos, sin = ctx.saved_tensors
        gin = g.detach()
        if gin.dtype not in (torch.float32, torch.bfloat16):
            gin = gin.float()
        return _C.rope_apply(gin.contiguous(), cos, sin, transpose=True).to(ctx.dt), None, None


def apply_rotary_emb(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x (B, N, H, d); cos / sin (>= N, rotary_dim / 2) (cinema/rotary.py:25-50)."""
    ro_dim = cos.shape[-1] * 2
    if ro_dim > x.shape[-1]:
        raise ValueError(f"Rotary dim {ro_dim} is larger than the last dimension of x {x.shape[-1]}")
    return _RopeFn.apply(x, cos.float().contiguous(), sin.float().contiguous())


class RotaryEmbedding(torch.nn.Module):
    """cinema/rotary.py:53-128."""

    def __init__(self, dim: int, base: float = 10000.0, scaling_factor: float = 1.0, device=None) -> None:
        super().__init__()
        self.dim = dim
        self.base = float(base)
        self.scaling_factor = scaling_factor
        self.device = device
        self.n_tokens = 0
        self.cos = None
        self.sin = None
        inv_freq = 1 / (self.base ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)

    def update_cos_sin(self, n_tokens: int, device: torch.device, dtype: torch.dtype) -> None:
        stale = (self.cos is None or n_tokens > self.n_tokens or self.cos.device != device or self.cos.dtype != dtype
                 or (self.training and self.cos.is_inference()))
        if stale:
            self.n_tokens = n_tokens
            t = torch.arange(n_tokens, device=device, dtype=self.inv_freq.dtype) / self.scaling_factor
            freqs = torch.outer(t, self.inv_freq.to(device))
            self.cos = torch.cos(freqs).to(dtype)  # tables are rounded to the activation dtype, as in the reference
            self.sin = torch.sin(freqs).to(dtype)

    def forward(self, q: torch.Tensor, k: torch.Tensor, offset: int = 0) -> tuple[torch.Tensor, torch.Tensor]:
        if q.shape[1] != k.shape[1]:
            raise ValueError("q and k must have the same sequence length")
        self.update_cos_sin(q.shape[1] + offset, device=q.device, dtype=q.dtype)
        return (apply_rotary_emb(q, self.cos[offset:], self.sin[offset:]),
                apply_rotary_emb(k, self.cos[offset:], self.sin[offset:]))
