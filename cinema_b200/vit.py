"""ViT building blocks with the reference's module API (cinema/vit.py) on top of the sm_100a kernels.

Same class names, constructor arguments, attribute names and ``state_dict`` keys as the reference
(``attn.q`` / ``attn.kv`` / ``attn.proj``, ``mlp.fc1`` / ``mlp.fc2``, ``norm1`` / ``norm2``, ``cls_token`` ...), so
checkpoints and the fine-tuning models that consume ``ViTEncoder`` drop in unchanged
(SURVEY.md section 8b).  The forward / backward math is NOT torch autograd over torch ops: each public
``forward`` is one ``torch.autograd.Function`` whose forward and backward are explicit chains of
C-ABI kernel launches (cinema_b200/engine.py).  Parameter gradients are accumulated by the
kernels directly into the flat gradient arena (``p.grad`` is a view of it), which is what the
single NCCL all-reduce and the optimiser consume.
"""

from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from cinema_b200 import _C, engine
from cinema_b200.arena import ensure_arena
from cinema_b200.conv import Linear
from cinema_b200.rotary import RotaryEmbedding

F32, BF16 = torch.float32, torch.bfloat16


# ------------------------------------------------------------------------------------------
# patchify / unpatchify (cinema/vit.py:67-256)
# ------------------------------------------------------------------------------------------
def _check_divisible(spatial: tuple[int, ...], patch_size: tuple[int, ...]) -> None:
    for s, p in zip(spatial, patch_size):
        if s % p != 0:
            raise ValueError(f"Input size ({tuple(spatial)}) cannot be divided by patch size ({tuple(patch_size)}).")


def patchify(image: torch.Tensor, patch_size: tuple[int, ...]) -> torch.Tensor:
    """(B, C, *spatial) -> (B, prod(grid), prod(patch) * C), channel fastest inside a token
    ("nchpwqdr->nhwdpqrc", cinema/vit.py:67-161).  Bit-exact copy kernel; 2-D, 3-D and 4-D."""
    n = len(patch_size)
    if n not in (2, 3, 4):
        raise ValueError(f"Patchify only supports 2D, 3D, and 4D images, got {n}D.")
    if image.dim() != n + 2:
        raise ValueError(f"image of shape {tuple(image.shape)} does not match a {n}-D patch size")
    b, c, *spatial = image.shape
    _check_divisible(tuple(spatial), tuple(patch_size))
    if image.dtype not in (F32, BF16, torch.float16):
        raise ValueError(f"patchify supports 2- and 4-byte floating dtypes, got {image.dtype}")
    image = image.contiguous()
    grid = [s // p for s, p in zip(spatial, patch_size)]
    out = torch.empty((b, math.prod(grid), math.prod(patch_size) * c), dtype=image.dtype, device=image.device)
    _C.patchify(image, out, b, c, spatial, patch_size, inverse=False)
    return out


def unpatchify(x: torch.Tensor, patch_size: tuple[int, ...], grid_size: tuple[int, ...]) -> torch.Tensor:
    """Inverse of :func:`patchify` (cinema/vit.py:164-256)."""
    b, n_patches, chans = x.shape
    if n_patches != math.prod(grid_size):
        raise ValueError(f"Number of patches {n_patches} != product of grid size {math.prod(grid_size)}.")
    if chans % math.prod(patch_size) != 0:
        raise ValueError(f"Number of channels {chans} is not divisible by product of patch size {patch_size}.")
    if len(patch_size) != len(grid_size):
        raise ValueError(f"Patch size {patch_size} and grid size {grid_size} do not match.")
    if len(patch_size) not in (2, 3, 4):
        raise ValueError(f"Unpatchify only supports 2D, 3D, and 4D images, got {len(patch_size)}D.")
    c = chans // math.prod(patch_size)
    spatial = [g * p for g, p in zip(grid_size, patch_size)]
    x = x.contiguous()
    out = torch.empty((b, c, *spatial), dtype=x.dtype, device=x.device)
    _C.patchify(x, out, b, c, spatial, patch_size, inverse=True)
    return out


def patchify_2d(image, patch_size):
    return patchify(image, patch_size)


patchify_3d = patchify_4d = patchify_2d


def unpatchify_2d(x, patch_size, grid_size):
    return unpatchify(x, patch_size, grid_size)


unpatchify_3d = unpatchify_4d = unpatchify_2d


# ------------------------------------------------------------------------------------------
# init helpers and fixed sin-cos positional embedding (cinema/vit.py:32-64, 347-443)
# ------------------------------------------------------------------------------------------
def init_weights(m: nn.Module) -> None:
    """xavier-uniform Linear weights with zero bias, unit LayerNorm (cinema/vit.py:32-48)."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        if m.bias is not None:
            nn.init.zeros_(m.bias)
        if m.weight is not None:
            nn.init.ones_(m.weight)


def get_tokens(embed_dim: int, n_tokens: int) -> nn.Parameter:
    """Learnable (1, n_tokens, embed_dim) token ~ N(0, 0.02) (cinema/vit.py:51-64)."""
    token = nn.Parameter(torch.zeros(1, n_tokens, embed_dim))
    nn.init.normal_(token, std=0.02)
    return token


def get_nd_sincos_pos_embed(embed_dim: int, grid_size: tuple[int, ...]) -> np.ndarray:
    """(prod(grid), embed_dim) table.  Reproduces the reference exactly, including its use of
    ``np.meshgrid``'s default "xy" indexing (axis 0 and 1 swapped, cinema/vit.py:421) and the zero
    padding when embed_dim is not divisible by 2 * ndim (cinema/vit.py:398-405)."""
    axes = [np.arange(s, dtype=np.float32) for s in grid_size]
    coords = np.stack(np.meshgrid(*axes), axis=0)
    n_axes = coords.shape[0]
    per_axis = embed_dim // n_axes
    per_axis -= per_axis % 2
    if per_axis <= 0:
        raise ValueError(f"Embedding dimension must be divisible by 2, got {embed_dim}.")
    half = per_axis // 2
    freq = np.exp(-np.log(10000) * np.arange(half, dtype=np.float32) / half)
    parts = []
    for a in range(n_axes):
        ang = np.einsum("m,d->md", coords[a].reshape(-1), freq)
        parts += [np.sin(ang), np.cos(ang)]
    table = np.concatenate(parts, axis=1)
    tail = embed_dim - per_axis * n_axes
    if tail > 0:
        table = np.concatenate([table, np.zeros((table.shape[0], tail))], axis=1)
    return table


def get_pos_embed(embed_dim: int, grid_size: tuple[int, ...]) -> nn.Parameter:
    """Fixed (1, N, E) positional embedding, a non-trainable Parameter so that it is part of the state dict."""
    table = get_nd_sincos_pos_embed(embed_dim, grid_size)
    p = nn.Parameter(torch.zeros(1, math.prod(grid_size), embed_dim), requires_grad=False)
    p.data.copy_(torch.from_numpy(table).float().unsqueeze(0))
    return p


class Mlp(nn.Module):
    """fc2(drop(norm(drop(act(fc1 x))))) -- the timm 1.0.15 ``Mlp`` the reference plugs in as ``mlp_layer``
    (cinema/vit.py:570-575, cinema/mae/mae.py:310).  On the B200 path fc1 + GELU and fc2 + residual are fused GEMMs."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                 bias=True, drop=0.0, use_conv=False) -> None:
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = tuple(bias) if isinstance(bias, (tuple, list)) else (bias, bias)
        drop = tuple(drop) if isinstance(drop, (tuple, list)) else (drop, drop)
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop[0])
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])
        self.drop2 = nn.Dropout(drop[1])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise RuntimeError("cinema_b200.vit.Mlp is executed as part of a fused Block (no standalone torch path)")


# ------------------------------------------------------------------------------------------
# the one autograd node behind Attention / Block / ViTEncoder / ViTDecoder
# ------------------------------------------------------------------------------------------
def _to_f32_rows(x: torch.Tensor) -> torch.Tensor:
    return x.detach().to(F32).contiguous()


def _to_bf16(x32: torch.Tensor) -> torch.Tensor:
    out = torch.empty(x32.shape, dtype=BF16, device=x32.device)
    _C.cast_bf16(x32, out)
    return out


class _StackFn(torch.autograd.Function):
    """[cls |] x -> blocks (self- or cross-attention) -> [last n rows] -> [LayerNorm]."""

    @staticmethod
    def forward(ctx, x, k, anchor, spec):  # noqa: ARG004 - anchor ties the node to the parameters
        blocks, cls_token, norm, n_last, owner = spec
        train = any(ctx.needs_input_grad[:3])  # grad mode is off inside forward(); this reflects the caller's
        arena = ensure_arena(owner)
        arena.refresh_shadow()
        if train:
            arena.prepare_grads()
        b, n_in, d = x.shape
        x32 = _to_f32_rows(x)
        if cls_token is not None:
            xin = torch.empty((b, n_in + 1, d), dtype=F32, device=x32.device)
            _C.embed_rows(None, 0, cls_token.data.view(-1), None, None, b, 1, out=xin, out_off=0)
            _C.embed_rows(x32, 0, None, None, None, b, n_in, out=xin, out_off=1)
            x32 = xin
        n = x32.shape[1]
        ws = [engine.blockw(arena, blk, train) for blk in blocks]
        k16 = None
        kvs = []
        if k is not None:
            k16 = _to_bf16(_to_f32_rows(k)).view(-1, d)
            nk = k.shape[1]
            for w in ws:
                kv = engine.linear_fwd(k16, w.kv).view(b, nk, 2, w.n_heads, d // w.n_heads)
                kvs.append(kv)
        cur = x32.view(b * n, d)
        saved = []
        for i, w in enumerate(ws):
            kv = (kvs[i][:, :, 0], kvs[i][:, :, 1]) if k is not None else None
            cur, sv = engine.block_fwd(cur, w, b, kv, train)
            saved.append(sv)
        n_out = n if n_last is None else n_last
        nw = engine.normw(arena, norm, train) if norm is not None else None
        if nw is not None:
            _, y32, mean, rstd = engine.ln_fwd(cur, nw, want16=False, want32=True, stats=train)
        else:
            y32, mean, rstd = cur, None, None
        out = y32.view(b, n, d)
        if n_last is not None and n_last != n:
            out = out[:, n - n_last:].contiguous()
        if train:
            ctx.state = (ws, nw, saved, kvs, k16, cur if nw is not None else None, mean, rstd, b, n, d, n_out,
                         cls_token is not None, arena, cls_token, x.dtype, k.dtype if k is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ws, nw, saved, kvs, k16, xl, mean, rstd, b, n, d, n_out, has_cls, arena, cls_token, xdt, kdt) = ctx.state
        ctx.state = None
        dev = dout.device
        dy = dout.detach().to(F32).contiguous()
        if n_out != n:
            full = torch.zeros((b, n, d), dtype=F32, device=dev)
            _C.scatter_rows(dy, engine.arange_idx(b, n - n_out, n_out, dev), full)
            dy = full
        dy = dy.view(b * n, d)
        if nw is not None:
            dx32, dx16 = engine.ln_bwd(dy, xl, mean, rstd, nw)
        else:
            dx32, dx16 = dy.clone() if dy.data_ptr() == dout.data_ptr() else dy, _to_bf16(dy)
        dk32 = None
        if k16 is not None:
            dk32 = torch.zeros((k16.shape[0], d), dtype=F32, device=dev)
        for i in range(len(ws) - 1, -1, -1):
            w = ws[i]
            if k16 is not None:
                dkv = torch.empty_like(kvs[i])
                dx32, dx16 = engine.block_bwd(dx32, dx16, w, b, saved[i], (kvs[i][:, :, 0], kvs[i][:, :, 1]),
                                              (dkv[:, :, 0], dkv[:, :, 1]))
                dkv2 = dkv.view(-1, 2 * d)
                if w.kv.gw is not None:
                    _C.gemm(dkv2, k16, w.kv.gw, a_mn=True, b_mn=True, accumulate=True)
                if w.kv.gb is not None:
                    _C.colsum(dkv2, w.kv.gb)
                _C.gemm(dkv2, w.kv.w16, dk32, b_mn=True, accumulate=True)
            else:
                dx32, dx16 = engine.block_bwd(dx32, dx16, w, b, saved[i], None, None)
            saved[i] = None
        dx32 = dx32.view(b, n, d)
        if has_cls:
            if cls_token.requires_grad:
                _C.colsum_seg(dx32, 0, 1, arena.grad_view(cls_token).view(-1))
            dx_in = dx32[:, 1:]
        else:
            dx_in = dx32
        dk = dk32.view(b, -1, d).to(kdt) if dk32 is not None else None
        return dx_in.to(xdt), dk, None, None


def _anchor(module: nn.Module) -> torch.Tensor:
    for p in module.parameters():
        if p.requires_grad:
            return p
    return next(module.parameters())


# ------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    """Image -> patch tokens -> Linear (cinema/vit.py:259-344)."""

    def __init__(self, image_size, patch_size, in_chans, embed_dim, norm_layer=None, bias=True,
                 strict_image_size=False, dynamic_img_pad=False) -> None:
        super().__init__()
        self.grad_ckpt = False
        self.n_dims = len(image_size)
        self.patch_size = patch_size
        self.image_size = image_size
        self.grid_size = tuple(s // p for s, p in zip(image_size, patch_size))
        self.n_patches = math.prod(self.grid_size)
        self.strict_image_size = strict_image_size
        self.dynamic_img_pad = dynamic_img_pad
        self.proj = Linear(in_chans * math.prod(patch_size), embed_dim, bias=bias)
        nn.init.xavier_uniform_(self.proj.weight.data.view(self.proj.weight.shape[0], -1))
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        self.proj.set_grad_ckpt(enable)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        spatial = tuple(image.shape[2:])
        if self.strict_image_size:
            if tuple(self.image_size) != spatial:
                raise ValueError(f"Input size ({image.shape}) doesn't match config (batch, channel) + {self.image_size}.")
        elif not self.dynamic_img_pad:
            for s, p in zip(spatial, self.patch_size):
                if s % p != 0:
                    raise ValueError(f"Input size ({spatial}) should be divisible by patch size ({self.patch_size}).")
        if self.dynamic_img_pad:
            pad: tuple[int, ...] = ()
            for s, p in zip(spatial, self.patch_size):
                pad = (0, (p - s % p) % p, *pad)
            image = torch.nn.functional.pad(image, pad)
        x = _PatchifyFn.apply(image, tuple(self.patch_size))
        return self.norm(self.proj(x))


class _PatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, patch_size):
        ctx.patch_size = patch_size
        ctx.grid = tuple(s // p for s, p in zip(image.shape[2:], patch_size))
        return patchify(image, patch_size)

    @staticmethod
    def backward(ctx, g):
        return unpatchify(g, ctx.patch_size, ctx.grid), None


class Attention(nn.Module):
    """Multi-head attention with separate q and kv projections and optional cross-attention
    (cinema/vit.py:446-522).  ``rotary=True`` is accepted for constructor compatibility; in the
    reference the rotation is applied over the *head* axis and cancels in q.k^T,
    so the observable output equals ``rotary=False`` and the fused kernel computes exactly that."""

    def __init__(self, dim, n_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0, proj_drop=0.0,
                 norm_layer=nn.LayerNorm, norm_eps=1e-5, rotary=False) -> None:
        super().__init__()
        if dim % n_heads != 0:
            raise ValueError(f"dim {dim} should be divisible by n_heads {n_heads}")
        self.n_heads = n_heads
        self.head_dim = dim // n_heads
        self.scale = self.head_dim ** -0.5
        self.fused_attn = True
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.q_norm = norm_layer(self.head_dim, eps=norm_eps) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim, eps=norm_eps) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.rotary = RotaryEmbedding(self.head_dim) if rotary else None

    def _arena_groups(self):
        if self.q.bias is None:
            return []
        return [[self.q.weight, self.kv.weight], [self.q.bias, self.kv.bias]]

    def forward(self, q: torch.Tensor, k: torch.Tensor | None = None) -> torch.Tensor:
        if k is not None and self.rotary is not None:
            raise ValueError("Rotary positional embedding is not supported with different query and key.")
        if not isinstance(self.q_norm, nn.Identity):
            raise NotImplementedError("the B200 attention path needs qk_norm=False (as in every CineMA config)")
        if self.training and (self.attn_drop.p > 0 or self.proj_drop.p > 0):
            raise NotImplementedError("attention / projection dropout is not part of the MAE hot path")
        return _AttentionFn.apply(q, k, _anchor(self), self)


class _AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, anchor, mod):  # noqa: ARG004
        train = any(ctx.needs_input_grad[:3])
        arena = ensure_arena(mod)
        arena.refresh_shadow()
        if train:
            arena.prepare_grads()
        b, n, d = x.shape
        h, hd = mod.n_heads, mod.head_dim
        wq = engine.linw(arena, mod.q.weight, mod.q.bias, train)
        wkv = engine.linw(arena, mod.kv.weight, mod.kv.bias, train)
        wp = engine.linw(arena, mod.proj.weight, mod.proj.bias, train)
        x16 = _to_bf16(_to_f32_rows(x)).view(b * n, d)
        k16 = _to_bf16(_to_f32_rows(k)).view(-1, d) if k is not None else x16
        nk = k.shape[1] if k is not None else n
        q2 = engine.linear_fwd(x16, wq)
        kv2 = engine.linear_fwd(k16, wkv)
        kv5 = kv2.view(b, nk, 2, h, hd)
        o, lse = engine.attn_fwd(q2.view(b, n, h, hd), kv5[:, :, 0], kv5[:, :, 1], mod.scale)
        y = engine.linear_fwd(o.view(b * n, d), wp, out_dtype=F32)
        if train:
            ctx.state = (wq, wkv, wp, x16, k16, q2, kv2, o, lse, b, n, nk, d, h, hd, mod.scale, k is not None, x.dtype,
                         k.dtype if k is not None else None)
        return y.view(b, n, d)

    @staticmethod
    def backward(ctx, dy):
        wq, wkv, wp, x16, k16, q2, kv2, o, lse, b, n, nk, d, h, hd, scale, cross, xdt, kdt = ctx.state
        ctx.state = None
        dy16 = _to_bf16(dy.detach().to(F32).contiguous()).view(b * n, d)
        do2 = engine.linear_bwd(dy16, o.view(b * n, d), wp)
        dq2 = torch.empty_like(q2)
        dkv2 = torch.empty_like(kv2)
        kv5, dkv5 = kv2.view(b, nk, 2, h, hd), dkv2.view(b, nk, 2, h, hd)
        engine.attn_bwd(q2.view(b, n, h, hd), kv5[:, :, 0], kv5[:, :, 1], o, do2.view(b, n, h, hd), lse,
                        dq2.view(b, n, h, hd), dkv5[:, :, 0], dkv5[:, :, 1], scale)
        dx = engine.linear_bwd(dq2, x16, wq, dx_dtype=F32)
        if cross:
            dk = engine.linear_bwd(dkv2, k16, wkv, dx_dtype=F32)
            return dx.view(b, n, d).to(xdt), dk.view(b, nk, d).to(kdt), None, None
        if wkv.gw is not None:
            _C.gemm(dkv2, k16, wkv.gw, a_mn=True, b_mn=True, accumulate=True)
        if wkv.gb is not None:
            _C.colsum(dkv2, wkv.gb)
        _C.gemm(dkv2, wkv.w16, dx, b_mn=True, accumulate=True)
        return dx.view(b, n, d).to(xdt), None, None, None


class Block(nn.Module):
    """Pre-LN transformer block (cinema/vit.py:525-609)."""

    def __init__(self, dim, n_heads, mlp_ratio, norm_layer, norm_eps, drop_path, qkv_bias, rotary, act_layer,
                 mlp_layer, qk_norm=False, proj_drop=0.0, attn_drop=0.0, init_values=None) -> None:
        super().__init__()
        if init_values:
            raise NotImplementedError("LayerScale is never enabled by the reference configs (cinema/vit.py:561)")
        if act_layer is not nn.GELU:
            raise NotImplementedError("the fused MLP epilogue implements exact-erf GELU (nn.GELU), as all CineMA configs use")
        self.grad_ckpt = False
        self.norm1 = norm_layer(dim, eps=norm_eps)
        self.attn = Attention(dim, n_heads=n_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, attn_drop=attn_drop,
                              proj_drop=proj_drop, norm_layer=norm_layer, norm_eps=norm_eps, rotary=rotary)
        self.ls1 = nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim, eps=norm_eps)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=proj_drop)
        self.ls2 = nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        # Kept for API compatibility (cinema/vit.py:579-585).  Activations are saved, never recomputed:
        # a ViT-B step at 16 frame-sets per GPU keeps < 10 GB of the 180 GB HBM3e.
        self.grad_ckpt = enable

    def forward(self, q: torch.Tensor, k: torch.Tensor | None = None) -> torch.Tensor:
        return _StackFn.apply(q, k, _anchor(self), ([self], None, None, None, self))


class DropPath(nn.Module):
    """Stochastic depth per sample (timm ``DropPath``, used at cinema/vit.py:562,577 when fine-tuning with drop_path > 0).
    Inside a ``Block`` it never runs as a module: the per-sample keep / keep_prob factors are drawn by
    ``engine.draw_drop_scales`` and applied in the epilogue of the branch's last GEMM.  Called on its own it is the plain
    elementwise definition."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True) -> None:
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        t = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            t.div_(keep)
        return x * t

    def extra_repr(self) -> str:
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


class ViTEncoder(nn.Module):
    """cls token + ``depth`` blocks + final LayerNorm (cinema/vit.py:612-698)."""

    def __init__(self, embed_dim, depth, n_heads, mlp_ratio, qkv_bias, norm_layer, norm_eps, rotary, act_layer,
                 mlp_layer, drop_path) -> None:
        super().__init__()
        self.grad_ckpt = False
        self.cls_token = get_tokens(embed_dim=embed_dim, n_tokens=1)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, n_heads=n_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer,
                  norm_eps=norm_eps, rotary=rotary, act_layer=act_layer, mlp_layer=mlp_layer, drop_path=drop_path)
            for _ in range(depth)
        ])
        self.norm = norm_layer(embed_dim, eps=norm_eps)
        self.apply(init_weights)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        for blk in self.blocks:
            blk.set_grad_ckpt(enable)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, n, D) tokens with positional embedding -> (B, 1 + n, D)."""
        return _StackFn.apply(x, None, _anchor(self), (list(self.blocks), self.cls_token, self.norm, None, self))

    def feature_forward(self, x: torch.Tensor) -> torch.Tensor:
        """All block outputs, the last one normalised: (B, 1 + n, D, depth) (cinema/vit.py:680-698).
        Inference-only on the B200 path (nothing in the reference back-propagates through it)."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            with torch.no_grad():
                return self.feature_forward(x)
        feats = []
        cur = _StackFn.apply(x, None, _anchor(self), ([self.blocks[0]], self.cls_token, None, None, self))
        for i, blk in enumerate(self.blocks):
            if i > 0:
                cur = _StackFn.apply(cur, None, _anchor(self), ([blk], None, None, None, self))
            if i != len(self.blocks) - 1:
                feats.append(cur)
        b, n, d = cur.shape
        nw = engine.normw(ensure_arena(self), self.norm, False)
        _, y32, _, _ = engine.ln_fwd(cur.reshape(b * n, d), nw, want16=False, want32=True, stats=False)
        feats.append(y32.view(b, n, d))
        return torch.stack(feats, dim=-1)


class ViTDecoder(nn.Module):
    """``depth`` blocks (self- or cross-attention) + LayerNorm over the masked tokens (cinema/vit.py:701-781)."""

    def __init__(self, embed_dim, depth, n_heads, mlp_ratio, qkv_bias, norm_layer, norm_eps, rotary, act_layer,
                 mlp_layer, drop_path) -> None:
        super().__init__()
        self.grad_ckpt = False
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, n_heads=n_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer,
                  norm_eps=norm_eps, rotary=rotary, act_layer=act_layer, mlp_layer=mlp_layer, drop_path=drop_path)
            for _ in range(depth)
        ])
        self.norm = norm_layer(embed_dim)
        self.apply(init_weights)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        for blk in self.blocks:
            blk.set_grad_ckpt(enable)

    def forward(self, x_q: torch.Tensor, x_k: torch.Tensor | None, n_enc_masked: int) -> torch.Tensor:
        return _StackFn.apply(x_q, x_k, _anchor(self), (list(self.blocks), None, self.norm, n_enc_masked, self))


def get_vit_config(size: str) -> dict[str, int]:
    """Size table of the reference (cinema/vit.py:784-831)."""
    table = {
        "tiny": (16, 1, 2, 16, 1, 2),
        "base": (768, 12, 12, 512, 8, 16),
        "large": (1024, 24, 16, 512, 8, 16),
        "huge": (1280, 32, 16, 512, 8, 16),
    }
    if size not in table:
        raise ValueError(f"size must be in ['tiny', 'base', 'large', 'huge'], got {size}.")
    keys = ("enc_embed_dim", "enc_depth", "enc_n_heads", "dec_embed_dim", "dec_depth", "dec_n_heads")
    return dict(zip(keys, table[size]))
