"""CPU oracle for the CineMA MAE-ViT hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *functional* restatement (plain fp32 PyTorch ops over a flat
``state_dict``) of the algorithm the reference implements with nn.Modules in
``cinema/vit.py``, ``cinema/rotary.py``, ``cinema/conv.py``, ``cinema/convvit.py``
and ``cinema/mae/mae.py``.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or as the
CPU baseline -- never on the product path.  ``cinema_b200`` must never import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference`` (with stub modules for the absent ``timm`` / ``omegaconf``)
and stores inputs, masks, weights, outputs and gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors and
against the reference's own known-answer tests (``cinema/rotary_test.py:9-13``,
``cinema/convvit_test.py:21-50``, ``cinema/mae/mae_test.py:30-32``).
Third-party arithmetic restated here: ``timm==1.0.15`` ``Mlp`` = fc2(GELU_erf(fc1 x))
(drop=0, norm=Identity); the timm-equivalence tests of the reference
(``cinema/vit_test.py:130-143``) cannot run offline, so that boundary is pinned only
through the shimmed reference run.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F  # noqa: N812

Tensor = torch.Tensor
StateDict = dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration (mirrors the CineMA constructor arguments, cinema/mae/mae.py:288-313)
# --------------------------------------------------------------------------------------
@dataclass
class MAEConfig:
    image_size_dict: dict[str, tuple[int, ...]]
    in_chans_dict: dict[str, int]
    enc_patch_size_dict: dict[str, tuple[int, ...]]
    enc_scale_factor_dict: dict[str, tuple[int, ...]]
    enc_conv_chans: list[int]
    enc_conv_n_blocks: int
    enc_embed_dim: int
    enc_depth: int
    enc_n_heads: int
    dec_embed_dim: int
    dec_depth: int
    dec_n_heads: int
    mlp_ratio: int = 4
    norm_target: bool = False
    cross_attn: bool = True
    norm_eps: float = 1e-5
    rotary: bool = False
    conv_norm_eps: float = 1e-6  # cinema/conv.py:190
    views: list[str] = field(default_factory=list)

    def __post_init__(self) -> None:
        if not self.views:
            self.views = list(self.image_size_dict.keys())

    # cinema/convvit.py:85
    def patch_sizes(self, view: str) -> list[tuple[int, ...]]:
        return [tuple(self.enc_patch_size_dict[view])] + [tuple(self.enc_scale_factor_dict[view])] * len(
            self.enc_conv_chans
        )

    # cinema/convvit.py:88-113 -- grid of the ViT tokens for the configured image size
    def grid_size(self, view: str) -> tuple[int, ...]:
        size = tuple(self.image_size_dict[view])
        for ps in self.patch_sizes(view):
            size = tuple(s // p for s, p in zip(size, ps))
        return size

    # cinema/mae/mae.py:207-228
    def dec_patch_size(self, view: str) -> tuple[int, ...]:
        out = (1,) * len(self.image_size_dict[view])
        for ps in self.patch_sizes(view):
            out = tuple(a * b for a, b in zip(out, ps))
        return out

    def n_patches(self, view: str) -> int:
        return math.prod(self.grid_size(view))


VIT_SIZES = {  # cinema/vit.py:798-831
    "tiny": dict(enc_embed_dim=16, enc_depth=1, enc_n_heads=2, dec_embed_dim=16, dec_depth=1, dec_n_heads=2),
    "base": dict(enc_embed_dim=768, enc_depth=12, enc_n_heads=12, dec_embed_dim=512, dec_depth=8, dec_n_heads=16),
    "large": dict(enc_embed_dim=1024, enc_depth=24, enc_n_heads=16, dec_embed_dim=512, dec_depth=8, dec_n_heads=16),
    "huge": dict(enc_embed_dim=1280, enc_depth=32, enc_n_heads=16, dec_embed_dim=512, dec_depth=8, dec_n_heads=16),
}


def make_config(
    size: str = "base",
    views: tuple[str, ...] = ("sax", "lax_2c", "lax_3c", "lax_4c"),
    sax_size: tuple[int, int, int] = (192, 192, 16),
    lax_size: tuple[int, int] = (256, 256),
    **overrides: int,
) -> MAEConfig:
    """Config of cinema/mae/config.yaml:13-55 routed through get_model (cinema/mae/mae.py:231-282)."""
    dims = dict(VIT_SIZES[size]) if size in VIT_SIZES else {}
    dims.update(overrides)
    is3d = {v: v == "sax" for v in views}
    return MAEConfig(
        image_size_dict={v: tuple(sax_size) if is3d[v] else tuple(lax_size) for v in views},
        in_chans_dict={v: 1 for v in views},
        enc_patch_size_dict={v: (4, 4, 1) if is3d[v] else (4, 4) for v in views},
        enc_scale_factor_dict={v: (2, 2, 1) if is3d[v] else (2, 2) for v in views},
        enc_conv_chans=[64, 128],
        enc_conv_n_blocks=2,
        **dims,
    )


# --------------------------------------------------------------------------------------
# patchify / unpatchify, any rank (cinema/vit.py:67-256)
# --------------------------------------------------------------------------------------
def patchify(image: Tensor, patch_size: tuple[int, ...]) -> Tensor:
    """(B, C, S1..Sn) -> (B, prod(grid), prod(patch)*C); channel is the fastest axis of a token.

    Restates the einsum "nchpwqdr->nhwdpqrc" family (cinema/vit.py:86,112,140) as one
    generic permute: split every spatial axis into (grid, patch), move all grid axes
    first, then all patch axes, then channels.
    """
    n = len(patch_size)
    if n not in (2, 3, 4):
        raise ValueError(f"Patchify only supports 2D, 3D, and 4D images, got {n}D.")
    b, c, *spatial = image.shape
    if len(spatial) != n:
        raise ValueError(f"image rank {len(spatial)} does not match patch rank {n}")
    for s, p in zip(spatial, patch_size):
        if s % p != 0:
            raise ValueError(f"Input size ({s}) cannot be divided by patch size ({p}).")
    grid = [s // p for s, p in zip(spatial, patch_size)]
    split = []
    for g, p in zip(grid, patch_size):
        split += [g, p]
    x = image.reshape(b, c, *split)  # (B, C, g1, p1, g2, p2, ...)
    grid_axes = [2 + 2 * i for i in range(n)]
    patch_axes = [3 + 2 * i for i in range(n)]
    x = x.permute(0, *grid_axes, *patch_axes, 1).contiguous()
    return x.reshape(b, math.prod(grid), math.prod(patch_size) * c)


def unpatchify(x: Tensor, patch_size: tuple[int, ...], grid_size: tuple[int, ...]) -> Tensor:
    """Inverse of :func:`patchify` (cinema/vit.py:164-256)."""
    b, n_patches, chans = x.shape
    if n_patches != math.prod(grid_size):
        raise ValueError(f"Number of patches {n_patches} != product of grid size {math.prod(grid_size)}.")
    if chans % math.prod(patch_size) != 0:
        raise ValueError(f"Number of channels {chans} is not divisible by product of patch size.")
    if len(patch_size) != len(grid_size):
        raise ValueError(f"Patch size {patch_size} and grid size {grid_size} do not match.")
    n = len(patch_size)
    if n not in (2, 3, 4):
        raise ValueError(f"Unpatchify only supports 2D, 3D, and 4D images, got {n}D.")
    c = chans // math.prod(patch_size)
    x = x.reshape(b, *grid_size, *patch_size, c)
    # axes now: 0 | 1..n (grid) | n+1..2n (patch) | 2n+1 (chan)
    order = [0, 2 * n + 1]
    for i in range(n):
        order += [1 + i, 1 + n + i]
    x = x.permute(*order).contiguous()
    return x.reshape(b, c, *[g * p for g, p in zip(grid_size, patch_size)])


# --------------------------------------------------------------------------------------
# fixed N-D sin-cos positional embedding (cinema/vit.py:347-443), float64 math then .float()
# --------------------------------------------------------------------------------------
def sincos_pos_embed(embed_dim: int, grid_size: tuple[int, ...]) -> Tensor:
    """(1, prod(grid), embed_dim).  Keeps the reference's ``np.meshgrid`` default
    ``indexing="xy"`` (cinema/vit.py:421), which swaps the first two axes, and the
    zero padding when embed_dim is not divisible by 2*ndim (cinema/vit.py:398-405)."""
    axes = [np.arange(s, dtype=np.float32) for s in grid_size]
    grid = np.stack(np.meshgrid(*axes), axis=0)  # xy indexing on purpose
    n = grid.shape[0]
    d = embed_dim // n
    d -= d % 2
    half = d // 2
    omega = np.arange(half, dtype=np.float32)
    omega = np.exp(-np.log(10000) * omega / half)  # cinema/vit.py:373-374
    chunks = []
    for i in range(n):
        ang = np.einsum("m,d->md", grid[i].reshape(-1), omega)
        chunks.append(np.concatenate([np.sin(ang), np.cos(ang)], axis=1))
    emb = np.concatenate(chunks, axis=1)
    pad = embed_dim - d * n
    if pad > 0:
        emb = np.concatenate([emb, np.zeros((emb.shape[0], pad))], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)


# --------------------------------------------------------------------------------------
# rotary (cinema/rotary.py) -- standalone contract (B, N, H, d)
# --------------------------------------------------------------------------------------
def rotate_half(x: Tensor) -> Tensor:
    """cinema/rotary.py:12-22."""
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def rotary_tables(n_tokens: int, dim: int, dtype: torch.dtype, base: float = 10000.0, scaling: float = 1.0):
    """cos/sin of shape (n_tokens, dim/2) in ``dtype`` (cinema/rotary.py:81,103-106)."""
    inv_freq = 1 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    t = torch.arange(n_tokens, dtype=torch.float32) / scaling
    freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs).to(dtype), torch.sin(freqs).to(dtype)


def apply_rotary(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """x: (B, N, H, d); cos/sin: (>=N, ro_dim/2).  cinema/rotary.py:25-50."""
    ro = cos.shape[-1] * 2
    if ro > x.shape[-1]:
        raise ValueError(f"Rotary dim {ro} is larger than the last dimension of x {x.shape[-1]}")
    n = x.shape[1]
    c = torch.cat([cos[:n], cos[:n]], dim=-1)[:, None, :].to(x.device)  # "s d -> s 1 (2 d)"
    s = torch.cat([sin[:n], sin[:n]], dim=-1)[:, None, :].to(x.device)
    head = x[..., :ro]
    return torch.cat([head * c + rotate_half(head) * s, x[..., ro:]], dim=-1)


def rotary_qk(q: Tensor, k: Tensor, dim: int, offset: int = 0) -> tuple[Tensor, Tensor]:
    """RotaryEmbedding.forward (cinema/rotary.py:108-128); sequence length is taken from axis 1."""
    if q.shape[1] != k.shape[1]:
        raise ValueError("q and k must have the same sequence length")
    cos, sin = rotary_tables(q.shape[1] + offset, dim, q.dtype)
    return apply_rotary(q, cos[offset:], sin[offset:]), apply_rotary(k, cos[offset:], sin[offset:])


# --------------------------------------------------------------------------------------
# transformer pieces (cinema/vit.py:446-781)
# --------------------------------------------------------------------------------------
def _lin(sd: StateDict, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def _ln(sd: StateDict, name: str, x: Tensor, eps: float) -> Tensor:
    w = sd[f"{name}.weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[f"{name}.bias"], eps)


def attention(sd: StateDict, name: str, q_in: Tensor, k_in: Tensor | None, n_heads: int, rotary: bool = False):
    """Attention.forward (cinema/vit.py:482-522): separate q and kv projections, kv rows are
    [k | v] (reshape(B,N,2,H,d), cinema/vit.py:499), softmax(q k^T / sqrt(d)) v, output proj.

    ``rotary=True`` reproduces the reference call exactly: RotaryEmbedding receives (B,H,N,d)
    although its contract is (B,N,H,d) (cinema/vit.py:498-503 vs cinema/rotary.py:112-113), so
    the rotation angle is indexed by the *head* and cancels in q.k^T."""
    if k_in is None:
        k_in = q_in
    elif rotary:
        raise ValueError("Rotary positional embedding is not supported with different query and key.")
    b, nq, ch = q_in.shape
    nk = k_in.shape[1]
    d = ch // n_heads
    q = _lin(sd, f"{name}.q", q_in).reshape(b, nq, n_heads, d).permute(0, 2, 1, 3)
    kv = _lin(sd, f"{name}.kv", k_in).reshape(b, nk, 2, n_heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    if rotary:
        q, k = rotary_qk(q, k, d)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(b, nq, ch)
    return _lin(sd, f"{name}.proj", o)


def mlp(sd: StateDict, name: str, x: Tensor) -> Tensor:
    """timm 1.0.15 Mlp with act=GELU(erf), drop=0, norm=Identity (used at cinema/vit.py:570-575)."""
    return _lin(sd, f"{name}.fc2", F.gelu(_lin(sd, f"{name}.fc1", x)))


def block(sd: StateDict, name: str, q: Tensor, k: Tensor | None, n_heads: int, eps: float, rotary: bool = False,
          drop: tuple[Tensor, Tensor] | None = None):
    """Block.forward (cinema/vit.py:587-609).  Only q is layer-normed; k is used as given.  ``drop``: the per-sample
    factors (B,) of drop_path1 / drop_path2 in training mode (timm DropPath: Bernoulli(keep) / keep), None = identity."""
    h = attention(sd, f"{name}.attn", _ln(sd, f"{name}.norm1", q, eps), k, n_heads, rotary)
    q = q + (h if drop is None else h * drop[0].to(h.dtype).view(-1, 1, 1))
    h = mlp(sd, f"{name}.mlp", _ln(sd, f"{name}.norm2", q, eps))
    q = q + (h if drop is None else h * drop[1].to(h.dtype).view(-1, 1, 1))
    return q


def vit_encoder(sd: StateDict, name: str, x: Tensor, depth: int, n_heads: int, eps: float, rotary: bool = False,
                return_all: bool = False, drop_scales: list[Tensor] | None = None):
    """ViTEncoder.forward / feature_forward (cinema/vit.py:661-698).  ``drop_scales``: 2 * depth stochastic-depth
    factor vectors in call order (see :func:`block`)."""
    cls = sd[f"{name}.cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1)
    feats = []
    for i in range(depth):
        x = block(sd, f"{name}.blocks.{i}", x, None, n_heads, eps, rotary,
                  None if drop_scales is None else (drop_scales[2 * i], drop_scales[2 * i + 1]))
        if i != depth - 1:
            feats.append(x)
    x = _ln(sd, f"{name}.norm", x, eps)
    if return_all:
        return torch.stack([*feats, x], dim=-1)
    return x


def vit_decoder(sd: StateDict, name: str, x_q: Tensor, x_k: Tensor | None, n_masked: int, depth: int, n_heads: int,
                eps: float):
    """ViTDecoder.forward (cinema/vit.py:749-781); final LayerNorm uses the default eps (cinema/vit.py:738)."""
    for i in range(depth):
        x_q = block(sd, f"{name}.blocks.{i}", x_q, x_k, n_heads, eps)
    x_q = x_q[:, x_q.shape[1] - n_masked:, :]
    return _ln(sd, f"{name}.norm", x_q, 1e-5)


# --------------------------------------------------------------------------------------
# conv stem (cinema/conv.py, cinema/convvit.py:24-291)
# --------------------------------------------------------------------------------------
def upsample_mask(mask: Tensor, scale_factor: tuple[int, ...]) -> Tensor:
    """Nearest-neighbour integer upsampling of a (B, *spatial) mask (cinema/convvit.py:24-51)."""
    if mask.ndim != len(scale_factor) + 1:
        raise ValueError(
            f"mask must have the same number of dimensions as scale_factor except batch, "
            f"got {mask.ndim} and {len(scale_factor)}."
        )
    for axis, f in enumerate(scale_factor):
        mask = mask.repeat_interleave(f, dim=axis + 1)
    return mask


def _conv(sd: StateDict, name: str, x: Tensor, stride=1, padding=0, groups: int = 1) -> Tensor:
    fn = F.conv2d if x.ndim == 4 else F.conv3d
    return fn(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"), stride=stride, padding=padding, groups=groups)


def conv_layer_norm(sd: StateDict, name: str, x: Tensor, eps: float) -> Tensor:
    """ConvLayerNorm: LayerNorm over the channel axis of a channel-first tensor (cinema/conv.py:169-187)."""
    x = x.movedim(1, -1)
    x = _ln(sd, name, x, eps)
    return x.movedim(-1, 1).contiguous()


def conv_norm_act(sd: StateDict, name: str, x: Tensor, stride: tuple[int, ...], eps: float) -> Tensor:
    """ConvNormActBlock with kernel == stride, "valid" padding, layer norm, GELU (cinema/conv.py:257-273)."""
    x = _conv(sd, f"{name}.conv", x, stride=stride)
    return F.gelu(conv_layer_norm(sd, f"{name}.norm", x, eps))


def masked_conv_block(sd: StateDict, name: str, x: Tensor, vis: Tensor | None, eps: float) -> Tensor:
    """MaskedConvBlock.forward (cinema/conv.py:400-415).  ``vis`` is 1 where visible."""
    h = _conv(sd, f"{name}.conv1", conv_layer_norm(sd, f"{name}.norm1", x, eps))
    if vis is not None:
        h = vis.unsqueeze(1).to(h.dtype) * h
    h = _conv(sd, f"{name}.dw_conv", h, padding=2, groups=h.shape[1])  # k=5, "same"
    x = x + _conv(sd, f"{name}.conv2", h)
    h = conv_layer_norm(sd, f"{name}.norm2", x, eps)
    h = _conv(sd, f"{name}.mlp.fc2", F.gelu(_conv(sd, f"{name}.mlp.fc1", h)))
    return x + h


def downsample_encoder(sd: StateDict, name: str, cfg: MAEConfig, view: str, image: Tensor, mask: Tensor | None):
    """DownsampleEncoder.forward (cinema/convvit.py:165-207).  Returns (skips, tokens(B, n_patches, D)); an input
    whose size differs from the configured image size gets the resampled positional table (cinema/convvit.py:139-163)."""
    b = image.shape[0]
    sizes = cfg.patch_sizes(view)
    grid = tuple(s // p for s, p in zip(image.shape[2:], cfg.dec_patch_size(view)))
    n_levels = len(cfg.enc_conv_chans)
    vis_masks: list[Tensor | None] = [None] * n_levels
    if mask is not None:
        m = mask.reshape(b, *grid)
        for lvl, ps in zip(range(n_levels - 1, -1, -1), sizes[:0:-1]):
            m = upsample_mask(m, ps)
            vis_masks[lvl] = ~m
    skips = []
    x = image
    for lvl in range(n_levels):
        x = conv_norm_act(sd, f"{name}.conv_blocks.{lvl}.patch_embed", x, sizes[lvl], cfg.conv_norm_eps)
        for j in range(cfg.enc_conv_n_blocks):
            x = masked_conv_block(sd, f"{name}.conv_blocks.{lvl}.conv.{j}", x, vis_masks[lvl], cfg.conv_norm_eps)
        skips.append(x)
    tok = _lin(sd, f"{name}.patch_embed.proj", patchify(x, sizes[-1]))
    tok = _lin(sd, f"{name}.linear", tok) + interpolate_pos_embed(sd[f"{name}.pos_embed"], cfg.grid_size(view), grid)
    return skips, tok


def interpolate_pos_embed(pos: Tensor, grid_from: tuple[int, ...], grid_to: tuple[int, ...]) -> Tensor:
    """DownsampleEncoder.interpolate_pos_encoding (cinema/convvit.py:139-163): bicubic (2-D) / trilinear (3-D)
    resampling of the (1, n, D) table, identity when the grids agree."""
    if tuple(grid_from) == tuple(grid_to):
        return pos
    d = pos.shape[-1]
    t = pos.float().reshape(1, *grid_from, d).movedim(-1, 1)
    t = F.interpolate(t, size=tuple(grid_to), mode={2: "bicubic", 3: "trilinear"}[len(grid_to)], antialias=False)
    return t.movedim(1, -1).reshape(1, -1, d).to(pos.dtype)


def _gather_rows(x: Tensor, sel: Tensor) -> Tensor:
    """x[sel].reshape(B, -1, C) for a boolean (B, n) selector with the same count in every row."""
    b, _, c = x.shape
    return x[sel].reshape(b, -1, c)


def multi_scale_fusion(sd: StateDict, name: str, cfg: MAEConfig, view: str, skips: list[Tensor], x: Tensor,
                       mask: Tensor | None) -> Tensor:
    """MultiScaleFusion.forward (cinema/convvit.py:265-291)."""
    sizes = cfg.patch_sizes(view)
    for lvl, skip in enumerate(skips):
        k = tuple(math.prod(ps[a] for ps in sizes[lvl + 1:]) for a in range(len(sizes[0])))
        down = _conv(sd, f"{name}.down_convs.{lvl}", skip, stride=k)
        down = down.flatten(2).transpose(1, 2)
        if mask is not None:
            down = _gather_rows(down, ~mask)
        x = x + down
    return _ln(sd, f"{name}.norm", x, cfg.norm_eps)


# --------------------------------------------------------------------------------------
# MAE pieces (cinema/mae/mae.py)
# --------------------------------------------------------------------------------------
def random_patch_mask(batch: int, n_patches: int, mask_ratio: float, device="cpu",
                      generator: torch.Generator | None = None) -> Tensor:
    """get_batch_random_patch_mask (cinema/mae/mae.py:30-65): 1 = removed, exactly int(n(1-r)) zeros per row."""
    if mask_ratio < 0:
        raise ValueError(f"mask_ratio must be positive, got {mask_ratio}.")
    if mask_ratio == 0:
        return torch.zeros((batch, n_patches), dtype=torch.bool, device=device)
    noise = torch.rand(batch, n_patches, device=device, generator=generator)
    rank = torch.argsort(torch.argsort(noise, dim=1), dim=1)
    return rank >= int(n_patches * (1 - mask_ratio))


def masked_mse(target: Tensor, pred: Tensor, mask: Tensor, norm_target: bool, eps: float = 1e-6):
    """mse_loss (cinema/mae/mae.py:107-152): per-patch mean / unbiased std metrics, optional target
    normalisation, mean squared error over masked patches only."""
    mean = target.mean(dim=-1, keepdim=True)
    std = target.var(dim=-1, keepdim=True) ** 0.5
    metrics = {"target_mean": mean.mean(), "target_std": std.mean()}
    if norm_target:
        target = (target - mean) / (std + eps)
    tgt = target[mask].reshape(pred.shape)
    loss = ((pred.float() - tgt.detach()) ** 2).mean()
    metrics["mse_loss"] = loss
    if norm_target and tgt.shape[1] > 0:
        metrics["normed_target_max"] = tgt.max()
        metrics["pred_max"] = pred.max()
    return loss, metrics


def mae_forward(sd: StateDict, cfg: MAEConfig, image_dict: dict[str, Tensor], enc_masks: dict[str, Tensor],
                taps: dict[str, Tensor] | None = None):
    """CineMA.forward (cinema/mae/mae.py:504-612) with the masks supplied by the caller.

    Returns (loss, pred_dict, metrics).  ``taps`` (optional dict) receives named intermediates."""
    views = list(image_dict.keys())
    if any(v not in cfg.views for v in views):
        raise ValueError(f"views {views} must be in {cfg.views}.")
    b = image_dict[views[0]].shape[0]
    tap = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)

    xs, skips_all, n_keep, n_masked = [], [], [], []
    for v in views:
        m = enc_masks[v]
        skips, tok = downsample_encoder(sd, f"enc_down_dict.{v}", cfg, v, image_dict[v], m)
        tap(f"tokens_all.{v}", tok)
        tok = _gather_rows(tok, ~m)
        skips_all.append(skips)
        xs.append(tok)
        n_keep.append(tok.shape[1])
        n_masked.append(m.shape[1] - tok.shape[1])
    tap("enc_in", torch.cat(xs, dim=1))

    x = vit_encoder(sd, "encoder", torch.cat(xs, dim=1), cfg.enc_depth, cfg.enc_n_heads, cfg.norm_eps, cfg.rotary)
    tap("enc_out", x)
    parts = list(torch.split(x, [1, *n_keep], dim=1))
    for i, v in enumerate(views):
        parts[i + 1] = multi_scale_fusion(sd, f"enc_fusion_dict.{v}", cfg, v, skips_all[i], parts[i + 1], enc_masks[v])
    tap("fused", torch.cat(parts, dim=1))

    x = _lin(sd, "dec_linear", torch.cat(parts, dim=1))
    parts = torch.split(x, [1, *n_keep], dim=1)
    vis, msk = [], []
    for i, v in enumerate(views):
        pe = sd[f"dec_embed_dict.{v}.pos_embed"].expand(b, -1, -1)  # cinema/mae/mae.py:97
        m = enc_masks[v]
        vis.append(parts[i + 1] + _gather_rows(pe, ~m))
        msk.append(sd[f"dec_embed_dict.{v}.mask_token"] + _gather_rows(pe, m))
    if cfg.cross_attn:
        x_q = torch.cat([parts[0], *msk], dim=1)
        x_k = torch.cat(vis, dim=1)
    else:
        x_q = torch.cat([parts[0], *vis, *msk], dim=1)
        x_k = None
    tap("dec_q", x_q)
    x = vit_decoder(sd, "decoder", x_q, x_k, sum(n_masked), cfg.dec_depth, cfg.dec_n_heads, cfg.norm_eps)
    tap("dec_out", x)
    outs = torch.split(x, n_masked, dim=1)

    preds, losses, metrics = {}, [], {}
    for i, v in enumerate(views):
        pred = _lin(sd, f"pred_head_dict.{v}", outs[i])
        preds[v] = pred
        target = patchify(image_dict[v], cfg.dec_patch_size(v))
        loss_v, met_v = masked_mse(target, pred, enc_masks[v], cfg.norm_target)
        metrics.update({f"{v}_{k}": val for k, val in met_v.items()})
        if torch.isfinite(loss_v):
            losses.append(loss_v)
    loss = sum(losses) / len(losses) if losses else torch.tensor(float("nan"))
    metrics["loss"] = loss
    return loss, preds, metrics


def mae_feature_forward(sd: StateDict, cfg: MAEConfig, image_dict: dict[str, Tensor]) -> dict[str, Tensor]:
    """CineMA.feature_forward (cinema/mae/mae.py:457-502): no masking, all tokens."""
    views = list(image_dict.keys())
    xs, skips_all, n_keep = [], [], []
    for v in views:
        skips, tok = downsample_encoder(sd, f"enc_down_dict.{v}", cfg, v, image_dict[v], None)
        skips_all.append(skips)
        xs.append(tok)
        n_keep.append(tok.shape[1])
    x = vit_encoder(sd, "encoder", torch.cat(xs, dim=1), cfg.enc_depth, cfg.enc_n_heads, cfg.norm_eps, cfg.rotary)
    parts = list(torch.split(x, [1, *n_keep], dim=1))
    for i, v in enumerate(views):
        parts[i + 1] = multi_scale_fusion(sd, f"enc_fusion_dict.{v}", cfg, v, skips_all[i], parts[i + 1], None)
    return dict(zip(["cls", *views], parts))


# --------------------------------------------------------------------------------------
# ConvViT: classification / regression fine-tuning model (cinema/convvit.py:334-561)
# --------------------------------------------------------------------------------------
def convvit_config(kw: dict) -> MAEConfig:
    """ConvViT constructor arguments -> the config object the stem / encoder restatements read (decoder fields unused).
    The stem sees ``n_frames * in_chans`` channels (cinema/convvit.py:400)."""
    return MAEConfig(
        image_size_dict=kw["image_size_dict"], in_chans_dict={v: kw["n_frames"] * c for v, c in kw["in_chans_dict"].items()},
        enc_patch_size_dict=kw["enc_patch_size_dict"], enc_scale_factor_dict=kw["enc_scale_factor_dict"],
        enc_conv_chans=kw["enc_conv_chans"], enc_conv_n_blocks=kw["enc_conv_n_blocks"], enc_embed_dim=kw["enc_embed_dim"],
        enc_depth=kw["enc_depth"], enc_n_heads=kw["enc_n_heads"], dec_embed_dim=0, dec_depth=0, dec_n_heads=1,
        mlp_ratio=kw.get("mlp_ratio", 4), norm_eps=kw.get("norm_eps", 1e-5), rotary=kw.get("rotary", False),
    )


def convvit_feature_forward(sd: StateDict, cfg: MAEConfig, image_dict: dict[str, Tensor],
                            mask_dict: dict[str, Tensor] | None, drop_scales: list[Tensor] | None = None) -> dict[str, Tensor]:
    """ConvViT.feature_forward (cinema/convvit.py:464-510): the mask only reaches the stem; every token is encoded and
    the fusion runs unmasked."""
    views = list(image_dict.keys())
    xs, skips_all, ns = [], [], []
    for v in views:
        skips, tok = downsample_encoder(sd, f"enc_down_dict.{v}", cfg, v, image_dict[v],
                                        None if mask_dict is None else mask_dict[v])
        skips_all.append(skips), xs.append(tok), ns.append(tok.shape[1])
    x = vit_encoder(sd, "encoder", torch.cat(xs, dim=1), cfg.enc_depth, cfg.enc_n_heads, cfg.norm_eps, cfg.rotary,
                    drop_scales=drop_scales)
    parts = list(torch.split(x, [1, *ns], dim=1))
    for i, v in enumerate(views):
        parts[i + 1] = multi_scale_fusion(sd, f"enc_fusion_dict.{v}", cfg, v, skips_all[i], parts[i + 1], None)
    return dict(zip(["cls", *views], parts))


def convvit_forward(sd: StateDict, cfg: MAEConfig, image_dict: dict[str, Tensor], mask_dict: dict[str, Tensor] | None,
                    reduce: str = "all", drop_scales: list[Tensor] | None = None) -> Tensor:
    """ConvViT.forward (cinema/convvit.py:512-561): per-view heads on the patch means (+ cls head), averaged."""
    x = convvit_feature_forward(sd, cfg, image_dict, mask_dict, drop_scales)
    if reduce == "cls":
        return _lin(sd, "pred_head_dict.cls", x["cls"])[:, 0]
    if reduce not in ("patch", "all"):
        raise NotImplementedError(f"Unsupported reduce method {reduce}.")
    logits = [_lin(sd, f"pred_head_dict.{v}", x[v].mean(dim=1, keepdim=True)) for v in cfg.views]
    if reduce == "all":
        logits.append(_lin(sd, "pred_head_dict.cls", x["cls"]))
    return torch.concat(logits, dim=1).mean(dim=1)


# --------------------------------------------------------------------------------------
# ConvUNetR: segmentation fine-tuning model (cinema/segmentation/convunetr.py:25-106,214-485)
# --------------------------------------------------------------------------------------
def conv_res_block(sd: StateDict, name: str, x: Tensor, eps: float) -> Tensor:
    """ConvResBlock.forward (cinema/conv.py:329-346), layer norm, GELU, dropout 0, odd kernel with "same" padding."""
    k = sd[f"{name}.conv1.weight"].shape[2:]
    pad = tuple(kk // 2 for kk in k)
    h = _conv(sd, f"{name}.conv1", F.gelu(conv_layer_norm(sd, f"{name}.norm1", x, eps)), padding=pad)
    h = _conv(sd, f"{name}.conv2", F.gelu(conv_layer_norm(sd, f"{name}.norm2", h, eps)), padding=pad)
    return h + (_conv(sd, f"{name}.shortcut", x) if f"{name}.shortcut.weight" in sd else x)


def _deconv(sd: StateDict, name: str, x: Tensor) -> Tensor:
    w = sd[f"{name}.weight"]  # kernel == stride (cinema/segmentation/convunetr.py:66)
    fn = F.conv_transpose2d if x.ndim == 4 else F.conv_transpose3d
    return fn(x, w, sd.get(f"{name}.bias"), stride=tuple(w.shape[2:]))


def upsample_decoder(sd: StateDict, name: str, embeddings: list[Tensor | None], n_blocks: int, eps: float) -> Tensor:
    """UpsampleDecoder.forward (cinema/segmentation/convunetr.py:88-106)."""
    embeddings = list(embeddings)
    x = embeddings.pop()
    level = 0
    while f"{name}.blocks.{level}.up.weight" in sd:
        x = _deconv(sd, f"{name}.blocks.{level}.up", x)
        skip = embeddings.pop()
        if skip is not None:
            x = x + skip
        for j in range(n_blocks):
            x = conv_res_block(sd, f"{name}.blocks.{level}.conv.{j}", x, eps)
        level += 1
    return x


def convunetr_forward(sd: StateDict, cfg: MAEConfig, image_dict: dict[str, Tensor], n_layers_wo_skip: int,
                      drop_scales: list[Tensor] | None = None) -> dict[str, Tensor]:
    """ConvUNetR.forward (cinema/segmentation/convunetr.py:436-485).  ``cfg`` from :func:`convvit_config`-style arguments
    (stem / encoder fields); the decoder structure is read off the state dict."""
    views = list(image_dict.keys())
    eps = cfg.conv_norm_eps
    xs, skips_all, ns = [], [], []
    for v in views:
        skips, tok = downsample_encoder(sd, f"enc_down_dict.{v}", cfg, v, image_dict[v], None)
        skips_all.append(skips), xs.append(tok), ns.append(tok.shape[1])
    x = vit_encoder(sd, "encoder", torch.cat(xs, dim=1), cfg.enc_depth, cfg.enc_n_heads, cfg.norm_eps, cfg.rotary,
                    drop_scales=drop_scales)
    tokens = torch.split(x, [1, *ns], dim=1)[1:]
    preds = {}
    for i, v in enumerate(views):
        grid = tuple(s // p for s, p in zip(image_dict[v].shape[2:], cfg.dec_patch_size(v)))
        xv = tokens[i].permute(0, 2, 1).reshape(x.shape[0], x.shape[2], *grid)
        levels = [*skips_all[i], xv]
        j = 0
        while f"dec_down_blocks_dict.{v}.{j}.weight" in sd:
            w = sd[f"dec_down_blocks_dict.{v}.{j}.weight"]
            xv = _conv(sd, f"dec_down_blocks_dict.{v}.{j}", xv, stride=tuple(w.shape[2:]))
            levels.append(xv)
            j += 1
        emb: list[Tensor | None] = [conv_res_block(sd, f"dec_image_conv_block_dict.{v}", image_dict[v], eps)]
        emb += [None] * n_layers_wo_skip
        emb += [conv_res_block(sd, f"dec_conv_blocks_dict.{v}.{k}", lv, eps) for k, lv in enumerate(levels)]
        preds[v] = _conv(sd, f"pred_head_dict.{v}", upsample_decoder(sd, f"decoder_dict.{v}", emb, 2, eps))
    return preds


def convunetr_config(kw: dict) -> MAEConfig:
    return convvit_config({**kw, "n_frames": 1})


# --------------------------------------------------------------------------------------
# random state dict with the reference's key schema / shapes / init distributions
# (cinema/vit.py:32-64, cinema/mae/mae.py:349-442).  Used for CPU-baseline timing and for
# tests that do not need reference-identical random streams.
# --------------------------------------------------------------------------------------
def init_state_dict(cfg: MAEConfig, seed: int = 0) -> StateDict:
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}

    def xavier(*shape: int) -> Tensor:
        fan_out = shape[0] * math.prod(shape[2:])
        fan_in = shape[1] * math.prod(shape[2:])
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    def linear(name: str, fin: int, fout: int) -> None:
        sd[f"{name}.weight"] = xavier(fout, fin)
        sd[f"{name}.bias"] = torch.zeros(fout)

    def norm(name: str, dim: int) -> None:
        sd[f"{name}.weight"] = torch.ones(dim)
        sd[f"{name}.bias"] = torch.zeros(dim)

    def conv(name: str, cin: int, cout: int, k: tuple[int, ...], groups: int = 1) -> None:
        fan_in = cin // groups * math.prod(k)
        bound = 1 / math.sqrt(fan_in)  # torch default conv init (kaiming_uniform a=sqrt(5))
        sd[f"{name}.weight"] = (torch.rand(cout, cin // groups, *k, generator=g) * 2 - 1) * bound
        sd[f"{name}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def vit_block(name: str, dim: int) -> None:
        norm(f"{name}.norm1", dim)
        linear(f"{name}.attn.q", dim, dim)
        linear(f"{name}.attn.kv", dim, 2 * dim)
        linear(f"{name}.attn.proj", dim, dim)
        norm(f"{name}.norm2", dim)
        linear(f"{name}.mlp.fc1", dim, dim * cfg.mlp_ratio)
        linear(f"{name}.mlp.fc2", dim * cfg.mlp_ratio, dim)

    for v in cfg.views:
        nd = len(cfg.image_size_dict[v])
        sizes = cfg.patch_sizes(v)
        cin = cfg.in_chans_dict[v]
        base = f"enc_down_dict.{v}"
        for lvl, ch in enumerate(cfg.enc_conv_chans):
            conv(f"{base}.conv_blocks.{lvl}.patch_embed.conv", cin, ch, sizes[lvl])
            norm(f"{base}.conv_blocks.{lvl}.patch_embed.norm", ch)
            for j in range(cfg.enc_conv_n_blocks):
                p = f"{base}.conv_blocks.{lvl}.conv.{j}"
                norm(f"{p}.norm1", ch)
                norm(f"{p}.norm2", ch)
                conv(f"{p}.conv1", ch, ch, (1,) * nd)
                conv(f"{p}.conv2", ch, ch, (1,) * nd)
                conv(f"{p}.dw_conv", ch, ch, (5,) * nd, groups=ch)
                conv(f"{p}.mlp.fc1", ch, 4 * ch, (1,) * nd)
                conv(f"{p}.mlp.fc2", 4 * ch, ch, (1,) * nd)
            cin = ch
        linear(f"{base}.patch_embed.proj", cin * math.prod(sizes[-1]), cfg.enc_embed_dim)
        linear(f"{base}.linear", cfg.enc_embed_dim, cfg.enc_embed_dim)
        sd[f"{base}.pos_embed"] = sincos_pos_embed(cfg.enc_embed_dim, cfg.grid_size(v))
        fus = f"enc_fusion_dict.{v}"
        for lvl, ch in enumerate(cfg.enc_conv_chans):
            k = tuple(math.prod(ps[a] for ps in sizes[lvl + 1:]) for a in range(nd))
            conv(f"{fus}.down_convs.{lvl}", ch, cfg.enc_embed_dim, k)
        norm(f"{fus}.norm", cfg.enc_embed_dim)
        sd[f"dec_embed_dict.{v}.pos_embed"] = sincos_pos_embed(cfg.dec_embed_dim, cfg.grid_size(v))
        sd[f"dec_embed_dict.{v}.mask_token"] = torch.randn(1, 1, cfg.dec_embed_dim, generator=g) * 0.02
        linear(f"pred_head_dict.{v}", cfg.dec_embed_dim, math.prod(cfg.dec_patch_size(v)) * cfg.in_chans_dict[v])

    sd["encoder.cls_token"] = torch.randn(1, 1, cfg.enc_embed_dim, generator=g) * 0.02
    for i in range(cfg.enc_depth):
        vit_block(f"encoder.blocks.{i}", cfg.enc_embed_dim)
    norm("encoder.norm", cfg.enc_embed_dim)
    linear("dec_linear", cfg.enc_embed_dim, cfg.dec_embed_dim)
    for i in range(cfg.dec_depth):
        vit_block(f"decoder.blocks.{i}", cfg.dec_embed_dim)
    norm("decoder.norm", cfg.dec_embed_dim)
    return sd


def train_step_cpu(sd: StateDict, cfg: MAEConfig, image_dict: dict[str, Tensor], enc_masks: dict[str, Tensor]):
    """One forward + backward of the reference path on the CPU (fp32; the reference disables autocast on CPU,
    cinema/mae/pretrain.py:251).  Returns (loss, grads)."""
    params = {k: v.detach().clone().requires_grad_(not k.endswith("pos_embed")) for k, v in sd.items()}
    loss, _, _ = mae_forward(params, cfg, image_dict, enc_masks)
    loss.backward()
    return loss.detach(), {k: p.grad for k, p in params.items() if p.grad is not None}
