"""The ConvMAE stem (``DownsampleEncoder.conv_blocks``, cinema/convvit.py:87-113,186-201) evaluated on the
visible ViT patches only, token-major and channel-last, entirely through the C-ABI kernels.

Per stem level: kernel==stride patch conv (GEMM over gathered patches) -> channel LayerNorm + GELU -> n x
MaskedConvBlock, each ``x += conv2(dw5(conv1(LN x)))`` then ``x += fc2(GELU(fc1(LN x)))`` (cinema/conv.py:400-415): the
1x1 convs and the MLP are row-wise GEMMs with fused bias / GELU / residual epilogues, the depth-wise 5^n conv
looks its neighbours up through the token map (masked and out-of-image neighbours are zero, which is what
``mask * conv1(...)`` and "same" padding do in the reference).  Exactness of the visible-only evaluation:
csrc/stem.cu header.

Row layout of level l: ``x[(b, i) * P_l + p, c]`` with P_l = prod(f_l), f_l = level positions per ViT token.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from cinema_b200 import _C, engine
from cinema_b200.conv import ConvLayerNorm

F32, BF16 = torch.float32, torch.bfloat16


@dataclass
class DwW:
    w16: torch.Tensor
    bias: torch.Tensor | None
    gw: torch.Tensor | None
    gb: torch.Tensor | None


@dataclass
class ConvBlockW:
    norm1: engine.NormW
    conv1: engine.LinW
    dw: DwW
    conv2: engine.LinW
    norm2: engine.NormW
    fc1: engine.LinW
    fc2: engine.LinW


@dataclass
class LevelW:
    conv: engine.LinW
    norm: engine.NormW
    blocks: list[ConvBlockW]
    patch: tuple[int, ...]  # kernel == stride of this level's patch conv
    f: tuple[int, ...]      # level positions per ViT token per axis
    chans: int


def supported(down) -> bool:
    """The native stem implements the reference default (``norm="layer"``, GELU, no dropout / drop-path)."""
    for stage in down.conv_blocks:
        if not isinstance(stage.patch_embed.norm, ConvLayerNorm):
            return False
        if any(not isinstance(blk.drop_path, torch.nn.Identity) for blk in stage.conv):
            return False
    return True


def stem_weights(arena, down, train: bool) -> list[LevelW]:
    lin = lambda m: engine.linw(arena, m.weight, m.bias, train)  # noqa: E731
    levels = []
    stride = [1] * len(down.eff_patch_size)
    for stage, ps in zip(down.conv_blocks, down.patch_sizes[:-1]):
        stride = [s * p for s, p in zip(stride, ps)]
        f = tuple(e // s for e, s in zip(down.eff_patch_size, stride))
        blocks = []
        for blk in stage.conv:
            dwp = blk.dw_conv
            dw = DwW(arena.w16(dwp.weight), dwp.bias.data if dwp.bias is not None else None,
                     arena.grad_view(dwp.weight) if train and dwp.weight.requires_grad else None,
                     arena.grad_view(dwp.bias) if train and dwp.bias is not None and dwp.bias.requires_grad else None)
            blocks.append(ConvBlockW(engine.normw(arena, blk.norm1, train), lin(blk.conv1), dw, lin(blk.conv2),
                                     engine.normw(arena, blk.norm2, train), lin(blk.mlp.fc1), lin(blk.mlp.fc2)))
        levels.append(LevelW(lin(stage.patch_embed.conv), engine.normw(arena, stage.patch_embed.norm, train), blocks,
                             tuple(ps), f, stage.patch_embed.conv.weight.shape[0]))
    return levels


def level_view(x: torch.Tensor, n_tokens: int, f: tuple[int, ...]) -> torch.Tensor:
    """(T * P, C) token-major rows as a strided (T, C, *f) tensor: what the patch gather / scatter kernels address."""
    return x.view(n_tokens, *f, x.shape[-1]).movedim(-1, 1)


class Geometry:
    """Token map of one view for one call: visible ids, mask, slot and the ViT token grid."""

    def __init__(self, keep, mask, slot, grid_tok, b, nk) -> None:
        self.keep, self.mask, self.slot, self.grid_tok, self.b, self.nk = keep, mask, slot, tuple(grid_tok), b, nk


def _block_fwd(x, w: ConvBlockW, geo: Geometry, f, save: bool):
    h, _, mean1, rstd1 = engine.ln_fwd(x, w.norm1, stats=save)
    h1 = engine.linear_fwd(h, w.conv1)
    h2 = torch.empty_like(h1)
    _C.dwconv_tokens(h1, h2, w.dw.w16, w.dw.bias, geo.mask, geo.slot, geo.keep, geo.grid_tok, f)
    x1 = engine.linear_fwd(h2, w.conv2, out_dtype=F32, residual=x)
    g, _, mean2, rstd2 = engine.ln_fwd(x1, w.norm2, stats=save)
    pre, act = engine.linear_gelu_fwd(g, w.fc1)
    x2 = engine.linear_fwd(act, w.fc2, out_dtype=F32, residual=x1)
    return x2, ((x, mean1, rstd1, h, h1, h2, x1, mean2, rstd2, g, pre, act) if save else None)


def _block_bwd(dx32, dx16, w: ConvBlockW, geo: Geometry, f, saved, fc2_bias_done: bool = False, out_bias=None):
    """``fc2_bias_done`` / ``out_bias``: bias gradients fused into the producing LayerNorm backward (engine.block_bwd)."""
    x, mean1, rstd1, h, h1, h2, x1, mean2, rstd2, g, pre, act = saved
    c = x.shape[1]
    dpre = engine.linear_bwd(dx16, act, w.fc2, gelu_aux=pre, bias_done=fc2_bias_done, dx_colsum=w.fc1.gb)
    dg = engine.linear_bwd(dpre, g, w.fc1, bias_done=w.fc1.gb is not None)
    conv2_gb = engine.fusable_bias(w.conv2, c)
    dx32, dx16 = engine.ln_bwd(dg, x1, mean2, rstd2, w.norm2, dres=dx32, dx32=dx32, dxsum=conv2_gb)
    dh2 = engine.linear_bwd(dx16, h2, w.conv2, bias_done=conv2_gb is not None)
    if w.dw.gw is not None:
        _C.dwconv_tokens_wgrad(h1, dh2, w.dw.gw, w.dw.gb, geo.mask, geo.slot, geo.keep, geo.grid_tok, f)
    dh1 = torch.empty_like(dh2)
    _C.dwconv_tokens(dh2, dh1, w.dw.w16, None, geo.mask, geo.slot, geo.keep, geo.grid_tok, f, transpose=True)
    dh = engine.linear_bwd(dh1, h, w.conv1)
    return engine.ln_bwd(dh, x, mean1, rstd1, w.norm1, dres=dx32, dx32=dx32, dxsum=out_bias)


def stem_fwd(levels: list[LevelW], image32: torch.Tensor, geo: Geometry, save: bool):
    """image32 (B, Cin, *spatial) fp32 contiguous -> (list of level outputs x_l (T * P_l, C_l) fp32, saved state)."""
    t = geo.b * geo.nk
    outs, saved = [], []
    prev = None
    for li, lw in enumerate(levels):
        p_l = math.prod(lw.f)
        if li == 0:
            idx = _C.expand_token_index(geo.keep, geo.grid_tok, lw.f)
            level_grid = tuple(g * ff for g, ff in zip(geo.grid_tok, lw.f))
            rows = torch.empty((t * p_l, image32.shape[1] * math.prod(lw.patch)), dtype=BF16, device=image32.device)
            _C.gather_patches(image32, level_grid, lw.patch, idx, False, rows)
        else:
            pw = levels[li - 1]
            rows = torch.empty((t * p_l, pw.chans * math.prod(lw.patch)), dtype=BF16, device=image32.device)
            _C.gather_patches(level_view(prev, t, pw.f), lw.f, lw.patch, None, False, rows)
        y = engine.linear_fwd(rows, lw.conv, out_dtype=F32)
        _, x, mean, rstd = engine.ln_fwd(y, lw.norm, want16=False, want32=True, stats=save, act=True)
        bsaved = []
        for bw in lw.blocks:
            x, sv = _block_fwd(x, bw, geo, lw.f, save)
            bsaved.append(sv)
        outs.append(x)
        saved.append((rows, y, mean, rstd, bsaved) if save else None)
        prev = x
    return outs, saved


def stem_bwd(levels: list[LevelW], geo: Geometry, saved, dlevels: list[torch.Tensor]) -> None:
    """dlevels[l]: fp32 gradient of level l's output (already holding the fusion / patch-embed contributions);
    propagates down the stem, accumulating parameter gradients."""
    t = geo.b * geo.nk
    for li in range(len(levels) - 1, -1, -1):
        lw = levels[li]
        rows, y, mean, rstd, bsaved = saved[li]
        dx32 = dlevels[li]
        dx16 = torch.empty(dx32.shape, dtype=BF16, device=dx32.device)
        _C.cast_bf16(dx32, dx16)
        gb_of = lambda bi: engine.fusable_bias(lw.blocks[bi].fc2, lw.chans) if bi >= 0 else None  # noqa: E731,B023
        for bi in range(len(lw.blocks) - 1, -1, -1):
            done = bi < len(lw.blocks) - 1 and gb_of(bi) is not None  # the level's entry dx16 is a plain cast
            dx32, dx16 = _block_bwd(dx32, dx16, lw.blocks[bi], geo, lw.f, bsaved[bi], fc2_bias_done=done,
                                    out_bias=gb_of(bi - 1))
            bsaved[bi] = None
        # LayerNorm + GELU backward (dy of the patch conv), then the conv's wgrad / dgrad
        conv_gb = engine.fusable_bias(lw.conv, lw.chans)
        _, dy16 = engine.ln_bwd(dx32, y, mean, rstd, lw.norm, dx32=dx32, beta_act=lw.norm.beta, dxsum=conv_gb)
        drows = engine.linear_bwd(dy16, rows, lw.conv, need_dx=li > 0, bias_done=conv_gb is not None)
        if li > 0:
            pw = levels[li - 1]
            _C.scatter_patches(drows, level_view(dlevels[li - 1], t, pw.f), lw.f, lw.patch, None, False, accumulate=True)
        saved[li] = None
