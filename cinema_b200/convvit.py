"""ConvMAE-style multi-scale stem and fusion with the reference's module API (cinema/convvit.py:24-291).

``DownsampleEncoder`` / ``MultiScaleFusion`` keep the reference's constructor signatures, attributes
(``patch_sizes``, ``eff_patch_size``, ``patch_embed.grid_size`` ...) and parameter names.  Inside
``CineMA.forward`` their parameters are consumed by the fused B200 path (cinema_b200/mae.py), which
evaluates token embedding and skip fusion on the *visible* tokens only; the ``forward`` methods here
are the standalone (all-token) form used by fine-tuning models.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F  # noqa: N812
from torch import nn

from cinema_b200.conv import Conv2d, Conv3d, ConvNormActBlock, Linear, MaskedConvBlock
from cinema_b200.vit import PatchEmbed, get_pos_embed, init_weights


def upsample_mask(mask: torch.Tensor, scale_factor: tuple[int, ...]) -> torch.Tensor:
    """Nearest-neighbour integer upsampling of a (B, *spatial) mask (cinema/convvit.py:24-51)."""
    if mask.ndim != len(scale_factor) + 1:
        raise ValueError(
            f"mask must have the same number of dimensions as scale_factor except batch, "
            f"got {mask.ndim} and {len(scale_factor)}."
        )
    for axis, f in enumerate(scale_factor):
        if f != 1:
            mask = mask.repeat_interleave(int(f), dim=axis + 1)
    return mask


class DownsampleEncoder(nn.Module):
    """Strided conv stem with masked ConvMAE blocks, then patch embedding to ViT tokens (cinema/convvit.py:54-207)."""

    def __init__(self, image_size, in_chans, patch_size, scale_factor, conv_chans, conv_n_blocks, embed_dim, norm) -> None:
        super().__init__()
        self.grad_ckpt = False
        n_dims = len(image_size)
        self.patch_sizes = [patch_size] + [scale_factor] * len(conv_chans)
        size = tuple(image_size)
        eff = (1,) * n_dims
        chans_in = in_chans
        self.conv_blocks = nn.ModuleList()
        for ps, ch in zip(self.patch_sizes[:-1], conv_chans):
            stage = nn.Module()
            stage.patch_embed = ConvNormActBlock(n_dims=n_dims, in_chans=chans_in, out_chans=ch, norm=norm,
                                                 kernel_size=ps, stride=ps, padding="valid")
            stage.conv = nn.ModuleList([MaskedConvBlock(n_dims=n_dims, in_chans=ch, norm=norm) for _ in range(conv_n_blocks)])
            self.conv_blocks.append(stage)
            size = tuple(s // p for s, p in zip(size, ps))
            eff = tuple(e * p for e, p in zip(eff, ps))
            chans_in = ch
        self.eff_patch_size = tuple(e * p for e, p in zip(eff, self.patch_sizes[-1]))
        self.patch_embed = PatchEmbed(image_size=size, patch_size=self.patch_sizes[-1], in_chans=chans_in, embed_dim=embed_dim)
        self.linear = Linear(embed_dim, embed_dim)
        self.pos_embed = get_pos_embed(embed_dim=embed_dim, grid_size=self.patch_embed.grid_size)
        self.apply(init_weights)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        for stage in self.conv_blocks:
            stage.patch_embed.set_grad_ckpt(enable)
            for blk in stage.conv:
                blk.set_grad_ckpt(enable)
        self.patch_embed.set_grad_ckpt(enable)
        self.linear.set_grad_ckpt(enable)

    def interpolate_pos_encoding(self, grid_size: tuple[int, ...]) -> torch.Tensor:
        """Positional table resampled to another token grid (cinema/convvit.py:139-163)."""
        if tuple(grid_size) == tuple(self.patch_embed.grid_size):
            return self.pos_embed
        mode = {2: "bicubic", 3: "trilinear"}[len(grid_size)]
        d = self.pos_embed.shape[-1]
        pe = self.pos_embed.float().reshape(1, *self.patch_embed.grid_size, d).movedim(-1, 1)
        pe = F.interpolate(pe, size=tuple(grid_size), mode=mode, antialias=False)
        return pe.movedim(1, -1).reshape(1, -1, d).to(self.pos_embed.dtype)

    def conv_masks(self, mask: torch.Tensor | None, grid_size: tuple[int, ...]) -> list[torch.Tensor | None]:
        """Per-level visibility masks (1 = visible) from the ViT-grid mask (cinema/convvit.py:186-192)."""
        if mask is None:
            return [None] * len(self.conv_blocks)
        out: list[torch.Tensor | None] = []
        m = mask.reshape(mask.shape[0], *grid_size)
        for ps in self.patch_sizes[:0:-1]:
            m = upsample_mask(m, scale_factor=ps)
            out.insert(0, ~m)
        return out

    def conv_stem(self, image: torch.Tensor, mask: torch.Tensor | None) -> list[torch.Tensor]:
        """The conv part of forward: the list of per-level feature maps; the last one feeds ``patch_embed``."""
        grid = tuple(s // p for s, p in zip(image.shape[2:], self.eff_patch_size))
        skips = []
        x = image
        for stage, vis in zip(self.conv_blocks, self.conv_masks(mask, grid)):
            x = stage.patch_embed(x)
            for blk in stage.conv:
                x = blk(x, vis)
            skips.append(x)
        return skips

    def forward(self, image: torch.Tensor, mask: torch.Tensor | None):
        """-> (skips, tokens (B, n_patches, D)) over ALL tokens (cinema/convvit.py:165-207)."""
        grid = tuple(s // p for s, p in zip(image.shape[2:], self.eff_patch_size))
        skips = self.conv_stem(image, mask)
        x = self.linear(self.patch_embed(skips[-1] if skips else image)) + self.interpolate_pos_encoding(grid)
        return skips, x


class MultiScaleFusion(nn.Module):
    """x += Conv_{k=s}(skip) per level, then LayerNorm (cinema/convvit.py:210-291)."""

    def __init__(self, image_size, patch_size, scale_factor, conv_chans, embed_dim, norm_layer, norm_eps) -> None:
        super().__init__()
        self.grad_ckpt = False
        n_dims = len(image_size)
        patch_sizes = [patch_size] + [scale_factor] * len(conv_chans)
        grid = tuple(image_size)
        for ps in patch_sizes:
            grid = tuple(s // p for s, p in zip(grid, ps))
        size = tuple(image_size)
        conv_cls = Conv2d if n_dims == 2 else Conv3d
        self.down_convs = nn.ModuleList()
        for i, ch in enumerate(conv_chans):
            size = tuple(s // p for s, p in zip(size, patch_sizes[i]))
            k = tuple(s // g for s, g in zip(size, grid))
            self.down_convs.append(conv_cls(ch, embed_dim, kernel_size=k, stride=k, padding="valid"))
        self.norm = norm_layer(embed_dim, eps=norm_eps)
        self.apply(init_weights)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        for conv in self.down_convs:
            conv.set_grad_ckpt(enable)

    def forward(self, skips: list[torch.Tensor], x: torch.Tensor, mask: torch.Tensor | None) -> torch.Tensor:
        for skip, conv in zip(skips, self.down_convs):
            down = conv(skip).flatten(2).transpose(1, 2)
            if mask is not None:
                down = down[~mask].reshape(x.shape[0], -1, x.shape[-1])
            x = x + down
        return self.norm(x)


def n_tokens_of(grid: tuple[int, ...]) -> int:
    return math.prod(grid)
