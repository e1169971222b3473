"""ConvUNetR: UNETR-style segmentation model on the pre-trained encoder (cinema/segmentation/convunetr.py:25-534), with
the reference's constructor, attributes, state-dict keys and outputs.

Where the time goes (BASELINE.json config 4, SAX 192 x 192 x 16, ViT-B): the ViT encoder sees ALL 2304 + 1 tokens --
391 GF of Linear layers + 196 GF of attention per sample -- and runs on the tcgen05 GEMM / attention kernels of the MAE
path (token embedding -> 12 blocks -> final LayerNorm, hand-written backward, stochastic depth in the GEMM epilogue).
The conv stem has to be materialised densely here (its feature maps ARE the U-Net skips) and the decoder is dense
3 x 3 (x 3) convolutions / transposed convolutions up to full resolution: both run on cuDNN under bf16 autocast
("+ U-Net decoder convs (stock cuDNN)").
"""

from __future__ import annotations

import math

import torch
from torch import nn

from cinema_b200.conv import Conv2d, Conv3d, ConvResBlock, ConvTranspose2d, ConvTranspose3d
from cinema_b200.convvit import DownsampleEncoder, load_pretrain_weights
from cinema_b200.vit import Mlp, ViTEncoder, get_vit_config, init_weights


class UpsampleDecoder(nn.Module):
    """Transposed-conv upsampling with additive skips and residual conv blocks (cinema/segmentation/convunetr.py:25-106)."""

    def __init__(self, n_dims, chans, patch_size, scale_factor, norm, kernel_size: int = 3, n_blocks: int = 2,
                 dropout: float = 0.0) -> None:
        if n_dims not in {2, 3}:
            raise ValueError(f"Invalid n_dims, must be 2 or 3, got {n_dims}.")
        super().__init__()
        self.grad_ckpt = False
        deconv_cls = ConvTranspose2d if n_dims == 2 else ConvTranspose3d
        self.blocks = nn.ModuleList()
        n = len(chans)
        for i, ch in enumerate(chans[::-1]):  # deepest level first
            last = i == n - 1
            out_ch = ch if last else chans[-i - 2]
            up = patch_size if last else scale_factor
            block = nn.Module()
            block.up = deconv_cls(ch, out_ch, kernel_size=up, stride=up)
            block.conv = nn.ModuleList([
                ConvResBlock(n_dims=n_dims, in_chans=out_ch, out_chans=out_ch, dropout=dropout, kernel_size=kernel_size, norm=norm)
                for _ in range(n_blocks)
            ])
            self.blocks.append(block)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        for block in self.blocks:
            block.up.set_grad_ckpt(enable)
            for conv in block.conv:
                conv.set_grad_ckpt(enable)

    def forward(self, embeddings: list[torch.Tensor | None]) -> torch.Tensor:
        """``embeddings``: per-level features from full resolution to the deepest, ``None`` where a level has no skip;
        consumed from the end (the list is emptied, like the reference's)."""
        x = embeddings.pop()
        for block in self.blocks:
            x = block.up(x)
            skip = embeddings.pop()
            if skip is not None:
                x = x + skip
            for conv in block.conv:
                x = conv(x)
        return x


def check_conv_unetr_enc_dec_compatiblity(enc_patch_size, enc_scale_factor, enc_n_conv_layers, dec_depth, dec_patch_size,
                                          dec_scale_factor) -> tuple[int, int]:
    """-> (decoder levels finer than the first encoder level, i.e. without skip; extra downsampling levels below the ViT
    grid), or ValueError when the encoder and decoder pyramids cannot be aligned
    (cinema/segmentation/convunetr.py:109-166)."""
    if enc_n_conv_layers >= dec_depth:
        raise ValueError(f"enc_n_conv_layers {enc_n_conv_layers} must be less than dec_depth {dec_depth}.")
    if any(f < s for f, s in zip(enc_patch_size, dec_patch_size)):
        raise ValueError(f"enc_patch_size {enc_patch_size} must be greater than dec_patch_size {dec_patch_size}.")
    enc_patch_size, enc_scale_factor = tuple(enc_patch_size), tuple(enc_scale_factor)
    dec_patch_size, dec_scale_factor = tuple(dec_patch_size), tuple(dec_scale_factor)
    enc_total = tuple(p * s ** enc_n_conv_layers for p, s in zip(enc_patch_size, enc_scale_factor))
    n_layers_wo_skip = n_downsample_layers = None
    factor = dec_patch_size
    for i in range(dec_depth):
        if factor == enc_patch_size:
            n_layers_wo_skip = i
        if factor == enc_total:
            n_downsample_layers = dec_depth - 1 - i
        factor = tuple(f * s for f, s in zip(factor, dec_scale_factor))
    if n_layers_wo_skip is None:
        raise ValueError(f"enc_patch_size {enc_patch_size} must be equal to dec_patch_size {dec_patch_size} times certain "
                         f"number of {dec_scale_factor} .")
    if n_downsample_layers is None:
        raise ValueError(f"enc_factor {enc_total} must be equal to dec_patch_size {dec_patch_size} times certain number of "
                         f"{dec_scale_factor} .")
    return n_layers_wo_skip, n_downsample_layers


class _EncoderFn(torch.autograd.Function):
    """Token embedding of every view -> shared ViT encoder (cls + all patches) on the B200 kernels, hand-written backward.
    The inputs after ``anchor`` are the deepest dense stem map of each view (or nothing for a view without stem); their
    gradients go back to autograd, which continues through the dense stem."""

    @staticmethod
    def forward(ctx, model, views, images, anchor, *last_skips):  # noqa: ARG004
        from cinema_b200 import engine, mae
        from cinema_b200.arena import ensure_arena

        train = any(ctx.needs_input_grad)
        arena = ensure_arena(model)
        arena.refresh_shadow()
        if train:
            arena.prepare_grads()
        counts = model._dense_levels
        skips, pos = [], 0
        for c in counts:
            skips.append([s.detach().contiguous() for s in last_skips[pos:pos + c]])  # the gather kernels read NC(D)HW
            pos += c
        b, dev = images[0].shape[0], images[0].device
        imgs32 = [im.detach().to(torch.float32).contiguous() for im in images]
        grids, keep, n_keeps, masks = [], [], [], []
        for i, v in enumerate(views):
            down = model.enc_down_dict[v]
            grid = tuple(s // p for s, p in zip(images[i].shape[2:], down.eff_patch_size))
            n = math.prod(grid)
            grids.append(grid), n_keeps.append(n)
            keep.append(engine.arange_idx(b, 0, n, dev))
            masks.append(None)
        sources, stems = mae._build_sources(model, arena, views, imgs32, skips, keep, masks, keep, grids, n_keeps, b, train)
        _, out, st = mae._encode(model, arena, views, sources, keep, n_keeps, b, train, True, fuse=False)
        if train:
            ctx.state = dict(model=model, arena=arena, views=views, st=st, n_keeps=n_keeps, b=b, skips=skips, stems=stems,
                             sources=sources, needs=mae._needs(ctx.needs_input_grad[4:], counts))
        return out["enc"]

    @staticmethod
    def backward(ctx, g):
        from cinema_b200 import mae

        s = ctx.state
        ctx.state = None
        targets, dskips = mae._build_targets(s["views"], s["sources"], s["stems"], s["skips"], s["needs"])
        vs = mae._encode_bwd(s["model"], s["arena"], s["views"], s["st"], g.to(torch.float32), s["n_keeps"], s["b"], targets,
                             fuse=False)
        vs.join()
        return (None, None, None, None, *[x for per_view in dskips for x in per_view])


class ConvUNetR(nn.Module):
    """ConvUNetR (cinema/segmentation/convunetr.py:214-534): per-view conv stem, shared ViT encoder over the tokens of all
    views, per-view UNETR decoder -> {view: logits (B, out_chans, *image_size)}."""

    def __init__(self, image_size_dict, in_chans_dict, out_chans, enc_patch_size_dict, enc_scale_factor_dict, enc_conv_chans,
                 enc_conv_n_blocks, enc_embed_dim, enc_depth, enc_n_heads, dec_chans, dec_patch_size_dict,
                 dec_scale_factor_dict, dec_kernel_size: int = 3, mlp_ratio: int = 4, qkv_bias: bool = True,
                 norm_layer=nn.LayerNorm, norm_eps: float = 1e-5, rotary: bool = False, act_layer=nn.GELU, mlp_layer=None,
                 dropout: float = 0.0, drop_path: float = 0.0, norm: str = "layer", channels_last: bool = True) -> None:
        """``channels_last`` (extension, default on): the dense stem and the decoder keep their feature maps in channel-last
        strides between the channel LayerNorms instead of copying back to NC(D)HW after each one (see ``ConvLayerNorm``);
        shapes, values and the state dict are unaffected."""
        super().__init__()
        self.grad_ckpt = False
        self._dense_levels: list[int] = []
        self.views = list(image_size_dict.keys())
        for v in self.views:
            if len(image_size_dict[v]) not in {2, 3}:
                raise ValueError(f"Invalid image_size for {v}, must be 2D or 3D, got {image_size_dict[v]}.")
        align = [check_conv_unetr_enc_dec_compatiblity(
            enc_patch_size=enc_patch_size_dict[v], enc_scale_factor=enc_scale_factor_dict[v],
            enc_n_conv_layers=len(enc_conv_chans), dec_depth=len(dec_chans), dec_patch_size=dec_patch_size_dict[v],
            dec_scale_factor=dec_scale_factor_dict[v]) for v in self.views]
        if len({a[0] for a in align}) != 1:
            raise ValueError(f"n_layers_wo_skip_list {[a[0] for a in align]} must be the same for all views.")
        if len({a[1] for a in align}) != 1:
            raise ValueError(f"n_downsample_layers_list {[a[1] for a in align]} must be the same for all views.")
        self.n_layers_wo_skip, n_down = align[0]

        self.enc_down_dict = nn.ModuleDict({
            v: DownsampleEncoder(image_size=image_size_dict[v], in_chans=in_chans_dict[v], patch_size=enc_patch_size_dict[v],
                                 scale_factor=enc_scale_factor_dict[v], conv_chans=enc_conv_chans,
                                 conv_n_blocks=enc_conv_n_blocks, embed_dim=enc_embed_dim, norm=norm)
            for v in self.views
        })
        self.encoder = ViTEncoder(embed_dim=enc_embed_dim, depth=enc_depth, n_heads=enc_n_heads, mlp_ratio=mlp_ratio,
                                  qkv_bias=qkv_bias, norm_layer=norm_layer, norm_eps=norm_eps, rotary=rotary,
                                  act_layer=act_layer, mlp_layer=mlp_layer or Mlp, drop_path=drop_path)
        self.dec_image_conv_block_dict = nn.ModuleDict()
        self.dec_down_blocks_dict = nn.ModuleDict()
        self.dec_conv_blocks_dict = nn.ModuleDict()
        self.decoder_dict = nn.ModuleDict()
        self.pred_head_dict = nn.ModuleDict()
        for v in self.views:
            nd = len(image_size_dict[v])
            conv_cls = Conv2d if nd == 2 else Conv3d
            res = lambda cin, cout: ConvResBlock(n_dims=nd, in_chans=cin, out_chans=cout, kernel_size=dec_kernel_size,  # noqa: E731,B023
                                                 dropout=dropout, act_layer=act_layer, norm=norm)
            self.dec_image_conv_block_dict[v] = res(in_chans_dict[v], dec_chans[0])
            self.dec_down_blocks_dict[v] = nn.ModuleList([
                conv_cls(enc_embed_dim, enc_embed_dim, kernel_size=dec_scale_factor_dict[v], stride=dec_scale_factor_dict[v],
                         padding="valid") for _ in range(n_down)
            ])
            first = self.n_layers_wo_skip
            blocks = [res(ch, dec_chans[first + i]) for i, ch in enumerate(enc_conv_chans)]  # stem skips
            blocks += [res(enc_embed_dim, dec_chans[first + len(enc_conv_chans) + i]) for i in range(n_down + 1)]  # ViT grid and below
            self.dec_conv_blocks_dict[v] = nn.ModuleList(blocks)
            self.decoder_dict[v] = UpsampleDecoder(n_dims=nd, chans=dec_chans, patch_size=dec_patch_size_dict[v],
                                                   scale_factor=dec_scale_factor_dict[v], norm=norm)
            self.pred_head_dict[v] = conv_cls(dec_chans[0], out_chans, kernel_size=1)
        self.apply(init_weights)
        self.set_channels_last(channels_last)

    def set_channels_last(self, enable: bool = True) -> None:
        from cinema_b200.conv import ConvLayerNorm

        self.channels_last = enable
        for m in self.modules():
            if isinstance(m, ConvLayerNorm):
                m.keep_channels_last = enable

    def set_native_convs(self, enable: bool = True) -> None:
        """Opt-in: every ``ConvResBlock`` of the skip blocks and of the UNETR decoder whose 3^n convolutions have at least 32
        input channels runs them as K-concatenated tcgen05 GEMMs over a zero-haloed channel-last row space
        (``cinema_b200.conv_gemm``) instead of cuDNN; the others, the dense stem and the transposed convolutions stay on
        cuDNN.  Values match the cuDNN path to bf16 rounding; state dict and shapes are unaffected."""
        for m in self.modules():
            if isinstance(m, ConvResBlock):
                m.native = enable

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        """API compatibility (cinema/segmentation/convunetr.py:421-434); activations are kept, not recomputed."""
        self.grad_ckpt = enable
        self.encoder.set_grad_ckpt(enable)
        for v in self.views:
            self.enc_down_dict[v].set_grad_ckpt(enable)
            self.dec_image_conv_block_dict[v].set_grad_ckpt(enable)
            for blk in [*self.dec_down_blocks_dict[v], *self.dec_conv_blocks_dict[v]]:
                blk.set_grad_ckpt(enable)
            self.decoder_dict[v].set_grad_ckpt(enable)
            self.pred_head_dict[v].set_grad_ckpt(enable)

    def encode(self, image_dict: dict[str, torch.Tensor]):
        """-> (per-view lists of dense stem maps, encoder output (B, 1 + sum n_patches, D) fp32)."""
        views = list(image_dict.keys())
        if any(v not in self.views for v in views):
            raise ValueError(f"views {views} must be in self.input_keys {self.views}.")
        skips = [self.enc_down_dict[v].conv_stem(image_dict[v], None) for v in views]
        last = [sk[-1:] for sk in skips]
        self._dense_levels = [len(x) for x in last]
        first = image_dict[views[0]]
        backbone = [*self.enc_down_dict.parameters(), *self.encoder.parameters()]
        anchor = next((p for p in backbone if p.requires_grad), None)
        train = torch.is_grad_enabled() and (anchor is not None or any(s.requires_grad for per in last for s in per))
        with torch.set_grad_enabled(train):
            x = _EncoderFn.apply(self, views, [image_dict[v] for v in views],
                                 anchor if anchor is not None else first.new_zeros(()), *[s for per in last for s in per])
        return skips, x

    def forward(self, image_dict: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        """{view: (B, in_chans, *image_size)} -> {view: logits (B, out_chans, *image_size)}
        (cinema/segmentation/convunetr.py:436-485)."""
        views = list(image_dict.keys())
        first = image_dict[views[0]] if views else None
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=first is not None and first.is_cuda):
            skips, x = self.encode(image_dict)
            ns = [math.prod(s // p for s, p in zip(image_dict[v].shape[2:], self.enc_down_dict[v].eff_patch_size)) for v in views]
            tokens = torch.split(x, [1, *ns], dim=1)[1:]  # the cls token is not used by the decoder
            preds = {}
            for i, v in enumerate(views):
                grid = tuple(s // p for s, p in zip(image_dict[v].shape[2:], self.enc_down_dict[v].eff_patch_size))
                xv = tokens[i].permute(0, 2, 1).reshape(x.shape[0], x.shape[2], *grid)
                levels = [*skips[i], xv]
                for down in self.dec_down_blocks_dict[v]:
                    xv = down(xv)
                    levels.append(xv)
                emb = [self.dec_image_conv_block_dict[v](image_dict[v]), *([None] * self.n_layers_wo_skip)]
                emb += [blk(levels[j]) for j, blk in enumerate(self.dec_conv_blocks_dict[v])]
                preds[v] = self.pred_head_dict[v](self.decoder_dict[v](emb))
        return preds

    @classmethod
    def from_finetuned(cls, config=None, state_dict=None, **kwargs) -> "ConvUNetR":
        """Fine-tuned model from a reference-format config and state dict; the reference downloads both from the Hugging
        Face hub (cinema/segmentation/convunetr.py:487-520), offline they are passed in or read from local files."""
        if config is None:
            import yaml

            with open(kwargs["config_path"]) as f:
                config = yaml.safe_load(f)
        model = get_model(config)
        if state_dict is None and "weights_path" in kwargs:
            from safetensors.torch import load_file

            state_dict = load_file(kwargs["weights_path"])
        if state_dict is not None:
            model.load_state_dict(state_dict)
        return model

    @classmethod
    def from_pretrained(cls, config, freeze: bool, ckpt_path=None, **kwargs) -> "ConvUNetR":  # noqa: ARG003
        """Model from config with stem + encoder initialised from a local MAE checkpoint
        (cinema/segmentation/convunetr.py:522-534 with the hub download replaced by ``ckpt_path``)."""
        from pathlib import Path

        if ckpt_path is None:
            raise ValueError("ckpt_path (local MAE checkpoint) is required: there is no network to download one")
        model = get_model(config)
        views = config["model"]["views"] if isinstance(config, dict) else config.model.views
        return load_pretrain_weights(model=model, views=views, ckpt_path=Path(ckpt_path), freeze=freeze)


def get_model(config) -> ConvUNetR:
    """Config -> ConvUNetR as cinema/segmentation/convunetr.py:169-211 (nested dict or OmegaConf-like object)."""
    from cinema_b200.mae import _Cfg

    c = _Cfg(config)
    data = config["data"] if isinstance(config, dict) else config.data
    has_lax = ("lax" in data) if hasattr(data, "keys") else hasattr(data, "lax")

    def view_cfg(v: str):
        if v == "sax":
            return c.data.sax
        return c.data.lax if has_lax else getattr(c.data, v)

    views = [c.model.views] if isinstance(c.model.views, str) else list(c.model.views)
    m = c.model.convunetr
    vit = get_vit_config(m.size)
    nd = {v: 3 if v == "sax" else 2 for v in views}
    model = ConvUNetR(
        image_size_dict={v: tuple(view_cfg(v).patch_size) for v in views},
        in_chans_dict={v: view_cfg(v).in_chans for v in views},
        out_chans=c.model.out_chans,
        enc_patch_size_dict={v: tuple(m.enc_patch_size[:n]) for v, n in nd.items()},
        enc_scale_factor_dict={v: tuple(m.enc_scale_factor[:n]) for v, n in nd.items()},
        enc_conv_chans=list(m.enc_conv_chans), enc_conv_n_blocks=m.enc_conv_n_blocks,
        enc_embed_dim=vit["enc_embed_dim"], enc_depth=vit["enc_depth"], enc_n_heads=vit["enc_n_heads"],
        dec_chans=tuple(m.dec_chans),
        dec_patch_size_dict={v: tuple(m.dec_patch_size[:n]) for v, n in nd.items()},
        dec_scale_factor_dict={v: tuple(m.dec_scale_factor[:n]) for v, n in nd.items()},
        dropout=m.dropout, drop_path=m.drop_path,
    )
    model.set_grad_ckpt(c.grad_ckpt)
    return model
