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


# ------------------------------------------------------------------------------------------
# ConvViT: the classification / regression fine-tuning model on top of the pre-trained encoder
# ------------------------------------------------------------------------------------------
class _FeatureFn(torch.autograd.Function):
    """All-token encoder features {cls, view...} with a hand-written backward: conv stem -> token embedding ->
    ViT encoder -> multi-scale fusion, the same kernels and bookkeeping as the MAE step (cinema_b200/mae.py) with every
    token visible.  Inputs after ``anchor`` are the dense (cuDNN) skip maps of the views whose stem does not run natively
    (``model._dense_levels``); their gradients are returned to autograd."""

    @staticmethod
    def forward(ctx, model, views, images, anchor, *skips_flat):  # noqa: ARG004 - anchor ties the node to the parameters
        from cinema_b200 import engine, mae, stem  # noqa: F401 - late import: mae imports this module
        from cinema_b200.arena import ensure_arena

        train = any(ctx.needs_input_grad)
        arena = ensure_arena(model)
        arena.refresh_shadow()
        if train:
            arena.prepare_grads()
        counts = model._dense_levels
        skips, pos = [], 0
        for c in counts:
            skips.append([s.detach() for s in skips_flat[pos:pos + c]])
            pos += c
        b, dev = images[0].shape[0], images[0].device
        imgs32 = [im.detach().to(torch.float32).contiguous() for im in images]
        grids, keep, n_keeps, masks = [], [], [], []
        for i, v in enumerate(views):
            down = model.enc_down_dict[v]
            grid = tuple(s // p for s, p in zip(images[i].shape[2:], down.eff_patch_size))
            n = math.prod(grid)
            grids.append(grid), n_keeps.append(n)
            keep.append(engine.arange_idx(b, 0, n, dev))  # every token is visible: slot == token id
            masks.append(torch.zeros((b, n), dtype=torch.bool, device=dev))
        sources, stems = mae._build_sources(model, arena, views, imgs32, skips, keep, masks, keep, grids, n_keeps, b, train)
        _, fused, st = mae._encode(model, arena, views, sources, keep, n_keeps, b, train, True)
        if train:
            ctx.state = dict(model=model, arena=arena, views=views, st=st, n_keeps=n_keeps, b=b, skips=skips, stems=stems,
                             sources=sources, needs=mae._needs(ctx.needs_input_grad[4:], counts))
        return (fused["cls"].clone(), *[fused[v] for v in views])

    @staticmethod
    def backward(ctx, g_cls, *g_views):
        from cinema_b200 import mae, stem

        s = ctx.state
        ctx.state = None
        model, arena, views, st, n_keeps, b = s["model"], s["arena"], s["views"], s["st"], s["n_keeps"], s["b"]
        d = st.d
        d_f32 = torch.empty((b + b * sum(n_keeps), d), dtype=torch.float32, device=g_cls.device)
        d_f32[:b].copy_(g_cls.reshape(b, d))  # rows: [cls (B) | view 0 (B * n) | ...], the layout of mae._encode
        for i, g in enumerate(g_views):
            d_f32[st.foffs[i]:st.foffs[i] + b * n_keeps[i]].copy_(g.reshape(b * n_keeps[i], d))
        targets, dskips = mae._build_targets(views, s["sources"], s["stems"], s["skips"], s["needs"])
        vs = mae._encode_bwd(model, arena, views, st, d_f32, n_keeps, b, targets)
        for i, stem_state in enumerate(s["stems"]):
            if stem_state is not None:
                levels, geo, saved = stem_state
                with vs.view(i):
                    stem.stem_bwd(levels, geo, saved, [t[0] for t in mae._native_grad_buffers(targets[i])])
        vs.join()
        return (None, None, None, None, *[g for per_view in dskips for g in per_view])


class ConvViT(nn.Module):
    """Multi-view ConvMAE-stem ViT for classification / regression (cinema/convvit.py:334-614): same constructor,
    attributes, state-dict keys and outputs.  ``feature_forward`` runs (and differentiates) on the B200 kernels; the
    prediction heads are whatever ``head_layer`` builds (``nn.Linear`` by default), applied to token means in fp32."""

    def __init__(self, image_size_dict, in_chans_dict, n_frames, out_chans, enc_patch_size_dict, enc_scale_factor_dict,
                 enc_conv_chans, enc_conv_n_blocks, enc_embed_dim, enc_depth, enc_n_heads, mlp_ratio=4, qkv_bias=True,
                 norm_layer=nn.LayerNorm, norm_eps=1e-5, rotary=False, act_layer=nn.GELU, mlp_layer=None, drop_path=0.0,
                 norm="layer", head_layer=nn.Linear) -> None:
        super().__init__()
        from cinema_b200.vit import Mlp, ViTEncoder

        self.grad_ckpt = False
        self.native_stem = True  # False: dense cuDNN conv stem (reference evaluation order)
        self._dense_levels: list[int] = []
        self.views = list(image_size_dict.keys())
        self.n_frames = n_frames
        self.enc_down_dict = nn.ModuleDict({
            v: DownsampleEncoder(image_size=image_size_dict[v], in_chans=n_frames * in_chans_dict[v],
                                 patch_size=enc_patch_size_dict[v], scale_factor=enc_scale_factor_dict[v],
                                 conv_chans=enc_conv_chans, conv_n_blocks=enc_conv_n_blocks, embed_dim=enc_embed_dim, norm=norm)
            for v in self.views
        })
        self.enc_fusion_dict = nn.ModuleDict({
            v: MultiScaleFusion(image_size=image_size_dict[v], patch_size=enc_patch_size_dict[v],
                                scale_factor=enc_scale_factor_dict[v], conv_chans=enc_conv_chans, embed_dim=enc_embed_dim,
                                norm_layer=norm_layer, norm_eps=norm_eps)
            for v in self.views
        })
        self.encoder = ViTEncoder(embed_dim=enc_embed_dim, depth=enc_depth, n_heads=enc_n_heads, mlp_ratio=mlp_ratio,
                                  qkv_bias=qkv_bias, norm_layer=norm_layer, norm_eps=norm_eps, rotary=rotary,
                                  act_layer=act_layer, mlp_layer=mlp_layer or Mlp, drop_path=drop_path)
        self.apply(init_weights)
        self.pred_head_dict = nn.ModuleDict()  # built after init_weights, like the reference: default torch initialisation
        if head_layer is not None:
            for v in [*self.views, "cls"]:
                self.pred_head_dict[v] = head_layer(enc_embed_dim, out_chans)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        """API compatibility (cinema/convvit.py:452-462); nothing is recomputed on the B200 path."""
        self.grad_ckpt = enable
        for v in self.views:
            self.enc_down_dict[v].set_grad_ckpt(enable)
            self.enc_fusion_dict[v].set_grad_ckpt(enable)
        self.encoder.set_grad_ckpt(enable)
        for v in [*self.views, "cls"]:
            if v in self.pred_head_dict and hasattr(self.pred_head_dict[v], "set_grad_ckpt"):
                self.pred_head_dict[v].set_grad_ckpt(enable)

    def _backbone_parameters(self):
        for mod in (self.enc_down_dict, self.enc_fusion_dict, self.encoder):
            yield from mod.parameters()

    def feature_forward(self, image_dict: dict[str, torch.Tensor],
                        mask_dict: dict[str, torch.Tensor] | None) -> dict[str, torch.Tensor]:
        """-> {"cls": (B, 1, D), view: (B, n_patches, D)} fp32 (cinema/convvit.py:464-510).  ``mask_dict`` only hides
        patches from the depth-wise convolutions of the stem; every token is still embedded and encoded, so with a
        mask the stem runs densely (cuDNN) and the kernels take over at the token gather."""
        from cinema_b200 import stem

        views = list(image_dict.keys())
        if any(v not in self.views for v in views):
            raise ValueError(f"views {views} must be in self.input_keys {self.views}.")
        first = image_dict[views[0]]
        anchor = next((p for p in self._backbone_parameters() if p.requires_grad), None)
        train = torch.is_grad_enabled() and anchor is not None
        dense = []
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=first.is_cuda), \
                torch.set_grad_enabled(train):
            for v in views:
                down = self.enc_down_dict[v]
                if mask_dict is None and self.native_stem and stem.supported(down):
                    dense.append([])
                else:
                    m = None if mask_dict is None else mask_dict[v].to(device=first.device, dtype=torch.bool)
                    dense.append(down.conv_stem(image_dict[v], m))
        self._dense_levels = [len(x) for x in dense]
        flat = [s for per_view in dense for s in per_view]
        with torch.set_grad_enabled(train):
            cls, *feats = _FeatureFn.apply(self, views, [image_dict[v] for v in views],
                                           anchor if train else first.new_zeros(()), *flat)
        return dict(zip(["cls", *views], [cls, *feats]))

    def forward(self, image_dict: dict[str, torch.Tensor], mask_dict: dict[str, torch.Tensor] | None = None,
                reduce: str = "all") -> torch.Tensor:
        """-> logits (B, out_chans) (cinema/convvit.py:512-561).  ``reduce``: "patch" (mean of per-view heads over the
        patch means), "all" (the same plus the cls head) or "cls"."""
        x = self.feature_forward(image_dict=image_dict, mask_dict=mask_dict)
        if reduce == "cls":
            return self.pred_head_dict["cls"](x["cls"])[:, 0]
        if reduce not in ("patch", "all"):
            raise NotImplementedError(f"Unsupported reduce method {reduce}.")
        logits = [self.pred_head_dict[v](x[v].mean(dim=1, keepdim=True)) for v in self.views]
        if reduce == "all":
            logits.append(self.pred_head_dict["cls"](x["cls"]))
        return torch.concat(logits, dim=1).mean(dim=1)

    @classmethod
    def from_finetuned(cls, config=None, state_dict=None, **kwargs) -> "ConvViT":
        """Fine-tuned model from a reference-format config and state dict.  The reference fetches both from the Hugging
        Face hub (cinema/convvit.py:563-591); offline they are passed in or read from ``config_path`` / ``weights_path``."""
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
    def from_pretrained(cls, config, freeze: bool, ckpt_path=None, **kwargs) -> "ConvViT":  # noqa: ARG003
        """Model from config with the encoder initialised from an MAE checkpoint (cinema/convvit.py:593-613); the
        checkpoint is a local ``.pt`` / ``.safetensors`` file (no hub access here)."""
        from pathlib import Path

        model = get_model(config)
        if ckpt_path is None:
            raise ValueError("ckpt_path (local MAE checkpoint) is required: there is no network to download one")
        views = config["model"]["views"] if isinstance(config, dict) else config.model.views
        return load_pretrain_weights(model=model, views=views, ckpt_path=Path(ckpt_path), freeze=freeze)


def get_model(config) -> ConvViT:
    """Config -> ConvViT as cinema/convvit.py:294-331 (``config`` is a nested dict or an OmegaConf object)."""
    from cinema_b200.mae import _Cfg
    from cinema_b200.vit import get_vit_config

    c = _Cfg(config)
    views = [c.model.views] if isinstance(c.model.views, str) else list(c.model.views)
    vit_config = get_vit_config(c.model.convvit.size)
    data = config["data"] if isinstance(config, dict) else config.data
    get = (lambda k: data.get(k)) if hasattr(data, "get") else (lambda k: getattr(data, k, None))
    if get("class_column") is not None:
        out_chans = len(get(get("class_column")))
    elif get("regression_column") is not None:
        out_chans = 1
    else:
        out_chans = c.model.out_chans
    ndim = {v: 3 if v == "sax" else 2 for v in views}
    model = ConvViT(
        image_size_dict={v: tuple(c.data.sax.patch_size if v == "sax" else c.data.lax.patch_size) for v in views},
        n_frames=c.model.n_frames,
        in_chans_dict={v: c.data.sax.in_chans if v == "sax" else c.data.lax.in_chans for v in views},
        out_chans=out_chans,
        enc_patch_size_dict={v: tuple(c.model.convvit.enc_patch_size[:n]) for v, n in ndim.items()},
        enc_scale_factor_dict={v: tuple(c.model.convvit.enc_scale_factor[:n]) for v, n in ndim.items()},
        enc_conv_chans=list(c.model.convvit.enc_conv_chans), enc_conv_n_blocks=c.model.convvit.enc_conv_n_blocks,
        enc_embed_dim=vit_config["enc_embed_dim"], enc_depth=vit_config["enc_depth"], enc_n_heads=vit_config["enc_n_heads"],
        drop_path=c.model.convvit.drop_path,
    )
    model.set_grad_ckpt(c.grad_ckpt)
    return model


_DROPPED_FROM_MAE = ("mask", "decoder", "_head", "sax", "lax_2c", "lax_3c", "lax_4c", "fusion", "dec_linear", "pos_embed")


def load_pretrain_weights(model: nn.Module, views, ckpt_path, freeze: bool) -> nn.Module:
    """Initialise a fine-tuning model from an MAE checkpoint (cinema/convvit.py:616-704): keep the shared encoder and the
    stems (and fusions, if the model has them) of the requested views, drop decoder-side and positional tensors, tile
    the first stem conv over extra input channels (n_frames > 1 / multi-modal input), verify that exactly the per-view
    positional tables are missing, optionally freeze what was loaded."""
    from pathlib import Path

    ckpt_path = Path(ckpt_path)
    if ckpt_path.suffix == ".pt":
        pretrained = torch.load(ckpt_path, map_location="cpu")["model"]
    elif ckpt_path.suffix == ".safetensors":
        from safetensors.torch import load_file

        pretrained = load_file(str(ckpt_path), device="cpu")
    else:
        raise ValueError(f"Unsupported checkpoint format {ckpt_path.suffix}.")
    views = [views] if isinstance(views, str) else list(views)
    drop = [k for k in _DROPPED_FROM_MAE if k not in views and not (k == "fusion" and hasattr(model, "enc_fusion_dict"))]
    expected_missing = {f"enc_down_dict.{v}.pos_embed" for v in views}
    first_conv = {f"enc_down_dict.{v}.conv_blocks.0.patch_embed.conv.weight": v for v in views}
    state_dict = {}
    for key, val in pretrained.items():
        if any(tag in key for tag in drop):
            continue
        if key in first_conv:
            chans = model.enc_down_dict[first_conv[key]].conv_blocks[0].patch_embed.conv.weight.shape[1]
            if val.shape[1] != chans:
                if val.ndim not in (4, 5):
                    raise ValueError(f"Unsupported weight shape {val.shape}.")
                val = val.repeat(1, chans, *([1] * (val.ndim - 2)))
        state_dict[key] = val
    result = model.load_state_dict(state_dict, strict=False)
    missing = {k for k in result.missing_keys if "decoder" not in k and not k.startswith("dec_") and "head" not in k}
    if missing != expected_missing:
        raise ValueError(f"Missing keys from checkpoint: {sorted(missing)}, expected {sorted(expected_missing)}")
    if result.unexpected_keys:
        raise ValueError(f"Unexpected keys in checkpoint: {result.unexpected_keys}")
    if freeze:
        for name, p in model.named_parameters():
            if name in state_dict:
                p.requires_grad = False
    return model


def get_layer_id_for_vit(name: str, n_layers: int) -> int:
    """Layer index of a parameter for layer-wise lr decay (cinema/convvit.py:707-736): stem / embeddings 0, encoder
    block i -> i + 1, everything else (heads) ``n_layers``."""
    if name.startswith("enc_") or any(tag in name for tag in ("cls_token", "pos_embed", "patch_embed", "view_embed")):
        return 0
    if name.startswith("encoder.blocks"):
        return int(name.split(".")[2]) + 1
    return n_layers


def param_groups_lr_decay(model: nn.Module, no_weight_decay_list: list[str], weight_decay: float, layer_decay: float,
                          out_dir=None) -> list[dict]:
    """AdamW parameter groups with layer-wise lr decay (cinema/convvit.py:739-816): group = (layer id, decay or not),
    ``lr_scale = layer_decay ** (n_layers - layer_id)``; 1-D parameters and the listed names are not decayed.  With
    ``out_dir`` the group -> parameter-name table is written to ``param_group_names.json`` like the reference does."""
    n_layers = len(model.encoder.blocks) + 1
    names: dict[str, dict] = {}
    groups: dict[str, dict] = {}
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        no_decay = p.ndim == 1 or n in no_weight_decay_list
        layer_id = get_layer_id_for_vit(n, n_layers)
        key = f"layer_{layer_id}_{'no_decay' if no_decay else 'decay'}"
        if key not in groups:
            meta = {"lr_scale": layer_decay ** (n_layers - layer_id), "weight_decay": 0.0 if no_decay else weight_decay}
            names[key] = {**meta, "params": []}
            groups[key] = {**meta, "params": []}
        names[key]["params"].append(n)
        groups[key]["params"].append(p)
    if out_dir is not None:
        import json
        from pathlib import Path

        Path(out_dir).mkdir(parents=True, exist_ok=True)
        with open(Path(out_dir) / "param_group_names.json", "w", encoding="utf-8") as f:
            json.dump(names, f, indent=2)
    return list(groups.values())
