"""Generate golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

The reference (``/root/reference``, pure Python) is imported unmodified.  Its missing
third-party imports are satisfied by the small stub modules below:
``timm.layers`` {Mlp, DropPath, SwiGLU, to_2tuple, use_fused_attn},
``timm.models.vision_transformer`` {LayerScale}, ``omegaconf`` {DictConfig, OmegaConf}.
``cinema/__init__.py`` is skipped (it pulls monai) by registering an empty package whose
``__path__`` points at the reference tree.

The fixtures are small (tiny / reduced dims) so they can be committed; they carry inputs,
masks, the reference state_dict, and the reference outputs + a few gradients.  The GPU box
has no ``/root/reference``; tests only read the ``.pt`` files.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import torch
from torch import nn

REF = Path("/root/reference")
DROP_LOG: list = []
OUT = Path(__file__).resolve().parent


def install_shims() -> None:
    if "cinema" in sys.modules:
        return
    pkg = types.ModuleType("cinema")
    pkg.__path__ = [str(REF / "cinema")]
    sys.modules["cinema"] = pkg
    mae_pkg = types.ModuleType("cinema.mae")
    mae_pkg.__path__ = [str(REF / "cinema" / "mae")]
    sys.modules["cinema.mae"] = mae_pkg
    seg_pkg = types.ModuleType("cinema.segmentation")
    seg_pkg.__path__ = [str(REF / "cinema" / "segmentation")]
    sys.modules["cinema.segmentation"] = seg_pkg

    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    class Mlp(nn.Module):  # timm 1.0.15 layers/mlp.py semantics
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                     norm_layer=None, bias=True, drop=0.0, use_conv=False):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            bias = to_2tuple(bias)
            drop = to_2tuple(drop)
            self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
            self.act = act_layer()
            self.drop1 = nn.Dropout(drop[0])
            self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
            self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])
            self.drop2 = nn.Dropout(drop[1])

        def forward(self, x):
            return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

    class SwiGLU(nn.Module):
        pass

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            t = x.new_empty(shape).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                t.div_(keep)
            DROP_LOG.append(t.flatten().clone())  # per-sample factors in call order, replayed by the parity tests
            return x * t

    class LayerScale(nn.Module):
        def __init__(self, dim, init_values=1e-5, inplace=False):
            super().__init__()
            self.gamma = nn.Parameter(init_values * torch.ones(dim))

        def forward(self, x):
            return x * self.gamma

    layers.Mlp, layers.SwiGLU, layers.DropPath, layers.to_2tuple = Mlp, SwiGLU, DropPath, to_2tuple
    layers.use_fused_attn = lambda: True
    vt.LayerScale = LayerScale
    timm.layers, timm.models, models.vision_transformer = layers, models, vt
    sys.modules.update({"timm": timm, "timm.layers": layers, "timm.models": models,
                        "timm.models.vision_transformer": vt})

    oc = types.ModuleType("omegaconf")

    class DictConfig(dict):
        pass

    class OmegaConf:
        @staticmethod
        def load(path):
            raise NotImplementedError

    oc.DictConfig, oc.OmegaConf = DictConfig, OmegaConf
    sys.modules["omegaconf"] = oc


def build_reference_mae(kw: dict, seed: int):
    install_shims()
    from cinema.mae.mae import CineMA  # type: ignore

    torch.manual_seed(seed)
    return CineMA(**kw)


CASES = {
    # cfg-1 plumbing case of BASELINE.json (tiny dims, one SAX view) -- reference `tiny` size
    "mae_tiny_sax": dict(
        kw=dict(
            image_size_dict={"sax": (32, 32, 4)}, in_chans_dict={"sax": 1},
            enc_patch_size_dict={"sax": (4, 4, 1)}, enc_scale_factor_dict={"sax": (2, 2, 1)},
            enc_conv_chans=[8, 16], enc_conv_n_blocks=1,
            enc_embed_dim=16, enc_depth=1, enc_n_heads=2, dec_embed_dim=16, dec_depth=1, dec_n_heads=2,
        ),
        batch=2, ratio=0.75,
    ),
    # multi-view, head_dim 32 / 64 like ViT-B (enc 128/2 heads -> d=64, dec 64/2 -> d=32), two blocks each
    "mae_small_4view": dict(
        kw=dict(
            image_size_dict={"sax": (64, 64, 2), "lax_2c": (64, 64), "lax_3c": (64, 64), "lax_4c": (64, 64)},
            in_chans_dict={"sax": 1, "lax_2c": 1, "lax_3c": 1, "lax_4c": 1},
            enc_patch_size_dict={"sax": (4, 4, 1), "lax_2c": (4, 4), "lax_3c": (4, 4), "lax_4c": (4, 4)},
            enc_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2), "lax_3c": (2, 2), "lax_4c": (2, 2)},
            enc_conv_chans=[16, 32], enc_conv_n_blocks=2,
            enc_embed_dim=128, enc_depth=2, enc_n_heads=2, dec_embed_dim=64, dec_depth=2, dec_n_heads=2,
        ),
        batch=2, ratio=0.75,
    ),
    # self-attention decoder + normalised targets
    "mae_tiny_selfattn_normtarget": dict(
        kw=dict(
            image_size_dict={"sax": (32, 32, 2), "lax_2c": (32, 32)}, in_chans_dict={"sax": 1, "lax_2c": 1},
            enc_patch_size_dict={"sax": (4, 4, 1), "lax_2c": (4, 4)},
            enc_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)},
            enc_conv_chans=[8, 16], enc_conv_n_blocks=1,
            enc_embed_dim=32, enc_depth=2, enc_n_heads=2, dec_embed_dim=32, dec_depth=1, dec_n_heads=4,
            cross_attn=False, norm_target=True,
        ),
        batch=3, ratio=0.5,
    ),
    # same two switches at the head dims the tcgen05 attention kernels are built for (64 encoder / 32 decoder)
    "mae_hd_selfattn_normtarget": dict(
        kw=dict(
            image_size_dict={"sax": (32, 32, 2), "lax_2c": (48, 48)}, in_chans_dict={"sax": 1, "lax_2c": 1},
            enc_patch_size_dict={"sax": (4, 4, 1), "lax_2c": (4, 4)},
            enc_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)},
            enc_conv_chans=[8, 16], enc_conv_n_blocks=1,
            enc_embed_dim=64, enc_depth=2, enc_n_heads=1, dec_embed_dim=64, dec_depth=2, dec_n_heads=2,
            cross_attn=False, norm_target=True,
        ),
        batch=3, ratio=0.5,
    ),
}

GRAD_KEYS = [
    "encoder.blocks.0.attn.q.weight", "encoder.blocks.0.attn.kv.bias", "encoder.blocks.0.mlp.fc1.weight",
    "encoder.cls_token", "dec_linear.weight", "decoder.blocks.0.attn.kv.weight", "decoder.norm.weight",
]


def make_mae_case(name: str, spec: dict) -> None:
    install_shims()
    from cinema.mae import mae as ref_mae  # type: ignore

    model = build_reference_mae(spec["kw"], seed=0)
    model.train()
    g = torch.Generator().manual_seed(1)
    images = {v: torch.rand(spec["batch"], 1, *s, generator=g) for v, s in spec["kw"]["image_size_dict"].items()}
    # masks come from the reference function under a fixed seed, then get injected into forward
    torch.manual_seed(2)
    masks = {
        v: ref_mae.get_batch_random_patch_mask(spec["batch"], model.enc_down_dict[v].patch_embed.n_patches,
                                               spec["ratio"], torch.device("cpu"))
        for v in images
    }
    queue = [masks[v] for v in images]
    orig = ref_mae.get_batch_random_patch_mask
    ref_mae.get_batch_random_patch_mask = lambda **_: queue.pop(0)
    try:
        loss, preds, mask_out, metrics = model(images, spec["ratio"])
    finally:
        ref_mae.get_batch_random_patch_mask = orig
    loss.backward()
    named = dict(model.named_parameters())
    firsts = [k for k in named if k.endswith("conv_blocks.0.patch_embed.conv.weight")][:1]
    grads = {k: named[k].grad.clone() for k in GRAD_KEYS + firsts if k in named and named[k].grad is not None}
    model.eval()
    with torch.no_grad():
        feats = model.feature_forward(images)
    torch.save(
        dict(
            kw=spec["kw"], batch=spec["batch"], ratio=spec["ratio"],
            state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
            images=images, masks=masks,
            loss=loss.detach(), preds={k: v.detach() for k, v in preds.items()},
            metrics={k: v.detach() for k, v in metrics.items()}, grads=grads,
            feats={k: v.detach() for k, v in feats.items()},
        ),
        OUT / f"{name}.pt",
    )
    print(name, float(loss), {k: tuple(v.shape) for k, v in preds.items()})


CONVVIT_CASES = {
    # two views (3-D SAX + 2-D LAX), head_dim 32, two stem levels, 3 classes
    "convvit_2view": dict(
        kw=dict(
            image_size_dict={"sax": (32, 32, 4), "lax_2c": (32, 32)}, in_chans_dict={"sax": 1, "lax_2c": 1}, n_frames=1,
            out_chans=3, enc_patch_size_dict={"sax": (4, 4, 1), "lax_2c": (4, 4)},
            enc_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)}, enc_conv_chans=[8, 16], enc_conv_n_blocks=1,
            enc_embed_dim=64, enc_depth=2, enc_n_heads=2,
        ),
        batch=2, input_size=None,
    ),
    # video-style input (n_frames * in_chans channels), head_dim 64, input larger than the configured image size
    # (interpolated positional table), regression head
    "convvit_frames_resized": dict(
        kw=dict(
            image_size_dict={"sax": (32, 32, 2)}, in_chans_dict={"sax": 2}, n_frames=3, out_chans=1,
            enc_patch_size_dict={"sax": (4, 4, 1)}, enc_scale_factor_dict={"sax": (2, 2, 1)}, enc_conv_chans=[8, 16],
            enc_conv_n_blocks=2, enc_embed_dim=64, enc_depth=1, enc_n_heads=1, mlp_ratio=2,
        ),
        batch=3, input_size={"sax": (48, 32, 2)},
    ),
    # stochastic depth as in the fine-tuning configs (drop_path > 0, training mode): the drawn per-sample factors are
    # recorded (``drop_scales``, call order = attention branch, MLP branch per block) and replayed by the tests
    "convvit_droppath": dict(
        kw=dict(
            image_size_dict={"sax": (32, 32, 4), "lax_2c": (32, 32)}, in_chans_dict={"sax": 1, "lax_2c": 1}, n_frames=1,
            out_chans=2, enc_patch_size_dict={"sax": (4, 4, 1), "lax_2c": (4, 4)},
            enc_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)}, enc_conv_chans=[8, 16], enc_conv_n_blocks=1,
            enc_embed_dim=64, enc_depth=3, enc_n_heads=2, drop_path=0.4,
        ),
        batch=4, input_size=None,
    ),
}

CONVVIT_GRAD_KEYS = [
    "encoder.blocks.0.attn.q.weight", "encoder.blocks.0.mlp.fc2.bias", "encoder.cls_token", "encoder.norm.weight",
    "enc_fusion_dict.sax.down_convs.0.weight", "enc_fusion_dict.sax.norm.bias", "enc_down_dict.sax.linear.weight",
    "enc_down_dict.sax.conv_blocks.0.patch_embed.conv.weight", "enc_down_dict.sax.conv_blocks.1.conv.0.dw_conv.weight",
    "pred_head_dict.sax.weight", "pred_head_dict.cls.bias",
]


def make_convvit_case(name: str, spec: dict) -> None:
    """ConvViT (classification / regression fine-tuning model, cinema/convvit.py:334) forward for every ``reduce``,
    with and without a stem mask, plus gradients of a weighted logit sum."""
    install_shims()
    from cinema.convvit import ConvViT, param_groups_lr_decay  # type: ignore

    kw = spec["kw"]
    torch.manual_seed(10)
    model = ConvViT(**kw)
    model.train()  # drop_path = 0: train and eval agree
    g = torch.Generator().manual_seed(11)
    sizes = spec["input_size"] or kw["image_size_dict"]
    images = {v: torch.rand(spec["batch"], kw["n_frames"] * kw["in_chans_dict"][v], *sizes[v], generator=g) for v in sizes}
    mask_dict = {}
    for v in sizes:
        grid = tuple(s // p for s, p in zip(sizes[v], model.enc_down_dict[v].eff_patch_size))
        n = 1
        for x in grid:
            n *= x
        mask_dict[v] = torch.rand(spec["batch"], n, generator=g) > 0.5
    named = dict(model.named_parameters())
    w = torch.rand(spec["batch"], kw["out_chans"], generator=g)
    drop_scales = None
    if kw.get("drop_path", 0.0) > 0.0:
        torch.manual_seed(13)
        del DROP_LOG[:]
        out = model(images, None, "all")
        drop_scales = [t.clone() for t in DROP_LOG]
        assert len(drop_scales) == 2 * kw["enc_depth"] and any(bool((t == 0).any()) for t in drop_scales)
        (out * w).sum().backward()
        logits, logits_masked, grads_masked = {"all": out.detach()}, {}, {}
        grads = {k: named[k].grad.clone() for k in CONVVIT_GRAD_KEYS if k in named and named[k].grad is not None}
        model.eval()
        feats = {k: v.detach() for k, v in model.feature_forward(images, None).items()}  # eval: DropPath is the identity
        model.train()
    else:
        logits = {r: model(images, None, r).detach() for r in ("patch", "all", "cls")}
        logits_masked = {r: model(images, mask_dict, r).detach() for r in ("all",)}
        feats = {k: v.detach() for k, v in model.feature_forward(images, None).items()}
        out = model(images, None, "all")
        (out * w).sum().backward()
        grads = {k: named[k].grad.clone() for k in CONVVIT_GRAD_KEYS if k in named and named[k].grad is not None}
        model.zero_grad()
        out = model(images, mask_dict, "all")
        (out * w).sum().backward()
        grads_masked = {k: named[k].grad.clone() for k in CONVVIT_GRAD_KEYS if k in named and named[k].grad is not None}
    import json
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        groups = param_groups_lr_decay(model, ["encoder.cls_token"], weight_decay=0.05, layer_decay=0.75, out_dir=Path(tmp))
        group_names = json.loads((Path(tmp) / "param_group_names.json").read_text())
    assert len(groups) == len(group_names)
    torch.save(
        dict(kw=kw, images=images, mask_dict=mask_dict, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
             logits=logits, logits_masked=logits_masked, feats=feats, w=w, grads=grads, grads_masked=grads_masked,
             lr_decay_groups=group_names, drop_scales=drop_scales),
        OUT / f"{name}.pt",
    )
    print(name, {r: v.flatten()[:3].tolist() for r, v in logits.items()})


def make_pretrain_transfer_case() -> None:
    """MAE checkpoint -> fine-tuning model hand-off (cinema/convvit.py:616-704) on the ``mae_small_4view`` weights: which
    keys load, the channel-tiled first stem conv for n_frames = 2, which parameters end up frozen."""
    install_shims()
    import tempfile

    from cinema.convvit import ConvViT, load_pretrain_weights  # type: ignore

    spec = CASES["mae_small_4view"]
    mae = build_reference_mae(spec["kw"], seed=0)
    views = ["sax", "lax_4c"]
    kw = dict(
        image_size_dict={v: spec["kw"]["image_size_dict"][v] for v in views}, in_chans_dict={v: 1 for v in views}, n_frames=2,
        out_chans=2, enc_patch_size_dict={v: spec["kw"]["enc_patch_size_dict"][v] for v in views},
        enc_scale_factor_dict={v: spec["kw"]["enc_scale_factor_dict"][v] for v in views},
        enc_conv_chans=spec["kw"]["enc_conv_chans"], enc_conv_n_blocks=spec["kw"]["enc_conv_n_blocks"],
        enc_embed_dim=spec["kw"]["enc_embed_dim"], enc_depth=spec["kw"]["enc_depth"], enc_n_heads=spec["kw"]["enc_n_heads"],
    )
    torch.manual_seed(12)
    model = ConvViT(**kw)
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "ckpt.pt"
        torch.save({"model": mae.state_dict()}, path)
        model = load_pretrain_weights(model, views, path, freeze=True)
    frozen = sorted(n for n, p in model.named_parameters() if not p.requires_grad)
    sd = model.state_dict()
    mae_sd = mae.state_dict()
    same_as_mae = sorted(k for k in sd if k in mae_sd and sd[k].shape == mae_sd[k].shape and torch.equal(sd[k], mae_sd[k]))
    first = {k: sd[k].clone() for k in sd if k.endswith("conv_blocks.0.patch_embed.conv.weight")}
    torch.save(dict(kw=kw, views=views, frozen=frozen, same_as_mae=same_as_mae, first_conv=first, keys=list(sd.keys())),
               OUT / "convvit_transfer.pt")
    print("convvit_transfer", len(frozen), "frozen,", len(same_as_mae), "tensors identical to the MAE checkpoint")


CONVUNETR_CASE = dict(
    # segmentation fine-tuning model (cinema/segmentation/convunetr.py:214): 3-D SAX + 2-D LAX, the decoder pyramid of
    # the reference's ACDC config scaled down (dec_chans one level deeper than the ViT grid -> one extra down block)
    kw=dict(
        image_size_dict={"sax": (32, 32, 4), "lax_2c": (32, 32)}, in_chans_dict={"sax": 1, "lax_2c": 1}, out_chans=3,
        enc_patch_size_dict={"sax": (4, 4, 1), "lax_2c": (4, 4)}, enc_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)},
        enc_conv_chans=[8, 16], enc_conv_n_blocks=1, enc_embed_dim=64, enc_depth=2, enc_n_heads=2,
        dec_chans=(4, 8, 8, 16, 16), dec_patch_size_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)},
        dec_scale_factor_dict={"sax": (2, 2, 1), "lax_2c": (2, 2)},
    ),
    batch=2,
)

CONVUNETR_GRAD_KEYS = [
    "encoder.blocks.0.attn.kv.weight", "encoder.blocks.1.mlp.fc1.bias", "encoder.cls_token", "encoder.norm.bias",
    "enc_down_dict.sax.linear.weight", "enc_down_dict.sax.patch_embed.proj.weight",
    "enc_down_dict.lax_2c.conv_blocks.0.patch_embed.conv.weight", "enc_down_dict.sax.conv_blocks.1.conv.0.mlp.fc2.weight",
    "dec_down_blocks_dict.sax.0.weight", "dec_conv_blocks_dict.sax.2.conv1.weight", "dec_image_conv_block_dict.lax_2c.conv1.weight",
    "decoder_dict.sax.blocks.0.up.weight", "pred_head_dict.sax.weight",
]


def make_convunetr_case() -> None:
    install_shims()
    from cinema.segmentation.convunetr import ConvUNetR, check_conv_unetr_enc_dec_compatiblity  # type: ignore

    spec = CONVUNETR_CASE
    kw = spec["kw"]
    torch.manual_seed(20)
    model = ConvUNetR(**kw)
    model.train()
    g = torch.Generator().manual_seed(21)
    images = {v: torch.rand(spec["batch"], kw["in_chans_dict"][v], *s, generator=g) for v, s in kw["image_size_dict"].items()}
    preds = model(images)
    w = {v: torch.rand(p.shape, generator=g) for v, p in preds.items()}
    sum((preds[v] * w[v]).sum() for v in preds).backward()
    named = dict(model.named_parameters())
    grads = {k: named[k].grad.clone() for k in CONVUNETR_GRAD_KEYS}
    with torch.no_grad():
        sax_only = model({"sax": images["sax"]})["sax"]
    compat = {}
    for args in [((4, 4, 1), (2, 2, 1), 2, 5, (2, 2, 1), (2, 2, 1)), ((4, 4), (2, 2), 2, 4, (2, 2), (2, 2)),
                 ((4, 4), (2, 2), 1, 4, (1, 1), (2, 2)), ((2, 2), (2, 2), 2, 5, (1, 1), (2, 2))]:
        compat[args] = check_conv_unetr_enc_dec_compatiblity(*args)
    torch.save(dict(kw=kw, images=images, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
                    preds={k: v.detach() for k, v in preds.items()}, w=w, grads=grads, sax_only=sax_only, compat=compat,
                    n_layers_wo_skip=model.n_layers_wo_skip),
               OUT / "convunetr_2view.pt")
    print("convunetr_2view", {k: tuple(v.shape) for k, v in preds.items()}, compat)


def make_op_vectors() -> None:
    """Small op-level vectors from the reference functions themselves."""
    install_shims()
    from cinema import rotary as ref_rot  # type: ignore
    from cinema import vit as ref_vit  # type: ignore
    from cinema.convvit import upsample_mask  # type: ignore
    from cinema.mae.mae import get_batch_random_patch_mask, mse_loss  # type: ignore

    g = torch.Generator().manual_seed(3)
    out: dict = {}
    # patchify 2/3/4-D
    for nm, shape, ps in [("p2", (2, 3, 8, 12), (2, 4)), ("p3", (2, 2, 8, 6, 4), (4, 2, 1)),
                          ("p4", (1, 2, 4, 4, 2, 6), (2, 2, 1, 3))]:
        img = torch.rand(*shape, generator=g)
        tok = ref_vit.patchify(img, ps)
        grid = tuple(s // p for s, p in zip(shape[2:], ps))
        back = ref_vit.unpatchify(tok, ps, grid)
        out[nm] = dict(image=img, patch_size=ps, grid=grid, tokens=tok, back=back)
    # sincos pos-embed incl. a non-square grid and the 3-D zero-pad case
    out["pos"] = {f"{d}_{gs}": ref_vit.get_pos_embed(d, gs).data.clone()
                  for d, gs in [(16, (2, 3)), (32, (4, 4)), (512, (2, 3, 4)), (768, (12, 12, 16))][:3]}
    out["pos_sum_768_12_12_16"] = ref_vit.get_pos_embed(768, (12, 12, 16)).data.double().sum(dim=1)
    # rotary: standalone contract, partial rotary dim
    q = torch.rand(2, 7, 3, 12, generator=g)
    k = torch.rand(2, 7, 3, 12, generator=g)
    rot = ref_rot.RotaryEmbedding(10)
    rq, rk = rot(q, k)
    out["rotary"] = dict(q=q, k=k, rq=rq, rk=rk, dim=10)
    # attention with rotary=True vs False (reference quirk: identical up to fp error)
    torch.manual_seed(4)
    att = ref_vit.Attention(32, n_heads=4, qkv_bias=True, rotary=True)
    x = torch.rand(2, 9, 32, generator=g)
    out["attn_rotary"] = dict(sd={k2: v.detach().clone() for k2, v in att.state_dict().items()}, x=x,
                              y=att(x).detach())
    # cross attention, q != k
    torch.manual_seed(5)
    att2 = ref_vit.Attention(64, n_heads=2, qkv_bias=True)
    xq, xk = torch.rand(2, 11, 64, generator=g), torch.rand(2, 5, 64, generator=g)
    out["attn_cross"] = dict(sd={k2: v.detach().clone() for k2, v in att2.state_dict().items()}, q=xq, k=xk,
                             y=att2(xq, xk).detach())
    # upsample_mask
    m = torch.rand(2, 3, 2, 2, generator=g) > 0.5
    out["upsample"] = dict(mask=m, sf=(2, 2, 1), out=upsample_mask(m, (2, 2, 1)))
    # random mask keep counts
    torch.manual_seed(6)
    out["mask_counts"] = {f"{n}_{r}": get_batch_random_patch_mask(4, n, r, torch.device("cpu"))
                          for n, r in [(16, 0.75), (37, 0.6), (144, 0.75), (9, 0.0)]}
    # loss
    tgt, pred = torch.rand(2, 10, 8, generator=g), torch.rand(2, 6, 8, generator=g)
    mk = torch.zeros(2, 10, dtype=torch.bool)
    mk[0, [0, 2, 3, 5, 7, 9]] = True
    mk[1, [1, 2, 4, 6, 8, 9]] = True
    for nt in (False, True):
        loss, met = mse_loss(tgt, pred, mk, nt)
        out[f"loss_norm{int(nt)}"] = dict(target=tgt, pred=pred, mask=mk, loss=loss, metrics=met)
    torch.save(out, OUT / "ops.pt")
    print("ops.pt written")


def main() -> None:
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if not only or name in only:
            make_mae_case(name, spec)
    for name, spec in CONVVIT_CASES.items():
        if not only or name in only:
            make_convvit_case(name, spec)
    if not only or "convunetr_2view" in only:
        make_convunetr_case()
    if not only or "convvit_transfer" in only:
        make_pretrain_transfer_case()
    if not only or "ops" in only:
        make_op_vectors()


if __name__ == "__main__":
    main()
