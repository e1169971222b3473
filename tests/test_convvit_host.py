"""Host-logic tests (no GPU) of the fine-tuning model ``ConvViT`` (cinema/convvit.py:334-816): driven through the CPU
emulation of the C-ABI (tests/emu_c.py) it must reproduce the golden vectors generated from the real reference
(tests/golden/make_golden.py: convvit_*.pt), including the MAE-checkpoint hand-off and the layer-wise lr-decay groups.

Tolerances as in test_model_host.py: the emulation rounds to bf16 where the kernels do, the goldens are fp32:
3e-2 relative (norm-wise) on features / gradients; logits (means over tokens and views) 3e-2 of their scale.
"""

import math

import pytest
import torch

from cinema_b200 import CineMA, ConvViT
from cinema_b200 import convvit as bconv


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


CASES = ["convvit_2view", "convvit_frames_resized"]


def _model(g):
    model = ConvViT(**g["kw"])
    res = model.load_state_dict(g["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model


@pytest.mark.parametrize("case", CASES)
def test_convvit_forward_matches_reference_golden(case, golden_dir, emulated_kernels):
    g = torch.load(golden_dir / f"{case}.pt")
    model = _model(g)
    assert list(model.state_dict().keys()) == list(g["state_dict"].keys())
    model.eval()
    with torch.no_grad():
        feats = model.feature_forward(g["images"], None)
    assert set(feats) == set(g["feats"])
    for k, ref in g["feats"].items():
        assert feats[k].shape == ref.shape and feats[k].dtype == torch.float32
        assert rel(feats[k], ref) < 2e-2, k
    for reduce, ref in g["logits"].items():
        out = model(g["images"], None, reduce)
        assert out.shape == ref.shape
        assert float((out - ref).abs().max()) < 3e-2 * max(1.0, float(ref.abs().max())), reduce
    with pytest.raises(NotImplementedError):
        model(g["images"], None, "none")
    with pytest.raises(ValueError):
        model({"other": next(iter(g["images"].values()))})


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("masked", [False, True])
def test_convvit_backward_matches_reference_golden(case, masked, golden_dir, emulated_kernels):
    """Gradients of a weighted logit sum: native all-token stem (no mask) and dense stem under a patch mask."""
    g = torch.load(golden_dir / f"{case}.pt")
    model = _model(g)
    model.train()
    out = model(g["images"], g["mask_dict"] if masked else None, "all")
    ref = (g["logits_masked"] if masked else g["logits"])["all"]
    assert float((out - ref).abs().max()) < 3e-2 * max(1.0, float(ref.abs().max()))
    assert model._dense_levels == ([2] * len(model.views) if masked else [0] * len(model.views))
    (out * g["w"]).sum().backward()
    named = dict(model.named_parameters())
    want = g["grads_masked"] if masked else g["grads"]
    assert len(want) >= 9
    for k, ref_g in want.items():
        assert named[k].grad is not None, k
        assert rel(named[k].grad, ref_g) < 4e-2, (k, rel(named[k].grad, ref_g))
    for k, p in named.items():
        assert (p.grad is not None) == p.requires_grad, k


def test_convvit_frozen_backbone_trains_heads_only(golden_dir, emulated_kernels):
    g = torch.load(golden_dir / "convvit_2view.pt")
    model = _model(g)
    for mod in (model.enc_down_dict, model.enc_fusion_dict, model.encoder):
        for p in mod.parameters():
            p.requires_grad = False
    model.train()
    out = model(g["images"], None, "all")
    (out * g["w"]).sum().backward()
    named = dict(model.named_parameters())
    for k, ref_g in g["grads"].items():
        if k.startswith("pred_head_dict"):
            assert rel(named[k].grad, ref_g) < 3e-2, k
        else:
            assert named[k].grad is None, k


def test_set_grad_ckpt_propagates(golden_dir):
    g = torch.load(golden_dir / "convvit_2view.pt")
    model = ConvViT(**g["kw"])
    for flag in (True, False):
        model.set_grad_ckpt(flag)
        assert model.grad_ckpt == flag
        for m in model.children():  # cinema/convvit_test.py:134-137
            if hasattr(m, "grad_ckpt"):
                assert m.grad_ckpt == flag


@pytest.mark.parametrize(("name", "n_layers", "expected"), [
    # known answers of the reference's own test (cinema/convvit_test.py:228-254)
    ("cls_token", 13, 0), ("pos_embed", 13, 0), ("patch_embed", 13, 0), ("view_embed", 13, 0), ("enc_view_embed", 13, 0),
    ("encoder.cls_token", 13, 0), ("patch_embed.proj.weight", 13, 0), ("patch_embed.proj.bias", 13, 0),
    ("encoder.blocks.0.attn.q.weight", 13, 1), ("encoder.blocks.0.attn.kv.weight", 13, 1),
    ("encoder.blocks.0.mlp.fc2.weight", 13, 1), ("encoder.blocks.11.attn.q.weight", 13, 12),
    ("encoder.blocks.11.attn.proj.weight", 13, 12), ("encoder.blocks.11.mlp.fc1.weight", 13, 12),
    ("pred_head_dict.cls.weight", 13, 13),
])
def test_get_layer_id_for_vit(name, n_layers, expected):
    assert bconv.get_layer_id_for_vit(name, n_layers) == expected


@pytest.mark.parametrize("case", CASES)
def test_param_groups_lr_decay_match_reference(case, golden_dir, tmp_path):
    import json

    g = torch.load(golden_dir / f"{case}.pt")
    model = ConvViT(**g["kw"])
    groups = bconv.param_groups_lr_decay(model, ["encoder.cls_token"], weight_decay=0.05, layer_decay=0.75, out_dir=tmp_path)
    names = json.loads((tmp_path / "param_group_names.json").read_text())
    assert names == g["lr_decay_groups"]  # group names, order, lr scales, decay flags, member lists
    assert [len(x["params"]) for x in groups] == [len(x["params"]) for x in names.values()]
    assert sum(len(x["params"]) for x in groups) == sum(1 for p in model.parameters() if p.requires_grad)


def test_load_pretrain_weights_from_mae_checkpoint(golden_dir, tmp_path):
    """MAE checkpoint (the ``mae_small_4view`` golden weights) -> ConvViT with n_frames = 2, frozen backbone: the loaded
    tensors, the channel-tiled first stem conv and the frozen set equal what the reference function produced."""
    t = torch.load(golden_dir / "convvit_transfer.pt")
    mae_g = torch.load(golden_dir / "mae_small_4view.pt")
    mae = CineMA(**mae_g["kw"])
    mae.load_state_dict(mae_g["state_dict"])
    for path, payload in ((tmp_path / "ckpt.pt", None), (tmp_path / "ckpt.safetensors", None)):
        if path.suffix == ".pt":
            torch.save({"model": mae.state_dict()}, path)
        else:
            from safetensors.torch import save_file

            save_file({k: v.contiguous() for k, v in mae.state_dict().items()}, str(path))
        torch.manual_seed(0)
        model = ConvViT(**t["kw"])
        model = bconv.load_pretrain_weights(model, t["views"], path, freeze=True)
        sd = model.state_dict()
        assert list(sd.keys()) == t["keys"]
        assert sorted(n for n, p in model.named_parameters() if not p.requires_grad) == t["frozen"]
        same = sorted(k for k in sd if k in mae_g["state_dict"] and sd[k].shape == mae_g["state_dict"][k].shape
                      and torch.equal(sd[k], mae_g["state_dict"][k]))
        assert same == t["same_as_mae"]
        for k, ref in t["first_conv"].items():
            assert torch.equal(sd[k], ref), k
    # a view the checkpoint does not have / a model whose keys do not line up must be rejected like the reference does
    bad = ConvViT(**{**t["kw"], "enc_depth": t["kw"]["enc_depth"] + 1})
    with pytest.raises(ValueError):
        bconv.load_pretrain_weights(bad, t["views"], tmp_path / "ckpt.pt", freeze=False)


def test_get_model_from_reference_style_config():
    cfg = {
        "grad_ckpt": True,
        "data": {"sax": {"patch_size": [32, 32, 4], "in_chans": 1}, "lax": {"patch_size": [32, 32], "in_chans": 1},
                 "class_column": "pathology", "pathology": ["NOR", "DCM", "HCM", "MINF", "RV"]},
        "model": {"views": ["sax", "lax_4c"], "n_frames": 2, "out_chans": 99,
                  "convvit": {"size": "tiny", "enc_patch_size": [4, 4, 1], "enc_scale_factor": [2, 2, 1],
                              "enc_conv_chans": [8, 16], "enc_conv_n_blocks": 1, "drop_path": 0.0}},
    }
    model = bconv.get_model(cfg)
    assert model.views == ["sax", "lax_4c"] and model.grad_ckpt
    assert model.pred_head_dict["cls"].out_features == 5
    assert model.enc_down_dict["sax"].conv_blocks[0].patch_embed.conv.weight.shape[1] == 2
    assert model.enc_down_dict["lax_4c"].patch_sizes[0] == (4, 4)
    cfg["data"].pop("class_column")
    cfg["data"]["regression_column"] = "lvef"
    assert bconv.get_model(cfg).pred_head_dict["sax"].out_features == 1
    cfg["data"].pop("regression_column")
    cfg["model"]["views"] = "sax"
    m = bconv.get_model(cfg)
    assert m.views == ["sax"] and m.pred_head_dict["cls"].out_features == 99


def _replay_drop_scales(monkeypatch, scales, device="cpu"):
    """Make ``engine.draw_drop_scales`` return the factors the reference drew (golden ``drop_scales``), pair by pair."""
    from cinema_b200 import engine

    queue = [t.to(device) for t in scales]

    def fake(b, drop_prob, scale_by_keep, dev):  # noqa: ARG001
        s1, s2 = queue.pop(0), queue.pop(0)
        assert s1.numel() == b
        return s1.float().contiguous(), s2.float().contiguous()

    monkeypatch.setattr(engine, "draw_drop_scales", fake)
    return queue


def test_convvit_stochastic_depth_matches_reference_golden(golden_dir, emulated_kernels, monkeypatch):
    """drop_path > 0 in training mode (all fine-tuning configs of the reference): with the reference's own per-sample
    factors replayed, logits and gradients match; eval mode is the identity."""
    g = torch.load(golden_dir / "convvit_droppath.pt")
    model = _model(g)
    model.train()
    queue = _replay_drop_scales(monkeypatch, g["drop_scales"])
    out = model(g["images"], None, "all")
    assert not queue  # 2 * depth factor vectors consumed, in order
    ref = g["logits"]["all"]
    assert float((out - ref).abs().max()) < 3e-2 * max(1.0, float(ref.abs().max()))
    (out * g["w"]).sum().backward()
    named = dict(model.named_parameters())
    for k, ref_g in g["grads"].items():
        assert rel(named[k].grad, ref_g) < 4e-2, (k, rel(named[k].grad, ref_g))
    model.eval()
    with torch.no_grad():
        feats = model.feature_forward(g["images"], None)
    for k, f in g["feats"].items():
        assert rel(feats[k], f) < 2e-2, k


def test_drop_scale_distribution_and_module():
    from cinema_b200 import engine
    from cinema_b200.vit import DropPath

    torch.manual_seed(0)
    s1, s2 = engine.draw_drop_scales(4000, 0.25, True, torch.device("cpu"))
    for s in (s1, s2):
        assert all(v == 0.0 or abs(v - 1.0 / 0.75) < 1e-6 for v in s.unique().tolist())
        assert abs(float((s > 0).float().mean()) - 0.75) < 0.03
    assert not torch.equal(s1, s2)
    dp = DropPath(0.5)
    x = torch.ones(64, 3, 5)
    dp.train()
    y = dp(x)
    assert set(y.unique().tolist()) <= {0.0, 2.0} and (y.flatten(1).std(dim=1) == 0).all()  # whole samples dropped
    dp.eval()
    assert dp(x) is x


_SINGLE_VIEW_GRID = [
    ((16, 24), (16, 24), (4, 8), (2, 2), []), ((16, 24), (16, 24), (2, 4), (2, 2), [4]), ((16, 24), (16, 24), (1, 2), (2, 2), [4, 8]),
    ((24, 32), (16, 24), (1, 2), (2, 2), [4, 8]), ((16, 24), (24, 32), (1, 2), (2, 2), [4, 8]),
    ((16, 16, 16), (16, 16, 16), (4, 4, 4), (2, 2, 2), [4]), ((16, 16, 16), (16, 16, 16), (2, 2, 2), (2, 2, 2), [4, 8]),
    ((16, 16, 1), (16, 16, 1), (2, 2, 1), (2, 2, 1), [4, 8]), ((32, 24, 12), (16, 16, 16), (2, 2, 2), (2, 2, 2), [4]),
    ((32, 32, 4), (32, 32, 4), (4, 4, 1), (2, 2, 1), [4, 8]), ((32, 32, 4), (32, 32, 4), (8, 8, 1), (2, 2, 1), [4, 8]),
    ((32, 32, 4), (32, 32, 4), (4, 4, 1), (2, 2, 1), [4, 8, 16]), ((32, 32, 4), (32, 32, 4), (2, 2, 1), (2, 2, 1), [2, 2, 4, 4]),
]


@pytest.mark.parametrize("input_mask", [True, False])
@pytest.mark.parametrize("grid", range(len(_SINGLE_VIEW_GRID)))
def test_reference_test_grid_shapes_convvit(grid, input_mask, emulated_kernels):
    """The reference's ``TestConvViT.test_single_view`` grid (cinema/convvit_test.py:72-150): no stem / 1 - 4 stem levels,
    anisotropic patches, inputs smaller and larger than the configured size (resampled positional table), multi-frame
    multi-channel inputs, every reduce mode, with and without a stem mask; plus a backward pass."""
    import math

    image_size, input_size, ps, sf, chans = _SINGLE_VIEW_GRID[grid]
    n_frames, in_chans = (3, 3) if grid % 2 else (1, 1)
    torch.manual_seed(0)
    vit = ConvViT(image_size_dict={"sax": image_size}, n_frames=n_frames, in_chans_dict={"sax": in_chans}, out_chans=3,
                  enc_patch_size_dict={"sax": ps}, enc_scale_factor_dict={"sax": sf}, enc_conv_chans=chans, enc_conv_n_blocks=1,
                  enc_embed_dim=16, enc_depth=2, enc_n_heads=2, mlp_ratio=2)
    x = torch.rand(2, n_frames * in_chans, *input_size)
    n_patches = math.prod(s // p for s, p in zip(input_size, vit.enc_down_dict["sax"].eff_patch_size))
    mask_dict = {"sax": torch.rand(2, n_patches) > 0.5} if input_mask else None
    for reduce in ("patch", "all", "cls"):
        out = vit({"sax": x}, mask_dict=mask_dict, reduce=reduce)
        assert out.shape == (2, 3) and bool(torch.isfinite(out).all())
    out.sum().backward()
    assert vit.encoder.cls_token.grad is not None and bool(torch.isfinite(vit.encoder.cls_token.grad).all())


def test_finetune_loop_with_layerwise_lr_decay(golden_dir, emulated_kernels):
    """cinema_b200/examples/finetune.py: ConvViT + layer-wise lr-decay AdamW groups + cosine schedule with per-group
    lr_scale fits a fixed toy batch; the lr seen by every group is schedule x its lr_scale (cinema/optim.py:47-51)."""
    from cinema_b200.examples import finetune as ft
    from cinema_b200.train import cosine_lr

    g = torch.load(golden_dir / "convvit_2view.pt")
    model = _model(g)
    opt = ft.build_optimizer(model, lr=1e-3, weight_decay=0.05, layer_decay=0.75)
    scales = sorted({grp["lr_scale"] for grp in opt.param_groups})
    assert len(scales) == model.encoder.blocks.__len__() + 2 and abs(scales[-1] - 1.0) < 1e-12  # layers 0 .. depth + 1
    label = torch.tensor([0, 2])
    batches = [(g["images"], label)] * 4
    first = ft.finetune_one_epoch(model, batches, opt, ft.classification_loss, epoch=1, n_batches=4, n_epochs=10, n_warmup_epochs=1,
                                  lr=1e-3, min_lr=1e-6, clip_grad=1.0)
    want = cosine_lr(1 + 3 / 4, 1, 10, 1e-3, 1e-6)
    for grp in opt.param_groups:
        assert abs(grp["lr"] - want * grp["lr_scale"]) < 1e-12
    later = ft.finetune_one_epoch(model, batches, opt, ft.classification_loss, epoch=2, n_batches=4, n_epochs=10, n_warmup_epochs=1,
                                  lr=1e-3, min_lr=1e-6, clip_grad=1.0)
    assert later[-1] < first[0]
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_segmentation_loss_definition(emulated_kernels):
    """Foreground soft Dice + CE (cinema/segmentation/train.py:77-103): perfect one-hot logits -> ~0, uniform logits -> the
    hand-computed Dice of classes 1.. + log C cross-entropy."""
    from cinema_b200.examples import finetune as ft

    label = torch.randint(0, 3, (2, 6, 6, 4))
    perfect = torch.nn.functional.one_hot(label, 3).movedim(-1, 1).float() * 50.0
    assert float(ft.segmentation_loss(perfect, label)) < 1e-3
    uniform = torch.zeros(2, 3, 6, 6, 4)
    dice = []
    for b in range(2):  # per sample and class: 1 - (2 * n_c / 3) / (144 / 3 + n_c), probabilities are 1/3 everywhere
        for c in range(1, 3):  # include_background=False
            n_c = float((label[b] == c).sum())
            dice.append(1.0 - (2.0 * n_c / 3 + 1e-5) / (144 / 3 + n_c + 1e-5))
    want = math.log(3) + sum(dice) / len(dice)
    assert abs(float(ft.segmentation_loss(uniform, label)) - want) < 1e-5
