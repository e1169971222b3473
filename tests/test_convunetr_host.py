"""Host-logic tests (no GPU) of the segmentation model ``ConvUNetR`` (cinema/segmentation/convunetr.py) against golden
vectors generated from the real reference (tests/golden/make_golden.py: convunetr_2view.pt), through the CPU emulation of
the C-ABI.  The ViT encoder runs on the (emulated) kernels, stem and decoder on torch convolutions; tolerances as in
test_model_host.py (bf16 rounding inside the encoder only): 2e-2 relative on logits, 8e-2 on gradients (the key-projection gradients of the tiny random network amplify bf16 rounding most)."""

import pytest
import torch

from cinema_b200.segmentation import ConvUNetR, check_conv_unetr_enc_dec_compatiblity, get_model


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_convunetr_forward_backward_matches_reference_golden(golden_dir, emulated_kernels):
    g = torch.load(golden_dir / "convunetr_2view.pt")
    model = ConvUNetR(**g["kw"])
    assert list(model.state_dict().keys()) == list(g["state_dict"].keys())
    res = model.load_state_dict(g["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert model.n_layers_wo_skip == g["n_layers_wo_skip"]
    model.train()
    preds = model(g["images"])
    assert set(preds) == set(g["preds"])
    for v, ref in g["preds"].items():
        assert preds[v].shape == ref.shape
        assert rel(preds[v], ref) < 2e-2, v
    assert model._dense_levels == [1, 1]  # only the deepest stem map of each view enters the native encoder function
    sum((preds[v] * g["w"][v]).sum() for v in preds).backward()
    named = dict(model.named_parameters())
    for k, ref_g in g["grads"].items():
        assert named[k].grad is not None, k
        assert rel(named[k].grad, ref_g) < 8e-2, (k, rel(named[k].grad, ref_g))
    for k, p in named.items():
        assert (p.grad is not None) == p.requires_grad, k
    # a subset of the views
    with torch.no_grad():
        out = model({"sax": g["images"]["sax"]})
    assert list(out) == ["sax"] and rel(out["sax"], g["sax_only"]) < 2e-2
    with pytest.raises(ValueError):
        model({"lax_3c": g["images"]["lax_2c"]})


def test_convunetr_frozen_encoder(golden_dir, emulated_kernels):
    """Frozen stem + encoder (the reference's ``freeze=True`` fine-tuning): only decoder-side parameters get gradients."""
    g = torch.load(golden_dir / "convunetr_2view.pt")
    model = ConvUNetR(**g["kw"])
    model.load_state_dict(g["state_dict"])
    for p in [*model.enc_down_dict.parameters(), *model.encoder.parameters()]:
        p.requires_grad = False
    preds = model(g["images"])
    sum((preds[v] * g["w"][v]).sum() for v in preds).backward()
    named = dict(model.named_parameters())
    for k, ref_g in g["grads"].items():
        if k.startswith(("enc_down_dict", "encoder")):
            assert named[k].grad is None, k
        else:
            assert rel(named[k].grad, ref_g) < 8e-2, k


def test_encoder_decoder_compatibility_table(golden_dir):
    g = torch.load(golden_dir / "convunetr_2view.pt")
    for args, want in g["compat"].items():
        assert check_conv_unetr_enc_dec_compatiblity(*args) == want, args
    # the known-answer table of the reference's own test (cinema/segmentation/convunetr_test.py:23-35)
    for args, want in [(((4, 4), (2, 2), 2, 4, (2, 2), (2, 2)), (1, 0)), (((4, 1), (2, 1), 2, 4, (2, 1), (2, 1)), (1, 0)),
                       (((4, 4), (2, 2), 2, 5, (2, 2), (2, 2)), (1, 1)), (((2, 2), (2, 2), 2, 4, (2, 2), (2, 2)), (0, 1))]:
        assert check_conv_unetr_enc_dec_compatiblity(*args) == want, args
    # incompatible pyramids raise
    with pytest.raises(ValueError):
        check_conv_unetr_enc_dec_compatiblity((4, 4), (2, 2), 4, 4, (2, 2), (2, 2))  # as many conv layers as decoder levels
    with pytest.raises(ValueError):
        check_conv_unetr_enc_dec_compatiblity((2, 2), (2, 2), 1, 4, (4, 4), (2, 2))  # decoder patch larger than encoder patch
    with pytest.raises(ValueError):
        check_conv_unetr_enc_dec_compatiblity((3, 3), (2, 2), 1, 4, (1, 1), (2, 2))  # 3 is not 1 * 2^k


def test_get_model_from_acdc_style_config():
    cfg = {
        "grad_ckpt": False,
        "data": {"sax": {"patch_size": [64, 64, 4], "in_chans": 1}, "lax": {"patch_size": [64, 64], "in_chans": 1}},
        "model": {"views": ["sax", "lax_4c"], "out_chans": 4,
                  "convunetr": {"size": "tiny", "enc_patch_size": [4, 4, 1], "enc_scale_factor": [2, 2, 1],
                                "enc_conv_chans": [8, 16], "enc_conv_n_blocks": 1, "dec_chans": [4, 8, 16, 32, 64],
                                "dec_patch_size": [2, 2, 1], "dec_scale_factor": [2, 2, 1], "dropout": 0.0, "drop_path": 0.1}},
    }
    model = get_model(cfg)
    assert model.views == ["sax", "lax_4c"] and not model.grad_ckpt
    assert model.pred_head_dict["sax"].out_channels == 4 and model.pred_head_dict["lax_4c"].kernel_size == (1, 1)
    assert len(model.dec_down_blocks_dict["sax"]) == 1 and len(model.dec_conv_blocks_dict["sax"]) == 4
    assert model.encoder.blocks[0].drop_path1.drop_prob == 0.1
    model.set_grad_ckpt(True)
    assert model.decoder_dict["sax"].grad_ckpt and model.enc_down_dict["sax"].grad_ckpt


_UNETR_GRID = [
    ((16, 24), (4, 8), (2, 2), [], (2, 4, 8), (1, 2), (2, 2)),
    ((16, 24), (2, 4), (2, 2), [4], (2, 4, 8), (1, 2), (2, 2)),
    ((16, 24), (1, 2), (2, 2), [4, 8], (2, 4, 8), (1, 2), (2, 2)),
    ((16, 24, 16), (4, 8, 4), (2, 2, 2), [], (2, 4, 8), (1, 2, 1), (2, 2, 2)),
    ((16, 16, 16), (4, 4, 4), (2, 2, 2), [4], (2, 4, 8, 16), (1, 1, 1), (2, 2, 2)),
    ((16, 16, 16), (2, 2, 2), (2, 2, 2), [4, 8], (2, 4, 8, 16), (1, 1, 1), (2, 2, 2)),
    ((16, 16, 1), (2, 2, 1), (2, 2, 1), [4, 8], (2, 4, 8, 16), (1, 1, 1), (2, 2, 1)),
    ((32, 32, 4), (4, 4, 1), (2, 2, 1), [4, 8], (2, 4, 8, 16), (2, 2, 1), (2, 2, 1)),
    ((32, 32, 4), (8, 8, 1), (2, 2, 1), [4, 8], (2, 4, 8, 16, 32), (2, 2, 1), (2, 2, 1)),
    ((32, 32, 4), (4, 4, 1), (2, 2, 1), [4, 8, 16], (2, 4, 8, 16, 32), (2, 2, 1), (2, 2, 1)),
    ((32, 32, 4), (2, 2, 1), (2, 2, 1), [2, 2, 4, 4], (2, 4, 8, 16, 32), (2, 2, 1), (2, 2, 1)),
]


@pytest.mark.parametrize("grid", range(len(_UNETR_GRID)))
def test_reference_test_grid_shapes_convunetr(grid, emulated_kernels):
    """The reference's ``TestConvUNetR.test_single_view`` grid (cinema/segmentation/convunetr_test.py:69-139): models
    without a stem, 1 - 4 stem levels, decoders with and without extra downsampling levels, 2-D and 3-D; logits have the
    image's shape, the grad-ckpt switch propagates, and a backward pass reaches the stem."""
    image_size, eps, esf, chans, dec_chans, dps, dsf = _UNETR_GRID[grid]
    in_chans, out_chans = (3, 4) if grid % 2 else (1, 2)
    torch.manual_seed(0)
    unetr = ConvUNetR(image_size_dict={"view": image_size}, in_chans_dict={"view": in_chans}, out_chans=out_chans,
                      enc_patch_size_dict={"view": eps}, enc_scale_factor_dict={"view": esf}, enc_conv_chans=chans,
                      enc_conv_n_blocks=1, enc_embed_dim=16, enc_depth=len(dec_chans), enc_n_heads=2, dec_chans=dec_chans,
                      dec_patch_size_dict={"view": dps}, dec_scale_factor_dict={"view": dsf}, mlp_ratio=2)
    for flag in (True, False):
        unetr.set_grad_ckpt(flag)
        assert all(m.grad_ckpt == flag for m in unetr.children() if hasattr(m, "grad_ckpt"))
    x = torch.rand(2, in_chans, *image_size)
    logits = unetr({"view": x})["view"]
    assert logits.shape == (2, out_chans, *image_size) and bool(torch.isfinite(logits).all())
    logits.sum().backward()
    first = next(unetr.enc_down_dict["view"].parameters())
    assert first.grad is not None or not first.requires_grad


def test_reference_multi_view_convunetr(emulated_kernels):
    """cinema/segmentation/convunetr_test.py:141-214: a 2-D and a 3-D view through one shared encoder."""
    unetr = ConvUNetR(image_size_dict={"lax": (16, 16), "sax": (16, 24, 16)}, in_chans_dict={"lax": 1, "sax": 1}, out_chans=3,
                      enc_patch_size_dict={"lax": (4, 8), "sax": (4, 8, 4)}, enc_scale_factor_dict={"lax": (2, 2), "sax": (2, 2, 2)},
                      enc_conv_chans=[], enc_conv_n_blocks=1, enc_embed_dim=16, enc_depth=3, enc_n_heads=2, dec_chans=(2, 4, 8),
                      dec_patch_size_dict={"lax": (1, 2), "sax": (1, 2, 1)}, dec_scale_factor_dict={"lax": (2, 2), "sax": (2, 2, 2)},
                      mlp_ratio=2)
    images = {"lax": torch.rand(2, 1, 16, 16), "sax": torch.rand(2, 1, 16, 24, 16)}
    out = unetr(images)
    assert out["lax"].shape == (2, 3, 16, 16) and out["sax"].shape == (2, 3, 16, 24, 16)
    sum(v.sum() for v in out.values()).backward()
    assert unetr.encoder.cls_token.grad is None or bool(torch.isfinite(unetr.encoder.cls_token.grad).all())


def test_native_conv_switch_matches_the_cudnn_path(emulated_kernels, monkeypatch):
    """``ConvUNetR.set_native_convs``: the ConvResBlock convolutions with >= 32 input channels as K-concatenated GEMMs over the
    zero-haloed row space (channels padded to a multiple of 64) give the torch convolution's values to bf16 rounding, forward and
    backward; blocks with fewer input channels (the image block) stay on the torch path."""
    from cinema_b200.conv import ConvResBlock

    torch.manual_seed(0)
    monkeypatch.setattr(ConvResBlock, "_native_ok", lambda self, x: self.native and self.conv1.in_channels >= 32
                        and self.conv1.padding == "same" and tuple(self.conv1.kernel_size) == (3,) * (x.dim() - 2))
    for nd, cin, cout, shape in ((3, 64, 64, (2, 64, 5, 6, 4)), (2, 32, 32, (2, 32, 7, 6)), (3, 96, 40, (1, 96, 4, 4, 3))):
        blk = ConvResBlock(n_dims=nd, in_chans=cin, out_chans=cout, norm="layer")
        x = torch.randn(shape, requires_grad=True)
        ref = blk(x)
        ref.square().sum().backward()
        want = [x.grad.clone(), blk.conv1.weight.grad.clone(), blk.conv2.bias.grad.clone()]
        x.grad = None
        blk.zero_grad()
        blk.native = True
        out = blk(x)
        out.square().sum().backward()
        got = [x.grad, blk.conv1.weight.grad, blk.conv2.bias.grad]
        assert float((out - ref).norm() / ref.norm()) < 6e-3
        for g, w in zip(got, want):
            assert float((g - w).norm() / w.norm()) < 1.5e-2
