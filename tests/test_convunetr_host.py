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
    # known answers of the reference's own test (cinema/segmentation/convunetr_test.py:23-63): incompatible pyramids raise
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
