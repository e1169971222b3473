"""The reference's ``cinema/conv_test.py`` shape cases on the cinema_b200 conv modules (the dense forms used by fine-tuning
models and by the cuDNN stem path; CPU, no kernels involved) + numerics of each block against the oracle restatement."""

import pytest
import torch

from cinema_b200.conv import ConvLayerNorm, ConvMlp, ConvNormActBlock, ConvResBlock, MaskedConvBlock, get_conv_norm

SIZES = [(16, 8), (7, 16), (16, 8, 7)]


@pytest.mark.parametrize("image_size", SIZES)
@pytest.mark.parametrize("chans", [1, 8])
def test_conv_mlp_layernorm_and_norms(chans, image_size):  # cinema/conv_test.py:16-56
    x = torch.rand(2, chans, *image_size)
    mlp = ConvMlp(n_dims=len(image_size), in_features=chans)
    for flag in (True, False):
        mlp.set_grad_ckpt(flag)
        assert all(m.grad_ckpt == flag for m in mlp.children() if hasattr(m, "grad_ckpt"))
    assert mlp(x).shape == x.shape
    ln = ConvLayerNorm(chans)
    y = ln(x)
    assert y.shape == x.shape and y.is_contiguous()
    ln.keep_channels_last = True
    z = ln(x)
    assert z.shape == x.shape and torch.equal(z, y)  # same values, channel-last strides
    for norm in ("instance", "layer", "group"):
        assert get_conv_norm(n_dims=len(image_size), in_chans=chans, norm=norm)(x).shape == x.shape
    with pytest.raises(ValueError):
        get_conv_norm(n_dims=len(image_size), in_chans=chans, norm="batch")


@pytest.mark.parametrize("norm", ["instance", "layer", "group"])
@pytest.mark.parametrize("image_size", SIZES)
@pytest.mark.parametrize(("in_chans", "out_chans"), [(1, 8), (8, 8), (8, 1)])
def test_conv_norm_act_and_res_blocks(in_chans, out_chans, image_size, norm):  # cinema/conv_test.py:59-106
    x = torch.rand(2, in_chans, *image_size)
    for block in (ConvNormActBlock(n_dims=len(image_size), in_chans=in_chans, out_chans=out_chans, norm=norm),
                  ConvResBlock(n_dims=len(image_size), in_chans=in_chans, out_chans=out_chans, dropout=0.1, norm=norm)):
        for flag in (True, False):
            block.set_grad_ckpt(flag)
            assert all(m.grad_ckpt == flag for m in block.children() if hasattr(m, "grad_ckpt"))
        assert block(x).shape == (2, out_chans, *image_size)
    with pytest.raises(ValueError):
        ConvResBlock(n_dims=4, in_chans=in_chans, out_chans=out_chans, norm=norm)


@pytest.mark.parametrize("norm", ["instance", "layer", "group"])
@pytest.mark.parametrize("use_mask", [True, False])
@pytest.mark.parametrize("drop_path", [0.0, 0.1])
@pytest.mark.parametrize("image_size", SIZES)
def test_masked_conv_block(image_size, drop_path, use_mask, norm):  # cinema/conv_test.py:108-139
    block = MaskedConvBlock(n_dims=len(image_size), in_chans=8, dropout=0.1, drop_path=drop_path, norm=norm)
    block.set_grad_ckpt(True)
    assert all(m.grad_ckpt for m in block.children() if hasattr(m, "grad_ckpt"))
    x = torch.rand(2, 8, *image_size)
    mask = torch.rand(2, *image_size) > 0.5 if use_mask else None
    assert block(x, mask).shape == x.shape


def test_blocks_match_the_oracle_restatement():
    """Values: ConvNormActBlock / MaskedConvBlock (with mask) / ConvResBlock with layer norm equal the functional oracle
    (which is pinned to the real reference through the model-level goldens)."""
    from oracle import cinema_oracle as O

    torch.manual_seed(0)
    x = torch.rand(2, 4, 8, 6, 4)
    mask = torch.rand(2, 8, 6, 4) > 0.4
    cna = ConvNormActBlock(n_dims=3, in_chans=4, out_chans=8, norm="layer", kernel_size=(2, 2, 1), stride=(2, 2, 1), padding="valid")
    sd = {f"m.{k}": v for k, v in cna.state_dict().items()}
    torch.testing.assert_close(cna(x), O.conv_norm_act(sd, "m", x, (2, 2, 1), 1e-6), rtol=1e-5, atol=1e-6)
    mcb = MaskedConvBlock(n_dims=3, in_chans=4, norm="layer").eval()
    sd = {f"m.{k}": v for k, v in mcb.state_dict().items()}
    torch.testing.assert_close(mcb(x, mask), O.masked_conv_block(sd, "m", x, mask, 1e-6), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mcb(x, None), O.masked_conv_block(sd, "m", x, None, 1e-6), rtol=1e-5, atol=1e-6)
    for cin, cout in ((4, 4), (4, 6)):
        res = ConvResBlock(n_dims=3, in_chans=cin, out_chans=cout, norm="layer").eval()
        sd = {f"m.{k}": v for k, v in res.state_dict().items()}
        torch.testing.assert_close(res(x), O.conv_res_block(sd, "m", x, 1e-6), rtol=1e-5, atol=1e-6)
