"""The reference's own ``cinema/vit_test.py`` cases on the cinema_b200 building blocks (CPU, emulated kernels): patchify /
unpatchify up to 4-D with the expected token shapes, PatchEmbed, sincos tables, Attention / Block with separate query and
key lengths, the grad-ckpt switch -- plus numerics of the ``qkv_bias=False`` form (the reference default for a bare
``Attention``) against the oracle."""

import math

import pytest
import torch
from torch import nn

from cinema_b200 import vit as V


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize(("image_size", "patch_size", "in_chans", "expected"), [
    ((16, 16), (2, 4), 1, (32, 8)), ((16, 16), (2, 4), 3, (32, 24)), ((8, 12, 16), (2, 4, 8), 1, (24, 64)),
    ((8, 12, 16), (2, 4, 8), 3, (24, 192)), ((8, 12, 16, 9), (2, 4, 8, 3), 1, (72, 192)), ((8, 12, 16, 9), (2, 4, 8, 3), 3, (72, 576)),
])
def test_patchify_and_unpatchify(image_size, patch_size, in_chans, expected, emulated_kernels):  # cinema/vit_test.py:24-52
    image = torch.rand(2, in_chans, *image_size)
    x = V.patchify(image, patch_size)
    assert x.shape == (2, *expected)
    grid = tuple(s // p for s, p in zip(image_size, patch_size))
    recon = V.unpatchify(x, patch_size, grid)
    assert recon.shape == image.shape and torch.equal(recon, image)  # a pure permutation: bit-exact round trip
    with pytest.raises(ValueError):
        V.unpatchify(x, patch_size, tuple(g + 1 for g in grid))


@pytest.mark.parametrize(("image_size", "patch_size", "in_chans", "embed_dim", "expected"), [
    ((16, 16), (4, 4), 1, 4, (16, 4)), ((16, 16), (4, 2), 3, 4, (32, 4)), ((16, 16, 16), (4, 4, 4), 1, 4, (64, 4)),
    ((16, 16, 16), (4, 2, 4), 3, 4, (128, 4)), ((16, 16, 16, 9), (4, 4, 4, 3), 1, 5, (192, 5)), ((16, 16, 16, 9), (4, 2, 4, 3), 3, 7, (384, 7)),
])
@pytest.mark.parametrize("grad_ckpt", [True, False])
def test_patch_embed(image_size, patch_size, in_chans, embed_dim, expected, grad_ckpt, emulated_kernels):  # cinema/vit_test.py:55-95
    pe = V.PatchEmbed(image_size=image_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
    pe.set_grad_ckpt(grad_ckpt)
    assert pe.grad_ckpt == grad_ckpt
    out = pe(torch.rand(2, in_chans, *image_size))
    assert out.shape == (2, *expected)
    out.sum().backward()
    assert pe.proj.weight.grad is not None


@pytest.mark.parametrize("grid_size", [(3, 4), (2, 3, 4)])
def test_get_nd_sincos_pos_embed(grid_size):  # cinema/vit_test.py:110-121
    assert V.get_nd_sincos_pos_embed(16, grid_size).shape == (math.prod(grid_size), 16)
    p = V.get_pos_embed(16, grid_size)
    assert p.shape == (1, math.prod(grid_size), 16) and not p.requires_grad


def _block(qkv_bias=False, dim=16, heads=4):
    return V.Block(dim, heads, mlp_ratio=4, qkv_bias=qkv_bias, rotary=False, norm_layer=nn.LayerNorm, norm_eps=1e-5, drop_path=0.0,
                   act_layer=nn.GELU, mlp_layer=V.Mlp)


@pytest.mark.parametrize("n_q_tokens", [1, 5, 16])
@pytest.mark.parametrize("n_k_tokens", [1, 5, 16])
def test_attention_and_block_qk(n_q_tokens, n_k_tokens, emulated_kernels):  # cinema/vit_test.py:145-158,196-204
    q, k = torch.rand(2, n_q_tokens, 16), torch.rand(2, n_k_tokens, 16)
    torch.manual_seed(0)
    assert V.Attention(16)(q, k).shape == q.shape  # defaults: 8 heads, no qkv bias
    assert V.Attention(16)(q).shape == q.shape
    assert _block()(q, k).shape == q.shape


@pytest.mark.parametrize("grad_ckpt", [True, False])
def test_block_grad_ckpt_switch(grad_ckpt, emulated_kernels):  # cinema/vit_test.py:206-219
    blk = _block()
    blk.set_grad_ckpt(grad_ckpt)
    assert blk.grad_ckpt == grad_ckpt
    q = torch.rand(2, 5, 16)
    assert blk(q, torch.rand(2, 5, 16)).shape == q.shape
    with pytest.raises(ValueError):
        V.Attention(16, rotary=True)(q, torch.rand(2, 3, 16))  # rotary with q != k (cinema/vit.py:494-495)


@pytest.mark.parametrize(("n_q", "n_k"), [(5, 5), (16, 5), (3, 16)])
def test_block_without_qkv_bias_matches_oracle(n_q, n_k, emulated_kernels):
    from oracle import cinema_oracle as O

    torch.manual_seed(3)
    blk = _block(qkv_bias=False, dim=32)
    assert blk.attn.q.bias is None and blk.attn.kv.bias is None
    sd = {f"b.{k}": v.detach().clone().requires_grad_() for k, v in blk.state_dict().items()}
    named = dict(blk.named_parameters())
    for cross in (True, False):
        blk.zero_grad()
        for t in sd.values():
            t.grad = None
        q = torch.randn(2, n_q, 32, requires_grad=True)
        k = torch.randn(2, n_k, 32, requires_grad=True) if cross else None
        y = blk(q, k)
        w = torch.randn(y.shape)
        (y * w).sum().backward()
        qr = q.detach().clone().requires_grad_()
        kr = k.detach().clone().requires_grad_() if cross else None
        yr = O.block(sd, "b", qr, kr, 4, 1e-5)
        (yr * w).sum().backward()
        assert rel(y, yr) < 1e-2 and rel(q.grad, qr.grad) < 2e-2
        if cross:
            assert rel(k.grad, kr.grad) < 2e-2
        for n, p in named.items():
            assert rel(p.grad, sd[f"b.{n}"].grad) < 3e-2, (cross, n)
