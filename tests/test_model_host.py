"""Host-logic tests (no GPU): the cinema_b200 modules driven through the CPU emulation of the C-ABI
(tests/emu_c.py) must reproduce the golden vectors generated from the real reference
(tests/golden/make_golden.py).  This pins the layout bookkeeping, the index arithmetic, the
hand-written backward chains and the parameter arena; the kernels themselves are checked on the
GPU (tests/test_*_gpu.py).

Tolerances: the emulation rounds to bf16 where the kernels do, the golden vectors are fp32, so
outputs / gradients agree to a few bf16 ulps accumulated over the network depth -- 3e-2 relative
(norm-wise) is asserted, 1e-3 relative on the loss.  Metrics computed in fp32 (target mean / std) must
match to 1e-5; masks bit-exactly.
"""

import math

import pytest
import torch

from cinema_b200 import CineMA
from cinema_b200 import vit as bvit


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


CASES = ["mae_tiny_sax", "mae_small_4view", "mae_tiny_selfattn_normtarget", "mae_hd_selfattn_normtarget"]


@pytest.mark.parametrize("case", CASES)
def test_mae_forward_backward_matches_reference_golden(case, golden_dir, emulated_kernels):
    g = torch.load(golden_dir / f"{case}.pt")
    model = CineMA(**g["kw"])
    missing = model.load_state_dict(g["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    loss, preds, masks, metrics = model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
    assert loss.dim() == 0
    assert abs(float(loss) - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    for v, p in preds.items():
        assert p.shape == g["preds"][v].shape
        assert torch.equal(masks[v], g["masks"][v])
        assert rel(p, g["preds"][v]) < 3e-2
    assert set(metrics) == set(g["metrics"])
    for k, val in g["metrics"].items():
        assert metrics[k].dim() == 0
        tol = 1e-5 if ("target_mean" in k or "target_std" in k) else 2e-2
        assert abs(float(metrics[k]) - float(val)) <= tol * max(1.0, abs(float(val))), k
    loss.backward()
    named = dict(model.named_parameters())
    for k, ref in g["grads"].items():
        assert named[k].grad is not None, k
        assert rel(named[k].grad, ref) < 3e-2, k
    # every trainable parameter received a gradient view of the flat arena; fixed tables did not
    for k, p in named.items():
        assert (p.grad is not None) == p.requires_grad, k


@pytest.mark.parametrize("case", ["mae_small_4view"])
def test_dense_stem_path_matches_too(case, golden_dir, emulated_kernels):
    """``native_stem=False`` keeps the reference evaluation order (dense conv stem, then gather); same results."""
    g = torch.load(golden_dir / f"{case}.pt")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.native_stem = False
    model.train()
    loss, preds, _, _ = model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
    assert model._dense_levels == [2, 2, 2, 2]
    assert abs(float(loss) - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    loss.backward()
    named = dict(model.named_parameters())
    for k, ref in g["grads"].items():
        assert rel(named[k].grad, ref) < 3e-2, k
    model.native_stem = True
    model.zero_grad()
    model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
    assert model._dense_levels == [0, 0, 0, 0]


@pytest.mark.parametrize("case", ["mae_tiny_sax", "mae_small_4view"])
def test_feature_forward_matches_reference_golden(case, golden_dir, emulated_kernels):
    g = torch.load(golden_dir / f"{case}.pt")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.eval()
    feats = model.feature_forward(g["images"])
    assert set(feats) == set(g["feats"])
    for k, ref in g["feats"].items():
        assert feats[k].shape == ref.shape
        assert rel(feats[k], ref) < 2e-2, k


def test_state_dict_schema_and_arena_roundtrip(golden_dir, emulated_kernels):
    """Keys / shapes equal the reference's; re-homing the parameters into the flat arena keeps them loadable."""
    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"])
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["state_dict"].keys())
    for k, v in g["state_dict"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    model.load_state_dict(g["state_dict"])
    model(g["images"], 0.75)  # builds the arena
    from cinema_b200.arena import ensure_arena

    arena = ensure_arena(model)
    assert arena.valid()
    for k, v in model.state_dict().items():
        assert torch.equal(v, g["state_dict"][k]), k
    # q / kv weights of an encoder block and the kv weights of all decoder blocks are adjacent in the arena
    blk = model.encoder.blocks[0].attn
    assert arena.adjacent([blk.q.weight, blk.kv.weight])
    assert arena.adjacent([b.attn.kv.weight for b in model.decoder.blocks])
    # loading again (in place) keeps the aliasing
    model.load_state_dict(g["state_dict"])
    assert arena.valid()


def test_grad_accumulation_and_zero_grad(golden_dir, emulated_kernels):
    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.train()
    loss, *_ = model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
    loss.backward()
    p = model.encoder.blocks[0].mlp.fc1.weight
    g1 = p.grad.clone()
    loss, *_ = model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
    loss.backward()
    assert rel(p.grad, 2 * g1) < 1e-5  # .grad accumulates like autograd
    model.zero_grad(set_to_none=True)
    assert p.grad is None
    loss, *_ = model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
    loss.backward()
    assert rel(p.grad, g1) < 1e-5


def test_mask_counts_and_errors(emulated_kernels):
    from cinema_b200.mae import get_batch_random_patch_mask

    m = get_batch_random_patch_mask(4, 37, 0.6, torch.device("cpu"))
    assert m.shape == (4, 37) and m.dtype == torch.bool
    assert ((~m).sum(1) == int(37 * (1 - 0.6))).all()  # cinema/mae/mae_test.py:30-32
    assert not get_batch_random_patch_mask(2, 9, 0.0, torch.device("cpu")).any()
    with pytest.raises(ValueError):
        get_batch_random_patch_mask(2, 9, -0.1, torch.device("cpu"))
    with pytest.raises(ValueError):
        bvit.patchify(torch.zeros(1, 1, 9, 8), (2, 2))
    with pytest.raises(ValueError):
        bvit.Attention(30, n_heads=4)


def test_vit_encoder_decoder_modules(golden_dir, emulated_kernels):
    """Standalone ViTEncoder / ViTDecoder / Block / Attention against the oracle on the golden weights."""
    from oracle import cinema_oracle as O

    g = torch.load(golden_dir / "mae_small_4view.pt")
    kw = g["kw"]
    model = CineMA(**kw)
    model.load_state_dict(g["state_dict"])
    sd = g["state_dict"]
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 11, kw["enc_embed_dim"], generator=gen).requires_grad_()
    y = model.encoder(x)
    xr = x.detach().clone().requires_grad_()
    yr = O.vit_encoder(sd, "encoder", xr, kw["enc_depth"], kw["enc_n_heads"], 1e-5)
    assert y.shape == yr.shape and rel(y, yr) < 2e-2
    w = torch.randn(y.shape, generator=gen)
    (y * w).sum().backward()
    (yr * w).sum().backward()
    assert rel(x.grad, xr.grad) < 3e-2
    # cross-attention decoder
    dd = kw["dec_embed_dim"]
    xq = torch.randn(2, 9, dd, generator=gen).requires_grad_()
    xk = torch.randn(2, 5, dd, generator=gen).requires_grad_()
    out = model.decoder(xq, xk, 8)
    xq_r, xk_r = xq.detach().clone().requires_grad_(), xk.detach().clone().requires_grad_()
    out_r = O.vit_decoder(sd, "decoder", xq_r, xk_r, 8, kw["dec_depth"], kw["dec_n_heads"], 1e-5)
    assert out.shape == out_r.shape == (2, 8, dd) and rel(out, out_r) < 2e-2
    w = torch.randn(out.shape, generator=gen)
    (out * w).sum().backward()
    (out_r * w).sum().backward()
    assert rel(xq.grad, xq_r.grad) < 3e-2 and rel(xk.grad, xk_r.grad) < 3e-2
    # a single block and a bare attention layer
    blk = model.decoder.blocks[0]
    assert rel(blk(xq.detach(), xk.detach()), O.block(sd, "decoder.blocks.0", xq.detach(), xk.detach(),
                                                       kw["dec_n_heads"], 1e-5)) < 2e-2
    att = model.encoder.blocks[1].attn
    xa = torch.randn(2, 7, kw["enc_embed_dim"], generator=gen)
    assert rel(att(xa), O.attention(sd, "encoder.blocks.1.attn", xa, None, kw["enc_n_heads"])) < 2e-2


def test_direct_train_step_equals_autograd_step(golden_dir, emulated_kernels):
    """``CineMA.train_step`` (forward + hand-written backward called back to back, no autograd engine) leaves the same
    loss and the same gradients as ``forward(...)[0].backward()``."""
    g = torch.load(golden_dir / "mae_small_4view.pt")
    grads = {}
    for mode in ("autograd", "direct"):
        model = CineMA(**g["kw"])
        model.load_state_dict(g["state_dict"])
        model.train()
        assert model.direct_step_supported()
        if mode == "autograd":
            loss, *_ = model(g["images"], g["ratio"], enc_mask_dict=g["masks"])
            loss.backward()
        else:
            loss = model.train_step(g["images"], g["ratio"], enc_mask_dict=g["masks"])
            assert not loss.requires_grad
        grads[mode] = (float(loss), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert grads["autograd"][0] == grads["direct"][0]
    assert set(grads["autograd"][1]) == set(grads["direct"][1])
    for k, ref in grads["autograd"][1].items():
        assert torch.equal(grads["direct"][1][k], ref), k
    model.native_stem = False
    assert not model.direct_step_supported()
    with pytest.raises(RuntimeError):
        model.train_step(g["images"], g["ratio"])


@pytest.mark.parametrize(("name", "views", "ratio", "batch"), [
    ("subset_of_views", ["lax_3c", "sax"], 0.75, 2),  # any subset, any order (cinema/mae/mae.py:521-523)
    ("batch_of_one", ["sax", "lax_2c", "lax_3c", "lax_4c"], 0.75, 1),
    ("single_2d_view", ["lax_4c"], 0.5, 2),
    # tiny grids at a high ratio: int(16 * 0.05) = 0 visible LAX tokens, 1 visible SAX token -- the reference's own tests
    # reach this corner (cinema/mae/mae_test.py:35-66, ratio 0.9 on 2..8-token grids)
    ("view_without_visible_tokens", ["sax", "lax_2c"], 0.95, 2),
])
def test_mae_edge_cases_against_oracle(name, views, ratio, batch, golden_dir, emulated_kernels):
    from oracle import cinema_oracle as O

    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.train()
    images = {v: g["images"][v][:batch] for v in views}
    torch.manual_seed(7)
    loss, preds, masks, metrics = model(images, ratio)
    loss.backward()
    sd = {k: v.clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    ref_loss, ref_preds, _ = O.mae_forward(sd, O.MAEConfig(**g["kw"]), images, masks)
    ref_loss.backward()
    assert list(preds) == views
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss))
    for v in views:
        n = masks[v].shape[1]
        n_keep = int(n * (1 - ratio))
        assert ((~masks[v]).sum(1) == n_keep).all()
        assert preds[v].shape == ref_preds[v].shape == (batch, n - n_keep, 256)
        assert rel(preds[v], ref_preds[v]) < 3e-2, v
    named = dict(model.named_parameters())
    checked = 0
    for k, p in sd.items():
        if p.grad is None or float(p.grad.norm()) < 1e-6:  # unused views / analytically zero gradients (one-key softmax)
            assert named[k].grad is None or float(named[k].grad.norm()) < 1e-6 or not k.startswith(("enc_", "dec_embed", "pred_head")), k
            continue
        assert rel(named[k].grad, p.grad) < 4e-2, (k, rel(named[k].grad, p.grad))
        checked += 1
    assert checked > 50


def test_mask_ratio_zero_and_no_keys(golden_dir, emulated_kernels):
    """enc_mask_ratio = 0: nothing to reconstruct -- empty predictions and a NaN loss, like the reference
    (cinema/mae/mae.py:604-608).  No visible token in ANY view (int(16 * 0.01) = 0): the cross-attention decoder attends
    over zero keys, which contributes 0 (SDPA semantics) -- loss, predictions and gradients still match the oracle."""
    from oracle import cinema_oracle as O

    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    with torch.no_grad():
        loss, preds, masks, metrics = model(g["images"], 0.0)
    assert math.isnan(float(loss)) and all(p.shape[1] == 0 for p in preds.values()) and not any(m.any() for m in masks.values())
    assert all(math.isfinite(float(metrics[f"{v}_target_mean"])) for v in preds)
    model.train()
    images = {v: g["images"][v] for v in ("lax_2c", "lax_3c")}
    loss, preds, masks, _ = model(images, 0.99)
    assert all(bool(m.all()) for m in masks.values())
    loss.backward()
    sd = {k: v.clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    ref_loss, ref_preds, _ = O.mae_forward(sd, O.MAEConfig(**g["kw"]), images, masks)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss))
    named = dict(model.named_parameters())
    for v in images:
        assert rel(preds[v], ref_preds[v]) < 2e-2
    for k, p in sd.items():
        if p.grad is not None and float(p.grad.norm()) > 1e-6:
            assert rel(named[k].grad, p.grad) < 3e-2, k


# the reference's own parametrisations (cinema/mae/mae_test.py:35-66, cinema/convvit_test.py:72-90,150-185): toy channel counts
# (2, 4: below the kernels' multiple-of-8 requirement, fine for the emulation), 3-level stems, anisotropic patches and scale
# factors, multi-channel / multi-frame inputs, inputs smaller / larger than the configured size, views left without tokens
_MAE_GRID = [
    ({"img1": (32, 32)}, {"img1": (2, 4)}, {"img1": (2, 2)}, [4, 8], {"img1": 1}),
    ({"img1": (32, 64)}, {"img1": (2, 4)}, {"img1": (2, 2)}, [2, 4, 8], {"img1": 1}),
    ({"img1": (32, 32), "img2": (32, 16)}, {"img1": (2, 4), "img2": (2, 2)}, {"img1": (2, 2), "img2": (2, 2)}, [4, 8], {"img1": 1, "img2": 3}),
    ({"img1": (8, 32, 32), "img2": (32, 16)}, {"img1": (2, 4, 1), "img2": (2, 4)}, {"img1": (2, 2, 1), "img2": (2, 2)}, [4, 8],
     {"img1": 1, "img2": 3}),
    ({"img1": (8, 16, 16), "img2": (16, 16), "img3": (16, 8)}, {"img1": (2, 4, 8), "img2": (2, 4), "img3": (2, 4)},
     {"img1": (2, 2, 1), "img2": (2, 2), "img3": (2, 1)}, [4, 8], {"img1": 1, "img2": 3, "img3": 2}),
]


@pytest.mark.parametrize("cross_attn", [True, False])
@pytest.mark.parametrize("enc_mask_ratio", [0.1, 0.5, 0.9])
@pytest.mark.parametrize("grid", range(len(_MAE_GRID)))
def test_reference_test_grid_shapes_mae(grid, enc_mask_ratio, cross_attn, emulated_kernels):
    """The assertions of the reference's ``TestCineMA`` (cinema/mae/mae_test.py:68-131) on its own parameter grid."""
    isd, psd, sfd, chans, icd = _MAE_GRID[grid]
    torch.manual_seed(0)
    mae = CineMA(image_size_dict=isd, in_chans_dict=icd, enc_patch_size_dict=psd, enc_scale_factor_dict=sfd, enc_conv_chans=chans,
                 enc_conv_n_blocks=1, enc_embed_dim=16, enc_depth=1, enc_n_heads=2, dec_embed_dim=8, dec_depth=1, dec_n_heads=2,
                 mlp_ratio=2, norm_target=bool(grid % 2), cross_attn=cross_attn)
    mae.set_grad_ckpt(True)
    assert all(m.grad_ckpt for m in mae.children() if hasattr(m, "grad_ckpt"))
    image_dict = {k: torch.rand(2, icd[k], *isd[k]) for k in isd}
    loss, pred_dict, mask_dict, metrics = mae(image_dict, enc_mask_ratio)
    ns_patches = [mae.enc_down_dict[k].patch_embed.n_patches for k in isd]
    assert sum(pred_dict[k].shape[1] for k in isd) == sum(n - int(n * (1 - enc_mask_ratio)) for n in ns_patches)
    for k in isd:
        assert pred_dict[k].shape[0] == 2 and pred_dict[k].shape[2] == math.prod(mae.dec_patch_size_dict[k]) * icd[k]
        assert mask_dict[k].shape == (2, mae.enc_down_dict[k].patch_embed.n_patches)
    assert loss.ndim == 0 and all(v.ndim == 0 for v in metrics.values())
    if torch.isfinite(loss):
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in mae.parameters() if p.grad is not None)


def test_mae_reconstruction_example(golden_dir, emulated_kernels):
    """cinema_b200/examples/inference.py (cinema/examples/inference/mae.py:57-83): visible patches keep the input voxels bit
    for bit, masked patches hold the model's predictions, the mask image marks exactly the masked patches."""
    from cinema_b200.examples.inference import mae_reconstruct, reconstruct_images

    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    torch.manual_seed(0)
    loss, recon, masks = mae_reconstruct(model, g["images"], 0.75)
    assert torch.isfinite(loss) and set(recon) == set(g["images"])
    for v, img in g["images"].items():
        assert recon[v].shape == img.shape and masks[v].shape == img.shape
        vis = masks[v] == 0
        assert torch.equal(recon[v][vis], img[vis])
        frac = float(masks[v].mean())
        n = model.enc_down_dict[v].patch_embed.n_patches
        assert abs(frac - (n - int(n * 0.25)) / n) < 1e-6
    # with the golden predictions / masks the pasted patches are exactly the reference's predictions
    grids = {v: model.enc_down_dict[v].patch_embed.grid_size for v in g["preds"]}
    recon, masks = reconstruct_images(g["images"], g["preds"], g["masks"], model.dec_patch_size_dict, grids)
    for v in g["preds"]:
        back = bvit.patchify(recon[v], model.dec_patch_size_dict[v])
        assert torch.equal(back[g["masks"][v]], g["preds"][v].reshape(-1, g["preds"][v].shape[-1]))
