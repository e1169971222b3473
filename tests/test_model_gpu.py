"""GPU parity tests of the module-level path (through the C-ABI) against the oracle and the committed
golden vectors generated from the real reference.

Tolerances.  The golden vectors are fp32 (the reference disables autocast on CPU).  The B200 path computes
GEMMs / attention in bf16 with fp32 accumulation exactly where ``torch.autocast(bf16)`` would, so the yardstick
is the reference's OWN bf16 noise: the oracle is run under CUDA bf16 autocast (the stock torch path) and our error
against the fp32 golden must not exceed 1.5 x that path's error (+ a small absolute floor).  The loss itself
must agree with the fp32 value to 1e-3 relative (BASELINE.json north_star); index work is bit-exact.
"""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cinema_b200 import CineMA, _C
    from cinema_b200 import vit as bvit
    from oracle import cinema_oracle as O

DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _to(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def _oracle_autocast(g):
    """the stock torch path: oracle modules math under CUDA bf16 autocast, fp32 params on the GPU."""
    cfg = O.MAEConfig(**g["kw"])
    params = {k: v.to(DEV).clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, preds, _ = O.mae_forward(params, cfg, _to(g["images"], DEV), _to(g["masks"], DEV))
    loss.backward()
    return loss.detach(), {k: v.detach().float() for k, v in preds.items()}, {k: p.grad for k, p in params.items() if p.grad is not None}


@pytest.mark.parametrize("native_stem", [True, False])
@pytest.mark.parametrize("case", ["mae_small_4view", "mae_hd_selfattn_normtarget"])
def test_mae_step_parity(case, native_stem, golden_dir):
    g = torch.load(golden_dir / f"{case}.pt")
    model = CineMA(**g["kw"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.native_stem = native_stem
    model.train()
    loss, preds, masks, metrics = model(_to(g["images"], DEV), g["ratio"], enc_mask_dict=_to(g["masks"], DEV))
    loss.backward()
    torch.cuda.synchronize()
    ref_loss, ref_preds, ref_grads = _oracle_autocast(g)
    # loss: 1e-3 relative to the fp32 reference value
    assert abs(float(loss) - float(g["loss"])) <= 1e-3 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    for v in preds:
        assert torch.equal(masks[v].cpu(), g["masks"][v])
        ours, stock = rel(preds[v], g["preds"][v]), rel(ref_preds[v], g["preds"][v])
        assert ours <= 1.5 * stock + 2e-3, (v, ours, stock)
    for k, val in g["metrics"].items():
        tol = 1e-5 if ("target_mean" in k or "target_std" in k) else 2e-2
        assert abs(float(metrics[k]) - float(val)) <= tol * max(1.0, abs(float(val))), k
    named = dict(model.named_parameters())
    for k, ref in g["grads"].items():
        ours, stock = rel(named[k].grad, ref), rel(ref_grads[k], ref)
        assert ours <= 1.5 * stock + 5e-3, (k, ours, stock)


def test_feature_forward_parity(golden_dir):
    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.eval()
    feats = model.feature_forward(_to(g["images"], DEV))
    for k, ref in g["feats"].items():
        assert feats[k].shape == ref.shape
        assert rel(feats[k], ref) < 3e-2, k


def test_vit_modules_against_oracle(golden_dir):
    g = torch.load(golden_dir / "mae_small_4view.pt")
    kw = g["kw"]
    model = CineMA(**kw).to(DEV)
    model.load_state_dict(g["state_dict"])
    sd = g["state_dict"]
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 300, kw["enc_embed_dim"], generator=gen)
    xr = x.clone().requires_grad_()
    xg = x.to(DEV).requires_grad_()
    y = model.encoder(xg)
    yr = O.vit_encoder(sd, "encoder", xr, kw["enc_depth"], kw["enc_n_heads"], 1e-5)
    assert rel(y, yr) < 2e-2
    w = torch.randn(yr.shape, generator=gen)
    (y * w.to(DEV)).sum().backward()
    (yr * w).sum().backward()
    assert rel(xg.grad, xr.grad) < 3e-2
    dd = kw["dec_embed_dim"]
    xq, xk = torch.randn(2, 261, dd, generator=gen), torch.randn(2, 140, dd, generator=gen)
    xq_r, xk_r = xq.clone().requires_grad_(), xk.clone().requires_grad_()
    xq_g, xk_g = xq.to(DEV).requires_grad_(), xk.to(DEV).requires_grad_()
    out = model.decoder(xq_g, xk_g, 260)
    out_r = O.vit_decoder(sd, "decoder", xq_r, xk_r, 260, kw["dec_depth"], kw["dec_n_heads"], 1e-5)
    assert out.shape == out_r.shape and rel(out, out_r) < 2e-2
    w = torch.randn(out_r.shape, generator=gen)
    (out * w.to(DEV)).sum().backward()
    (out_r * w).sum().backward()
    assert rel(xq_g.grad, xq_r.grad) < 3e-2 and rel(xk_g.grad, xk_r.grad) < 3e-2


def test_attention_module_against_reference_golden(golden_dir):
    ops = torch.load(golden_dir / "ops.pt")
    ac = ops["attn_cross"]  # dim 64, 2 heads -> head_dim 32, q != k, generated by the reference Attention
    att = bvit.Attention(64, n_heads=2, qkv_bias=True).to(DEV)
    att.load_state_dict(ac["sd"])
    y = att(ac["q"].to(DEV), ac["k"].to(DEV))
    assert y.shape == ac["y"].shape and rel(y, ac["y"]) < 1e-2


def test_public_patchify_bit_exact(golden_dir):
    ops = torch.load(golden_dir / "ops.pt")
    for nm in ("p2", "p3", "p4"):
        o = ops[nm]
        tok = bvit.patchify(o["image"].to(DEV), o["patch_size"])
        assert torch.equal(tok.cpu(), o["tokens"])
        assert torch.equal(bvit.unpatchify(tok, o["patch_size"], o["grid"]).cpu(), o["image"])
    with pytest.raises(ValueError):
        bvit.patchify(torch.zeros(1, 1, 9, 8, device=DEV), (2, 2))


def test_rotary_kernel_against_reference_golden(golden_dir):
    from cinema_b200.rotary import RotaryEmbedding

    r = torch.load(golden_dir / "ops.pt")["rotary"]
    rot = RotaryEmbedding(r["dim"])
    q = r["q"].to(DEV).requires_grad_()
    rq, rk = rot(q, r["k"].to(DEV))
    torch.testing.assert_close(rq.cpu(), r["rq"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rk.cpu(), r["rk"], rtol=1e-5, atol=1e-6)
    # backward = inverse rotation: <R q, w> differentiates to R^T w
    w = torch.randn_like(rq)
    (rq * w).sum().backward()
    qr = r["q"].clone().requires_grad_()
    c, s = O.rotary_tables(qr.shape[1], r["dim"], torch.float32)
    (O.apply_rotary(qr, c, s) * w.cpu()).sum().backward()
    torch.testing.assert_close(q.grad.cpu(), qr.grad, rtol=1e-5, atol=1e-6)
    # bf16 activations: tables are rounded to bf16 like the reference does
    rq16, _ = rot(r["q"].to(DEV).bfloat16(), r["k"].to(DEV).bfloat16())
    assert rq16.dtype == torch.bfloat16 and rel(rq16.float(), r["rq"]) < 1e-2


def test_fused_adamw_matches_torch():
    torch.manual_seed(0)
    n = 1 << 20
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 3
    ref = p.clone().requires_grad_()
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    p16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    for t in range(1, 4):
        ref.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_([ref], 5.0)
        opt.step()
        gn = torch.zeros(1, device=DEV)
        _C.sumsq(g, gn)
        torch.testing.assert_close(gn.sqrt()[0], norm, rtol=1e-4, atol=0)
        hyper = torch.tensor([1e-3, 1 - 0.9 ** t, 1 - 0.95 ** t, 0], device=DEV)
        _C.adamw_flat(p, g, m, v, p16, hyper, 0.9, 0.95, 1e-8, 0.05, gn, 5.0, 1.0)
        torch.testing.assert_close(p, ref.detach(), rtol=1e-5, atol=1e-6)
        assert torch.equal(p16, p.to(torch.bfloat16))
    # non-finite gradient norm skips the step (GradScaler semantics)
    before = p.clone()
    _C.adamw_flat(p, g, m, v, p16, hyper, 0.9, 0.95, 1e-8, 0.05, torch.tensor([float("inf")], device=DEV), 5.0, 1.0)
    assert torch.equal(p, before)


def test_trainer_graph_equals_eager_and_learns(golden_dir):
    from cinema_b200.train import MAETrainer

    g = torch.load(golden_dir / "mae_small_4view.pt")
    images = _to(g["images"], DEV)
    losses = {}
    for use_graph in (False, True):
        torch.manual_seed(0)
        model = CineMA(**g["kw"]).to(DEV)
        model.load_state_dict(g["state_dict"])
        model.train()
        tr = MAETrainer(model, lr=1e-3, use_cuda_graph=use_graph, graph_warmup=2)
        torch.manual_seed(1)
        losses[use_graph] = [float(tr.step(images)) for _ in range(8)]
    assert all(math.isfinite(x) for x in losses[True])
    assert losses[True][-1] < losses[True][0]  # the fixed batch is being fit
    # the replayed graph runs the same kernels as the eager step; masks come from the same RNG stream
    # only up to the capture point, so compare the eager prefix exactly and the level afterwards
    assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(losses[True][:2], losses[False][:2]))  # atomics reorder sums
    assert abs(losses[True][-1] - losses[False][-1]) < 0.1 * losses[False][0]


def test_full_size_vitb_step_properties():
    """BASELINE.json config: ViT-B, SAX 192x192x16 + 3 LAX 192x192, mask 0.75.  Too big for the CPU oracle in a test,
    so size-independent properties are checked: exact keep counts, finite loss, loss equals the mean of the per-view
    MSEs, predictions only depend on the sample (batch independence), and a deterministic repeat."""
    from bench import model_kwargs

    kw = model_kwargs("base", (192, 192, 16), (192, 192))
    torch.manual_seed(0)
    model = CineMA(**kw).to(DEV).train()
    gen = torch.Generator().manual_seed(3)
    images = {v: torch.rand(2, 1, *s, generator=gen).to(DEV) for v, s in kw["image_size_dict"].items()}
    torch.manual_seed(5)
    loss, preds, masks, metrics = model(images, 0.75)
    loss.backward()
    assert math.isfinite(float(loss))
    for v, m in masks.items():
        n = m.shape[1]
        assert ((~m).sum(1) == int(n * 0.25)).all()
        assert preds[v].shape == (2, n - int(n * 0.25), 256)
    mean_mse = sum(float(metrics[f"{v}_mse_loss"]) for v in masks) / len(masks)
    assert abs(float(loss) - mean_mse) < 1e-5
    # pixel-level check of the fused loss against plain torch on the returned predictions
    for v in masks:
        tgt = O.patchify(images[v], model.dec_patch_size_dict[v])
        ref = ((preds[v] - tgt[masks[v]].reshape(preds[v].shape)) ** 2).mean()
        assert abs(float(metrics[f"{v}_mse_loss"]) - float(ref)) <= 1e-5 * float(ref)
    # sample 0 alone, with its own mask, gives the same predictions for sample 0
    loss1, preds1, _, _ = model({v: t[:1] for v, t in images.items()}, 0.75, enc_mask_dict={v: m[:1] for v, m in masks.items()})
    for v in masks:
        assert rel(preds1[v][0], preds[v][0]) < 2e-2
    g0 = model.encoder.blocks[0].mlp.fc1.weight.grad.clone()
    model.zero_grad(set_to_none=True)
    loss2, _, _, _ = model(images, 0.75, enc_mask_dict=masks)
    loss2.backward()
    assert abs(float(loss2) - float(loss)) < 1e-4 * abs(float(loss))
    assert rel(model.encoder.blocks[0].mlp.fc1.weight.grad, g0) < 1e-2  # atomics reorder fp32 sums only


def _parity_log(name, record):
    """Measured ours-vs-stock error ratios are appended to gpurun_out/parity_r02.json (copied to profiles/ by hand)."""
    import json
    from pathlib import Path

    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    out.mkdir(exist_ok=True)
    f = out / "parity_r02.json"
    data = json.loads(f.read_text()) if f.exists() else {}
    data[name] = record
    f.write_text(json.dumps(data, indent=1, sort_keys=True))


FULL_SIZE_CONFIGS = {
    # BASELINE.json configs[1]: SAX 192x192x16 alone, ViT-B (577 / 1729 / 576 tokens)
    "cfg2_sax_only_vitb": ("base", None),
    # configs[2] as BASELINE.json words it: 4 views, LAX 192x192 (685 / 2053 / 684), the bench workload
    "cfg3_vitb_lax192": ("base", 192),
    # configs[2] with the reference's default LAX 256x256 (cinema/__init__.py:10; 769 / 2305 / 768)
    "cfg3_vitb_lax256": ("base", 256),
    # configs[4]: ViT-L encoder (cinema/vit.py:807-822), 4 views, LAX 256x256
    "cfg5_vitl_lax256": ("large", 256),
}
FULL_SIZE_GRADS = ("encoder.blocks.0.attn.q.weight", "encoder.blocks.{last}.mlp.fc2.weight", "decoder.blocks.0.attn.kv.weight",
                   "decoder.blocks.7.mlp.fc1.bias", "dec_linear.weight", "pred_head_dict.sax.weight",
                   "enc_down_dict.sax.conv_blocks.0.conv.0.dw_conv.weight", "enc_down_dict.sax.patch_embed.proj.weight",
                   "enc_fusion_dict.sax.down_convs.0.weight", "encoder.cls_token", "encoder.norm.weight",
                   "decoder.blocks.3.norm1.bias")


@pytest.mark.parametrize("name", list(FULL_SIZE_CONFIGS))
def test_full_size_matches_stock_torch_path(name):
    """BASELINE.json configs[1], [2] (both LAX sizes) and [4] at FULL size, B = 2: our step against (i) the fp32 evaluation
    of the same model on the GPU (oracle math, TF32 off) = the truth and (ii) the stock torch path (the same math under
    CUDA bf16 autocast: SDPA, cuBLASLt, cuDNN) on identical weights, inputs and masks.  Ours and stock are two bf16
    tensor-core evaluations of one fp32 model: the loss must hit the fp32 value to 1e-3 relative (north_star), and our
    error against the truth may not exceed 1.5 x the stock path's own error (+ a small floor), for the predictions of
    every view and a spread of gradients (first / last encoder block, decoder, stem, fusion, heads, norms, cls token).
    The measured errors and their ratios are logged (profiles/r02_parity.json)."""
    from bench import model_kwargs

    size, lax = FULL_SIZE_CONFIGS[name]
    kw = model_kwargs(size, (192, 192, 16), (lax or 192, lax or 192))
    if lax is None:
        for key in ("image_size_dict", "in_chans_dict", "enc_patch_size_dict", "enc_scale_factor_dict"):
            kw[key] = {"sax": kw[key]["sax"]}
    cfg = O.MAEConfig(**kw)
    sd = O.init_state_dict(cfg, seed=1)
    model = CineMA(**kw).to(DEV).train()
    model.load_state_dict(sd)
    gen = torch.Generator().manual_seed(11)
    images = {v: torch.rand(2, 1, *s, generator=gen).to(DEV) for v, s in kw["image_size_dict"].items()}
    torch.manual_seed(7)
    masks = {v: O.random_patch_mask(2, cfg.n_patches(v), 0.75, device=DEV) for v in kw["image_size_dict"]}

    loss, preds, _, _ = model(images, 0.75, enc_mask_dict=masks)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    preds = {k: v.detach().float().clone() for k, v in preds.items()}
    del model
    g = {"kw": kw, "state_dict": sd, "images": images, "masks": masks}
    stock_loss, stock_preds, stock_grads = _oracle_autocast(g)
    # fp32 truth on the GPU: no autocast, TF32 off for the convolutions (matmul default is already full fp32)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        params = {k: v.to(DEV).clone().requires_grad_(not k.endswith("pos_embed")) for k, v in sd.items()}
        true_loss, true_preds, _ = O.mae_forward(params, cfg, images, masks)
        true_loss.backward()
        true_grads = {k: p.grad for k, p in params.items() if p.grad is not None}
        true_preds = {k: v.detach() for k, v in true_preds.items()}
    finally:
        torch.backends.cudnn.allow_tf32 = tf32

    last = kw["enc_depth"] - 1
    rec = {"tokens": [cfg.n_patches(v) for v in kw["image_size_dict"]],
           "loss": {"ours": float(loss), "stock_bf16": float(stock_loss), "fp32": float(true_loss),
                    "ours_rel": abs(float(loss) - float(true_loss)) / abs(float(true_loss)),
                    "stock_rel": abs(float(stock_loss) - float(true_loss)) / abs(float(true_loss))},
           "preds": {}, "grads": {}}
    failures = []
    for v in preds:
        ours, stock = rel(preds[v], true_preds[v]), rel(stock_preds[v], true_preds[v])
        rec["preds"][v] = {"ours": ours, "stock": stock, "ratio": ours / max(stock, 1e-12), "ours_vs_stock": rel(preds[v], stock_preds[v])}
        if not ours <= 1.5 * stock + 2e-3:
            failures.append((v, ours, stock))
    for pat in FULL_SIZE_GRADS:
        k = pat.format(last=last)
        if k not in true_grads:
            continue
        ours, stock = rel(grads[k], true_grads[k]), rel(stock_grads[k], true_grads[k])
        rec["grads"][k] = {"ours": ours, "stock": stock, "ratio": ours / max(stock, 1e-12)}
        if not ours <= 1.5 * stock + 5e-3:
            failures.append((k, ours, stock))
    _parity_log(name, rec)
    assert rec["loss"]["ours_rel"] <= 1e-3, rec["loss"]
    assert not failures, failures


def test_same_seed_same_masks_as_the_reference_modules():
    """Drop-in check against the UNMODIFIED reference classes (baseline/_ref, when the install travelled): identical
    state dict + identical torch seed -> identical masks (bit-exact index work) and the same loss to 1e-3 relative
    against the reference under bf16 autocast."""
    from baseline import stock

    if not stock.available():
        pytest.skip("baseline/_ref is not installed on this box")
    from bench import model_kwargs

    kw = model_kwargs("base", (192, 192, 16), (192, 192))
    ref = stock.build_reference_model(kw, grad_ckpt=False, seed=3).to(DEV).train()
    model = CineMA(**kw).to(DEV).train()
    model.load_state_dict(ref.state_dict())
    gen = torch.Generator().manual_seed(5)
    images = {v: torch.rand(2, 1, *s, generator=gen).to(DEV) for v, s in kw["image_size_dict"].items()}
    torch.manual_seed(9)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref_loss, ref_preds, ref_masks, _ = ref(images, 0.75)
    torch.manual_seed(9)
    loss, preds, masks, _ = model(images, 0.75)
    for v in masks:
        assert torch.equal(masks[v], ref_masks[v]), v
        assert preds[v].shape == ref_preds[v].shape
        assert rel(preds[v], ref_preds[v].float()) < 3e-2, v
    assert abs(float(loss) - float(ref_loss)) <= 2e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss))


# ------------------------------------------------------------------------------------------
# ConvViT: the classification / regression fine-tuning model on the same kernels
# ------------------------------------------------------------------------------------------
def _convvit_oracle_autocast(g, masks):
    cfg = O.convvit_config(g["kw"])
    params = {k: v.to(DEV).clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.convvit_forward(params, cfg, _to(g["images"], DEV), None if masks is None else _to(masks, DEV), "all")
    (out.float() * g["w"].to(DEV)).sum().backward()
    return out.detach().float(), {k: p.grad for k, p in params.items() if p.grad is not None}


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("case", ["convvit_2view", "convvit_frames_resized"])
def test_convvit_parity(case, masked, golden_dir):
    """Logits (all reduce modes), features and gradients against the fp32 reference golden; yardstick = the stock torch
    bf16-autocast path's own error (x 1.5 + floor), as for the MAE step."""
    from cinema_b200 import ConvViT

    g = torch.load(golden_dir / f"{case}.pt")
    model = ConvViT(**g["kw"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.train()
    images = _to(g["images"], DEV)
    masks = g["mask_dict"] if masked else None
    out = model(images, None if masks is None else _to(masks, DEV), "all")
    (out * g["w"].to(DEV)).sum().backward()
    torch.cuda.synchronize()
    ref = (g["logits_masked"] if masked else g["logits"])["all"]
    stock_out, stock_grads = _convvit_oracle_autocast(g, masks)
    scale = max(1.0, float(ref.abs().max()))
    assert float((out.cpu() - ref).abs().max()) <= 1.5 * float((stock_out.cpu() - ref).abs().max()) + 5e-3 * scale
    named = dict(model.named_parameters())
    for k, ref_g in (g["grads_masked"] if masked else g["grads"]).items():
        ours, stock = rel(named[k].grad, ref_g), rel(stock_grads[k], ref_g)
        assert ours <= 1.5 * stock + 5e-3, (k, ours, stock)
    if not masked:
        model.eval()
        with torch.no_grad():
            feats = model.feature_forward(images, None)
            for k, f in g["feats"].items():
                assert rel(feats[k], f) < 3e-2, k
            for reduce, want in g["logits"].items():
                got = model(images, None, reduce)
                assert float((got.cpu() - want).abs().max()) <= 3e-2 * max(1.0, float(want.abs().max())), reduce


def test_convvit_full_size_vitb_sax_step():
    """ViT-B ConvViT on SAX 192 x 192 x 16 (2305 tokens through the encoder, config 4's encoder shape): one training
    step runs on the native path, logits finite and equal between two identical calls, every parameter gets a gradient."""
    from cinema_b200 import ConvViT

    torch.manual_seed(0)
    model = ConvViT(image_size_dict={"sax": (192, 192, 16)}, in_chans_dict={"sax": 1}, n_frames=1, out_chans=4,
                    enc_patch_size_dict={"sax": (4, 4, 1)}, enc_scale_factor_dict={"sax": (2, 2, 1)}, enc_conv_chans=[64, 128],
                    enc_conv_n_blocks=2, enc_embed_dim=768, enc_depth=12, enc_n_heads=12).to(DEV)
    model.train()
    x = {"sax": torch.rand(2, 1, 192, 192, 16, device=DEV)}
    out = model(x)
    assert out.shape == (2, 4) and bool(torch.isfinite(out).all())
    assert model._dense_levels == [0]
    torch.nn.functional.cross_entropy(out, torch.tensor([1, 3], device=DEV)).backward()
    for k, p in model.named_parameters():
        assert (p.grad is not None) == p.requires_grad, k
        if p.grad is not None:
            assert bool(torch.isfinite(p.grad).all()), k
    assert float(model.encoder.blocks[0].attn.q.weight.grad.abs().sum()) > 0
    with torch.no_grad():
        assert torch.equal(model(x), model(x))


def test_convvit_stochastic_depth_parity(golden_dir, monkeypatch):
    """drop_path > 0, training mode (every fine-tuning config of the reference): replaying the per-sample factors the
    reference drew, logits and gradients match the fp32 golden within the stock bf16 path's own error."""
    from cinema_b200 import ConvViT, engine

    g = torch.load(golden_dir / "convvit_droppath.pt")
    model = ConvViT(**g["kw"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.train()
    queue = [t.to(DEV).float().contiguous() for t in g["drop_scales"]]
    monkeypatch.setattr(engine, "draw_drop_scales", lambda b, p, k, dev: (queue.pop(0), queue.pop(0)))
    out = model(_to(g["images"], DEV), None, "all")
    assert not queue
    (out * g["w"].to(DEV)).sum().backward()
    torch.cuda.synchronize()
    cfg = O.convvit_config(g["kw"])
    params = {k: v.to(DEV).clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        stock = O.convvit_forward(params, cfg, _to(g["images"], DEV), None, "all",
                                  drop_scales=[t.to(DEV) for t in g["drop_scales"]])
    (stock.float() * g["w"].to(DEV)).sum().backward()
    ref = g["logits"]["all"]
    scale = max(1.0, float(ref.abs().max()))
    assert float((out.cpu() - ref).abs().max()) <= 1.5 * float((stock.float().cpu() - ref).abs().max()) + 5e-3 * scale
    named = dict(model.named_parameters())
    for k, ref_g in g["grads"].items():
        ours, theirs = rel(named[k].grad, ref_g), rel(params[k].grad, ref_g)
        assert ours <= 1.5 * theirs + 5e-3, (k, ours, theirs)
    monkeypatch.undo()
    # with the real generator: reproducible under a seed, different across seeds, identity in eval mode
    torch.manual_seed(1)
    a = model(_to(g["images"], DEV), None, "all")
    torch.manual_seed(1)
    b = model(_to(g["images"], DEV), None, "all")
    torch.manual_seed(2)
    c = model(_to(g["images"], DEV), None, "all")
    assert torch.equal(a, b) and not torch.equal(a, c)


# ------------------------------------------------------------------------------------------
# ConvUNetR: segmentation fine-tuning model (native ViT encoder, cuDNN stem / decoder)
# ------------------------------------------------------------------------------------------
def test_convunetr_parity(golden_dir):
    from cinema_b200.segmentation import ConvUNetR

    g = torch.load(golden_dir / "convunetr_2view.pt")
    model = ConvUNetR(**g["kw"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.train()
    images, w = _to(g["images"], DEV), _to(g["w"], DEV)
    preds = model(images)
    sum((preds[v].float() * w[v]).sum() for v in preds).backward()
    torch.cuda.synchronize()
    cfg = O.convunetr_config(g["kw"])
    params = {k: v.to(DEV).clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        stock = O.convunetr_forward(params, cfg, images, g["n_layers_wo_skip"])
    sum((stock[v].float() * w[v]).sum() for v in stock).backward()
    for v, ref in g["preds"].items():
        assert preds[v].shape == ref.shape
        ours, theirs = rel(preds[v].float(), ref), rel(stock[v].float(), ref)
        assert ours <= 1.5 * theirs + 5e-3, (v, ours, theirs)
    named = dict(model.named_parameters())
    for k, ref_g in g["grads"].items():
        ours, theirs = rel(named[k].grad, ref_g), rel(params[k].grad, ref_g)
        assert ours <= 1.5 * theirs + 1e-2, (k, ours, theirs)


def test_convunetr_full_size_acdc_step():
    """BASELINE.json config 4: SAX 192 x 192 x 16, ViT-B encoder over 2305 tokens, ACDC decoder pyramid
    (cinema/segmentation/acdc/config.yaml:41-66), 4 classes, stochastic depth 0.1: one training step, then frozen encoder."""
    from cinema_b200.segmentation import ConvUNetR

    torch.manual_seed(0)
    model = ConvUNetR(image_size_dict={"sax": (192, 192, 16)}, in_chans_dict={"sax": 1}, out_chans=4,
                      enc_patch_size_dict={"sax": (4, 4, 1)}, enc_scale_factor_dict={"sax": (2, 2, 1)}, enc_conv_chans=[64, 128],
                      enc_conv_n_blocks=2, enc_embed_dim=768, enc_depth=12, enc_n_heads=12, dec_chans=(32, 64, 128, 256, 512),
                      dec_patch_size_dict={"sax": (2, 2, 1)}, dec_scale_factor_dict={"sax": (2, 2, 1)}, drop_path=0.1).to(DEV)
    model.train()
    x = {"sax": torch.rand(2, 1, 192, 192, 16, device=DEV)}
    y = torch.randint(0, 4, (2, 192, 192, 16), device=DEV)
    out = model(x)["sax"]
    assert out.shape == (2, 4, 192, 192, 16) and bool(torch.isfinite(out).all())
    torch.nn.functional.cross_entropy(out.float(), y).backward()
    for k, p in model.named_parameters():
        assert (p.grad is not None) == p.requires_grad, k
        if p.grad is not None:
            assert bool(torch.isfinite(p.grad).all()), k
    assert float(model.encoder.blocks[0].attn.q.weight.grad.abs().sum()) > 0
    assert float(model.enc_down_dict["sax"].conv_blocks[0].patch_embed.conv.weight.grad.abs().sum()) > 0
    model.zero_grad(set_to_none=True)
    for p in [*model.enc_down_dict.parameters(), *model.encoder.parameters()]:
        p.requires_grad = False
    torch.nn.functional.cross_entropy(model(x)["sax"].float(), y).backward()
    assert model.encoder.blocks[0].attn.q.weight.grad is None
    assert float(model.decoder_dict["sax"].blocks[0].up.weight.grad.abs().sum()) > 0


def test_trainer_prefetch_equals_direct_upload(golden_dir):
    """``prefetch`` (copy stream + staging buffer) feeds the step the same bytes as the in-line upload: two trainers on the
    same weights / RNG stream, one fed through prefetch with a DIFFERENT batch each step, produce identical losses."""
    from cinema_b200.train import MAETrainer

    g = torch.load(golden_dir / "mae_small_4view.pt")
    batches = []
    for i in range(5):
        gen = torch.Generator().manual_seed(100 + i)
        batches.append({k: torch.rand(v.shape, generator=gen).pin_memory() for k, v in g["images"].items()})
    losses = {}
    for mode in ("direct", "prefetch"):
        torch.manual_seed(0)
        model = CineMA(**g["kw"]).to(DEV)
        model.load_state_dict(g["state_dict"])
        model.train()
        tr = MAETrainer(model, lr=1e-3, use_cuda_graph=True, graph_warmup=2)
        torch.manual_seed(1)
        out = []
        if mode == "prefetch":
            tr.prefetch(batches[0])
        for i, b in enumerate(batches):
            loss = tr.step(b)
            if mode == "prefetch" and i + 1 < len(batches):
                tr.prefetch(batches[i + 1])
            out.append(float(loss))
        losses[mode] = out
    # the two runs differ only through the order of fp32 atomics (split-K / bias-gradient accumulation), which five AdamW steps
    # amplify to ~1e-5 of the loss (one run in six exceeded 1e-5); a wrong or stale batch would show up at the 1e-2 level
    assert all(abs(a - b) <= 5e-4 * abs(a) for a, b in zip(losses["direct"], losses["prefetch"])), losses
    assert len(set(losses["direct"])) == len(batches)  # the batches really differ


@pytest.mark.parametrize("name", ["cfg2_sax_only_vitb", "cfg5_vitl_4view_lax256"])
def test_other_baseline_configs_train(name):
    """BASELINE.json configs 2 (SAX 192 x 192 x 16 alone, ViT-B) and 5 (ViT-L encoder, 4 views, the reference's LAX 256 x 256):
    a few optimiser steps through the CUDA-graph trainer on a fixed batch -- finite, decreasing loss, graph replay equals
    what eager produced before capture, parameter count of the reference's size table."""
    from bench import model_kwargs
    from cinema_b200.train import MAETrainer

    if name == "cfg2_sax_only_vitb":
        kw = model_kwargs("base", (192, 192, 16), (192, 192))
        for key in ("image_size_dict", "in_chans_dict", "enc_patch_size_dict", "enc_scale_factor_dict"):
            kw[key] = {"sax": kw[key]["sax"]}
        n_params = None
    else:
        kw = model_kwargs("large", (192, 192, 16), (256, 256))
        n_params = 24 * (4 * 1024 * 1024 + 4 * 1024 + 8 * 1024 * 1024 + 5 * 1024 + 4 * 1024)  # encoder blocks of ViT-L
    torch.manual_seed(0)
    model = CineMA(**kw).to(DEV).train()
    if n_params is not None:
        assert sum(p.numel() for p in model.encoder.blocks.parameters()) == n_params
    gen = torch.Generator().manual_seed(3)
    images = {v: torch.rand(2, 1, *s, generator=gen).to(DEV) for v, s in kw["image_size_dict"].items()}
    tr = MAETrainer(model, lr=2e-4, use_cuda_graph=True, graph_warmup=2)
    losses = [float(tr.step(images)) for _ in range(6)]
    assert all(math.isfinite(x) for x in losses), losses
    assert losses[-1] < losses[0], losses
    assert tr.use_graph and tr._g_fb is not None


def test_trainer_split_graph_equals_single_graph(golden_dir, monkeypatch):
    """The step captured as a CHAIN of CUDA graphs cut at the gradient stages (how the gradient all-reduce overlaps the
    backward on several GPUs; forced here on one rank) replays to the same losses and parameters as the single graph."""
    from cinema_b200.train import MAETrainer

    g = torch.load(golden_dir / "mae_small_4view.pt")
    images = _to(g["images"], DEV)
    out = {}
    for split in (False, True):
        monkeypatch.setenv("CB_FORCE_SPLIT", "1" if split else "0")
        torch.manual_seed(0)
        model = CineMA(**g["kw"]).to(DEV)
        model.load_state_dict(g["state_dict"])
        model.train()
        tr = MAETrainer(model, lr=1e-3, use_cuda_graph=True, graph_warmup=2)
        assert tr._direct and tr.overlap == split
        torch.manual_seed(1)
        losses = [float(tr.step(images)) for _ in range(6)]
        assert len(tr._fb_segments) == (3 if split else 0)
        out[split] = (losses, tr.arena.flat32.clone())
    assert all(abs(a - b) <= 1e-4 * abs(b) for a, b in zip(out[True][0], out[False][0]))  # atomics reorder fp32 sums
    assert rel(out[True][1], out[False][1]) < 1e-3


def test_device_feeder_on_cuda_matches_host_computation(tmp_path):
    """Input pipeline (cinema_b200/data.py): raw integer batches uploaded on the copy stream and scaled on the device equal
    the host-side ScaleIntensity of the same frames, batch after batch (double-buffered slots are not overwritten early)."""
    import numpy as np

    from cinema_b200 import data as D

    rng = np.random.default_rng(0)
    subjects = [(f"s{i}", {"sax": rng.integers(0, 3000, size=(40, 48, 6, 4)).astype(np.int16),
                           "lax_4c": rng.integers(0, 256, size=(56, 60, 4)).astype(np.uint8)}) for i in range(9)]
    D.write_shards(tmp_path, subjects)
    ds = D.CineShardDataset(tmp_path)
    sizes = {"sax": (48, 48, 8), "lax_4c": (64, 64)}
    mk = lambda pin: D.FrameBatcher(ds, D.ShardSampler(len(ds), seed=1), 2, sizes, n_frames=4, seed=3, pin_memory=pin)  # noqa: E731
    host = [{v: D.scale_intensity(r.images[v], r.lo[v], r.hi[v]).clone() for v in r.images} for r in mk(False)]
    feeder = D.DeviceFeeder(mk(True), DEV)
    got = []
    for batch in feeder:
        assert all(t.is_cuda and t.dtype == torch.float32 for t in batch.values())
        got.append({v: t.clone() for v, t in batch.items()})
    torch.cuda.synchronize()
    assert len(got) == len(host) == 4 and feeder.h2d_bytes_per_batch == 2 * (48 * 48 * 8 * 2 + 64 * 64) + 4 * 2 * 4
    for a, b in zip(host, got):
        for v in a:
            torch.testing.assert_close(b[v].cpu(), a[v], rtol=1e-6, atol=1e-7)


def test_trainer_steps_without_host_sync_match_synced_steps(golden_dir):
    """``MAETrainer.step`` never synchronises, so the host runs ahead of the device; the per-step scalars (lr, Adam bias
    corrections) travel through a ring of pinned slots guarded by events, so a queued H2D copy can never read a later
    step's values.  Twelve steps (more than the ring holds) issued back to back with a changing lr and NO host sync must
    give the same losses and parameters as the same steps with a device synchronisation after each."""
    from cinema_b200.train import MAETrainer, cosine_lr

    g = torch.load(golden_dir / "mae_small_4view.pt")
    images = _to(g["images"], DEV)
    out = {}
    for sync in (True, False):
        torch.manual_seed(0)
        model = CineMA(**g["kw"]).to(DEV)
        model.load_state_dict(g["state_dict"])
        model.train()
        tr = MAETrainer(model, lr=1e-3, use_cuda_graph=True, graph_warmup=2)
        torch.manual_seed(1)
        losses = []
        for i in range(12):
            tr.set_lr(cosine_lr(i, 3, 12, 2e-3, 1e-5))
            losses.append(tr.step(images).clone())
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        out[sync] = ([float(x) for x in losses], tr.arena.flat32.clone())
    # fp32 atomics reorder sums and twelve updates at lr 2e-3 amplify that; one wrong bias correction (e.g. step 1 using
    # step 2's: update x 0.53) moves the parameters by ~1e-2 and the following losses by > 1e-2
    worst = max(abs(a - b) / abs(b) for a, b in zip(out[False][0], out[True][0]))
    assert worst <= 2e-3, (worst, out[False][0], out[True][0])
    assert rel(out[False][1], out[True][1]) < 5e-3


def test_user_masks_with_ragged_counts_are_rejected(golden_dir):
    g = torch.load(golden_dir / "mae_small_4view.pt")
    model = CineMA(**g["kw"]).to(DEV).train()
    masks = {k: v.clone() for k, v in g["masks"].items()}
    v0 = next(iter(masks))
    row = masks[v0][1]
    row[int(torch.nonzero(row)[0])] = False  # sample 1 keeps one more patch than sample 0
    with pytest.raises(ValueError, match="different number of patches"):
        model(_to(g["images"], DEV), g["ratio"], enc_mask_dict=_to(masks, DEV))
