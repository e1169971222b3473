"""Pins oracle/cinema_oracle.py against vectors produced by the real reference
(tests/golden/make_golden.py) and against the reference's own known-answer tests."""

import math

import pytest
import torch

from oracle import cinema_oracle as O

TOL = dict(rtol=1e-5, atol=1e-6)  # fp32 CPU vs fp32 CPU, same torch kernels


def _load(golden_dir, name):
    return torch.load(golden_dir / name, weights_only=False)


def _cfg(kw):
    keys = {f for f in O.MAEConfig.__dataclass_fields__}
    return O.MAEConfig(**{k: v for k, v in kw.items() if k in keys})


@pytest.mark.parametrize("case", ["mae_tiny_sax", "mae_small_4view", "mae_tiny_selfattn_normtarget"])
def test_mae_forward_backward_matches_reference(golden_dir, case):
    g = _load(golden_dir, f"{case}.pt")
    cfg = _cfg(g["kw"])
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith("pos_embed"))
          for k, v in g["state_dict"].items()}
    loss, preds, metrics = O.mae_forward(sd, cfg, g["images"], g["masks"])
    torch.testing.assert_close(loss.detach(), g["loss"], **TOL)
    for v, p in g["preds"].items():
        torch.testing.assert_close(preds[v].detach(), p, **TOL)
    assert set(metrics) == set(g["metrics"])
    for k, m in g["metrics"].items():
        torch.testing.assert_close(metrics[k].detach(), m, **TOL)
        assert metrics[k].ndim == 0  # cinema/mae/mae_test.py:129-131
    loss.backward()
    assert g["grads"], "fixture carries no gradients"
    for k, gr in g["grads"].items():
        torch.testing.assert_close(sd[k].grad, gr, rtol=1e-4, atol=1e-6)
    # Sigma masked tokens (cinema/mae/mae_test.py:123)
    for v, m in g["masks"].items():
        assert preds[v].shape[1] == int(m.sum()) // m.shape[0]


@pytest.mark.parametrize("case", ["mae_tiny_sax", "mae_small_4view"])
def test_feature_forward_matches_reference(golden_dir, case):
    g = _load(golden_dir, f"{case}.pt")
    cfg = _cfg(g["kw"])
    with torch.no_grad():
        feats = O.mae_feature_forward(g["state_dict"], cfg, g["images"])
    assert list(feats) == list(g["feats"])
    for k, f in g["feats"].items():
        torch.testing.assert_close(feats[k], f, **TOL)


def test_state_dict_schema_matches_reference(golden_dir):
    g = _load(golden_dir, "mae_small_4view.pt")
    cfg = _cfg(g["kw"])
    mine = O.init_state_dict(cfg)
    assert set(mine) == set(g["state_dict"])
    for k, v in g["state_dict"].items():
        assert mine[k].shape == v.shape, k
        assert mine[k].dtype == v.dtype, k
    for k in mine:
        if k.endswith("pos_embed"):
            torch.testing.assert_close(mine[k], g["state_dict"][k], rtol=0, atol=0)


def test_patchify_unpatchify(golden_dir):
    ops = _load(golden_dir, "ops.pt")
    for nm in ("p2", "p3", "p4"):
        c = ops[nm]
        tok = O.patchify(c["image"], c["patch_size"])
        assert torch.equal(tok, c["tokens"])
        assert torch.equal(O.unpatchify(tok, c["patch_size"], c["grid"]), c["back"])
        assert torch.equal(c["back"], c["image"])  # round trip is exact
    with pytest.raises(ValueError):
        O.patchify(torch.zeros(1, 1, 5, 4), (2, 2))
    with pytest.raises(ValueError):
        O.unpatchify(torch.zeros(1, 4, 8), (2, 2), (2, 3))


def test_sincos_pos_embed(golden_dir):
    ops = _load(golden_dir, "ops.pt")
    for key, ref in ops["pos"].items():
        d, gs = key.split("_", 1)
        mine = O.sincos_pos_embed(int(d), eval(gs))
        assert torch.equal(mine, ref), key
    big = O.sincos_pos_embed(768, (12, 12, 16))
    torch.testing.assert_close(big.double().sum(dim=1), ops["pos_sum_768_12_12_16"], rtol=0, atol=0)


def test_rotate_half_known_answer():
    # cinema/rotary_test.py:9-13
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    assert torch.equal(O.rotate_half(x), torch.tensor([[-2.0, 1.0], [-4.0, 3.0]]))


def test_rotary_contract(golden_dir):
    r = _load(golden_dir, "ops.pt")["rotary"]
    rq, rk = O.rotary_qk(r["q"], r["k"], r["dim"])
    torch.testing.assert_close(rq, r["rq"], **TOL)
    torch.testing.assert_close(rk, r["rk"], **TOL)
    assert torch.equal(rq[..., r["dim"]:], r["q"][..., r["dim"]:])  # partial rotary dim untouched
    with pytest.raises(ValueError):
        O.rotary_qk(r["q"], r["k"][:, :3], r["dim"])


def test_attention_rotary_quirk_and_cross(golden_dir):
    ops = _load(golden_dir, "ops.pt")
    a = ops["attn_rotary"]
    sd = {f"attn.{k}": v for k, v in a["sd"].items()}
    y_rot = O.attention(sd, "attn", a["x"], None, 4, rotary=True)
    y_plain = O.attention(sd, "attn", a["x"], None, 4, rotary=False)
    torch.testing.assert_close(y_rot, a["y"], **TOL)
    # the reference's rotary call is a no-op on the attention output
    torch.testing.assert_close(y_rot, y_plain, rtol=1e-4, atol=1e-6)
    c = ops["attn_cross"]
    sd = {f"attn.{k}": v for k, v in c["sd"].items()}
    torch.testing.assert_close(O.attention(sd, "attn", c["q"], c["k"], 2), c["y"], **TOL)
    with pytest.raises(ValueError):
        O.attention(sd, "attn", c["q"], c["k"], 2, rotary=True)


def test_upsample_mask_known_answers(golden_dir):
    # cinema/convvit_test.py:21-50 (restated)
    m = torch.tensor([[True, False]])
    assert torch.equal(O.upsample_mask(m, (2,)), torch.tensor([[True, True, False, False]]))
    m = torch.tensor([[[True, False], [False, True]]])
    exp = torch.tensor([[[True, True, False, False]] * 2 + [[False, False, True, True]] * 2])
    assert torch.equal(O.upsample_mask(m, (2, 2)), exp)
    u = _load(golden_dir, "ops.pt")["upsample"]
    assert torch.equal(O.upsample_mask(u["mask"], u["sf"]), u["out"])
    with pytest.raises(ValueError):
        O.upsample_mask(torch.zeros(1, 2, 2, dtype=torch.bool), (2,))


def test_random_mask_counts(golden_dir):
    # cinema/mae/mae_test.py:30-32: exactly int(n (1-r)) kept per row; same RNG stream as the reference
    ops = _load(golden_dir, "ops.pt")
    torch.manual_seed(6)
    for key, ref in ops["mask_counts"].items():
        n, r = key.split("_")
        n, r = int(n), float(r)
        mine = O.random_patch_mask(4, n, r)
        assert mine.shape == (4, n) and mine.dtype == torch.bool
        assert torch.equal((~mine).sum(dim=1), torch.full((4,), int(n * (1 - r)) if r > 0 else n))
        assert torch.equal(mine, ref)
    with pytest.raises(ValueError):
        O.random_patch_mask(1, 4, -0.1)


@pytest.mark.parametrize("norm", [0, 1])
def test_masked_mse(golden_dir, norm):
    c = _load(golden_dir, "ops.pt")[f"loss_norm{norm}"]
    loss, met = O.masked_mse(c["target"], c["pred"], c["mask"], bool(norm))
    torch.testing.assert_close(loss, c["loss"], **TOL)
    assert set(met) == set(c["metrics"])
    for k in met:
        torch.testing.assert_close(met[k], c["metrics"][k], **TOL)


def test_config_shapes():
    cfg = O.make_config("base")
    assert cfg.grid_size("sax") == (12, 12, 16) and cfg.n_patches("sax") == 2304
    assert cfg.grid_size("lax_2c") == (16, 16) and cfg.dec_patch_size("sax") == (16, 16, 1)
    assert math.prod(cfg.dec_patch_size("lax_4c")) == 256


@pytest.mark.parametrize("case", ["convvit_2view", "convvit_frames_resized"])
def test_convvit_oracle_matches_reference(golden_dir, case):
    """ConvViT restatement (logits for every reduce mode, stem mask, resized input with interpolated positional table,
    gradients) against the real reference's outputs."""
    g = _load(golden_dir, f"{case}.pt")
    cfg = O.convvit_config(g["kw"])
    sd = {k: v.clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    with torch.no_grad():
        feats = O.convvit_feature_forward(sd, cfg, g["images"], None)
        for k, f in g["feats"].items():
            torch.testing.assert_close(feats[k], f, **TOL)
        for reduce, ref in g["logits"].items():
            torch.testing.assert_close(O.convvit_forward(sd, cfg, g["images"], None, reduce), ref, **TOL)
        torch.testing.assert_close(O.convvit_forward(sd, cfg, g["images"], g["mask_dict"], "all"),
                                   g["logits_masked"]["all"], **TOL)
    for masks, want in ((None, g["grads"]), (g["mask_dict"], g["grads_masked"])):
        for p in sd.values():
            p.grad = None
        (O.convvit_forward(sd, cfg, g["images"], masks, "all") * g["w"]).sum().backward()
        for k, gr in want.items():
            torch.testing.assert_close(sd[k].grad, gr, rtol=1e-4, atol=1e-6)


def test_convvit_oracle_stochastic_depth_matches_reference(golden_dir):
    """Training-mode DropPath: replaying the per-sample factors the reference drew reproduces its logits and gradients."""
    g = _load(golden_dir, "convvit_droppath.pt")
    cfg = O.convvit_config(g["kw"])
    sd = {k: v.clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    out = O.convvit_forward(sd, cfg, g["images"], None, "all", drop_scales=g["drop_scales"])
    torch.testing.assert_close(out.detach(), g["logits"]["all"], **TOL)
    (out * g["w"]).sum().backward()
    for k, gr in g["grads"].items():
        torch.testing.assert_close(sd[k].grad, gr, rtol=1e-4, atol=1e-6)
    with torch.no_grad():  # eval mode: identity
        feats = O.convvit_feature_forward(sd, cfg, g["images"], None)
    for k, f in g["feats"].items():
        torch.testing.assert_close(feats[k], f, **TOL)


def test_convunetr_oracle_matches_reference(golden_dir):
    """Segmentation model restatement (stem -> encoder -> UNETR decoder) against the real reference: logits per view and
    gradients of a weighted logit sum."""
    g = _load(golden_dir, "convunetr_2view.pt")
    cfg = O.convunetr_config(g["kw"])
    sd = {k: v.clone().requires_grad_(not k.endswith("pos_embed")) for k, v in g["state_dict"].items()}
    preds = O.convunetr_forward(sd, cfg, g["images"], g["n_layers_wo_skip"])
    for v, ref in g["preds"].items():
        torch.testing.assert_close(preds[v].detach(), ref, rtol=1e-4, atol=1e-5)
    sum((preds[v] * g["w"][v]).sum() for v in preds).backward()
    for k, gr in g["grads"].items():
        torch.testing.assert_close(sd[k].grad, gr, rtol=2e-4, atol=1e-5)
    with torch.no_grad():
        out = O.convunetr_forward(sd, cfg, {"sax": g["images"]["sax"]}, g["n_layers_wo_skip"])
    torch.testing.assert_close(out["sax"], g["sax_only"], rtol=1e-4, atol=1e-5)
