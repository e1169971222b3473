"""Informational timing of the fine-tuning models on one B200 (not part of bench.py's contract line):

    python tools/bench_finetune.py [--batch 4] [--steps 10]

BASELINE.json config 4 (ConvUNetR, SAX 192 x 192 x 16, ViT-B encoder over 2305 tokens, ACDC decoder pyramid, 4 classes,
cross-entropy) and the ConvViT classifier on the same input, forward + backward per step, against the STOCK torch path
on the same GPU: the oracle's functional restatement of the reference modules (F.linear / F.layer_norm / F.conv3d /
SDPA ...) under ``torch.autocast(bf16)`` with autograd -- the same yardstick as tests/perf_stock_torch_gpu.py.  Both arms
see the same weights and inputs; the reference's grad-ckpt default (one more forward of recompute) is OFF for the stock
arm, which favours it.  Writes gpurun_out/finetune_bench.json.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from cinema_b200 import ConvViT  # noqa: E402
from cinema_b200.segmentation import ConvUNetR  # noqa: E402
from oracle import cinema_oracle as O  # noqa: E402

DEV = "cuda"


def timed(fn, steps: int, warmup: int) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--channels-last", type=int, default=1)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    b = a.batch
    stem = dict(image_size_dict={"sax": (192, 192, 16)}, in_chans_dict={"sax": 1}, enc_patch_size_dict={"sax": (4, 4, 1)},
                enc_scale_factor_dict={"sax": (2, 2, 1)}, enc_conv_chans=[64, 128], enc_conv_n_blocks=2, enc_embed_dim=768,
                enc_depth=12, enc_n_heads=12)
    x = {"sax": torch.rand(b, 1, 192, 192, 16, device=DEV)}
    out = {"batch": b, "steps": a.steps, "channels_last": bool(a.channels_last), "torch": torch.__version__, "input": "SAX 192x192x16, ViT-B, 2305 encoder tokens"}

    # ---- segmentation (config 4)
    seg_kw = dict(**stem, out_chans=4, dec_chans=(32, 64, 128, 256, 512), dec_patch_size_dict={"sax": (2, 2, 1)},
                  dec_scale_factor_dict={"sax": (2, 2, 1)})
    seg = ConvUNetR(**seg_kw, channels_last=bool(a.channels_last)).to(DEV).train()
    y = torch.randint(0, 4, (b, 192, 192, 16), device=DEV)

    from cinema_b200.segmentation.loss import segmentation_loss, segmentation_loss_restated

    y1 = y.unsqueeze(1)

    def ours_seg():  # the reference's loss (cross-entropy + foreground Dice, cinema/segmentation/train.py:77-103), fused kernels
        seg.zero_grad(set_to_none=True)
        segmentation_loss(seg(x)["sax"], y1)[0].backward()

    cfg = O.convunetr_config({k: v for k, v in seg_kw.items()})
    params = {k: v.detach().clone().requires_grad_(not k.endswith("pos_embed")) for k, v in seg.state_dict().items()}

    def stock_seg():
        for p in params.values():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = O.convunetr_forward(params, cfg, x, seg.n_layers_wo_skip)["sax"]
        segmentation_loss_restated(logits, y1)[0].backward()  # the same loss as torch ops (what the stock script runs)

    t_ours, t_stock = timed(ours_seg, a.steps, a.warmup), timed(stock_seg, a.steps, a.warmup)
    # opt-in: the ConvResBlock convolutions with >= 32 input channels as tcgen05 GEMMs (ConvUNetR.set_native_convs)
    seg.set_native_convs(True)
    try:
        ref_logits = None
        seg.set_native_convs(False)
        with torch.no_grad():
            ref_logits = seg(x)["sax"].float()
        seg.set_native_convs(True)
        with torch.no_grad():
            nat_logits = seg(x)["sax"].float()
        rel = float((nat_logits - ref_logits).norm() / ref_logits.norm())
        t_nat = timed(ours_seg, a.steps, a.warmup)
        out["convunetr_acdc_native_convs"] = {"ours_ms": round(t_nat, 2), "speedup_vs_stock": round(t_stock / t_nat, 2),
                                              "logits_rel_l2_vs_cudnn_path": rel}
    except Exception as exc:  # noqa: BLE001 -- the opt-in path must not take the benchmark down
        out["convunetr_acdc_native_convs"] = {"error": repr(exc)[:300]}
    seg.set_native_convs(False)
    out["convunetr_acdc"] = {"ours_ms": round(t_ours, 2), "stock_torch_ms": round(t_stock, 2),
                             "ours_volumes_per_s": round(b / t_ours * 1e3, 1), "stock_volumes_per_s": round(b / t_stock * 1e3, 1),
                             "speedup": round(t_stock / t_ours, 2)}
    del seg, params
    torch.cuda.empty_cache()

    # ---- classification
    cls_kw = dict(**stem, n_frames=1, out_chans=4)
    clf = ConvViT(**cls_kw).to(DEV).train()
    lab = torch.randint(0, 4, (b,), device=DEV)

    def ours_cls():
        clf.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(clf(x), lab).backward()

    ccfg = O.convvit_config(cls_kw)
    cparams = {k: v.detach().clone().requires_grad_(not k.endswith("pos_embed")) for k, v in clf.state_dict().items()}

    def stock_cls():
        for p in cparams.values():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = O.convvit_forward(cparams, ccfg, x, None, "all")
        torch.nn.functional.cross_entropy(logits.float(), lab).backward()

    t_ours, t_stock = timed(ours_cls, a.steps, a.warmup), timed(stock_cls, a.steps, a.warmup)
    out["convvit_classifier"] = {"ours_ms": round(t_ours, 2), "stock_torch_ms": round(t_stock, 2),
                                 "ours_volumes_per_s": round(b / t_ours * 1e3, 1),
                                 "stock_volumes_per_s": round(b / t_stock * 1e3, 1), "speedup": round(t_stock / t_ours, 2)}
    print(json.dumps(out))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "finetune_bench.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
