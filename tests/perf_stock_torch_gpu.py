"""Informational: the STOCK torch path on the same B200 (not collected by pytest, not part of bench.py).

    python tests/perf_stock_torch_gpu.py [--batch 16] [--steps 10] [--lax 192]

Runs the oracle's functional restatement of the reference step -- the same ATen / cuBLASLt / cuDNN / SDPA calls
the reference modules make (F.linear, F.layer_norm, F.gelu, F.conv{2,3}d, F.scaled_dot_product_attention,
boolean-mask gathers) -- on cuda:0 under ``torch.autocast(bf16)`` with autograd backward and a fused torch AdamW
(cinema/mae/pretrain.py:251-261,365-367).  The reference's ``grad_ckpt: True`` default (cinema/mae/config.yaml:3)
would add one forward of recompute; it is left OFF here, which favours the stock path.  north_star asks to beat
"the reference's stock PyTorch/SDPA path on the same box": this is that number, written to
gpurun_out/stock_torch.json and quoted in BASELINE.md / profiles/.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from bench import model_kwargs, synthetic_batch, train_gflop_per_sample  # noqa: E402
from oracle import cinema_oracle as oracle  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--lax", type=int, default=192)
    ap.add_argument("--size", default="base")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True  # cinema/device.py:63
    kw = model_kwargs(a.size, (192, 192, 16), (a.lax, a.lax))
    cfg = oracle.MAEConfig(**kw)
    sd = oracle.init_state_dict(cfg, seed=0)
    params = {k: v.to(dev).requires_grad_(not k.endswith("pos_embed")) for k, v in sd.items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05,
                            fused=True)
    batch = {k: v.to(dev) for k, v in synthetic_batch(kw, a.batch, seed=0, pin=False).items()}

    def step() -> torch.Tensor:
        masks = {v: oracle.random_patch_mask(a.batch, cfg.n_patches(v), 0.75, device=dev) for v in kw["image_size_dict"]}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _, _ = oracle.mae_forward(params, cfg, batch, masks)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.requires_grad], 5.0)
        opt.step()
        return loss.detach()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.steps
    work = train_gflop_per_sample(kw)
    out = {"impl": "stock torch (oracle port on cuda, bf16 autocast, SDPA, fused AdamW, no grad-ckpt)", "batch": a.batch,
           "ms_per_step": round(ms, 3), "volumes_per_s": round(a.batch / ms * 1e3, 2),
           "step_tflops": round(a.batch / ms * work["train_gflop"], 1), "final_loss": float(loss),
           "torch": torch.__version__, "lax": a.lax, "size": a.size}
    print(json.dumps(out))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "stock_torch.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
