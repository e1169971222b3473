"""Stock arms: the UNMODIFIED reference modules (``baseline/_ref``, see ``install_reference.py``) driven the way
``cinema/mae/pretrain.py:242-272`` drives them.

* ``stock_gpu_step_time``: ``CineMA.forward`` under ``torch.autocast("cuda", bf16)`` (SDPA, cuBLASLt, cuDNN),
  ``cinema.optim.GradScaler`` (scale -> backward -> unscale -> ``clip_grad_norm_(5.0)`` -> step -> update),
  ``torch.optim.AdamW`` over timm-style weight-decay groups, ``optimizer.zero_grad()``, per-step
  ``metrics[...].item()`` and ``torch.cuda.synchronize()`` -- the reference loop, nothing of this repo on the path.
  This is the number BASELINE.json's north_star asks to beat ("the reference's stock PyTorch/SDPA path on the same box").
* ``cpu_step_time``: the same modules on the host cores in fp32 (the reference disables autocast without CUDA,
  ``cinema/mae/pretrain.py:251``); forward + backward + clip + AdamW.

timm / omegaconf are not installable offline: the stub modules below restate the handful of names the reference
imports (timm 1.0.15 ``Mlp`` = fc2(drop(norm(drop(act(fc1 x))))), ``DropPath``, ``LayerScale``,
``param_groups_weight_decay``), as ``tests/golden/make_golden.py`` does for the golden vectors.
"""

from __future__ import annotations

import os
import sys
import time
import types
from pathlib import Path

import torch
from torch import nn

REF_ROOT = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_ROOT / "cinema" / "mae" / "mae.py").exists()


def _install_shims() -> None:
    if "cinema.mae.mae" in sys.modules:
        return
    # the installed package's __init__ imports monai (absent): register the package without executing it
    for name, rel in (("cinema", "cinema"), ("cinema.mae", "cinema/mae")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [str(REF_ROOT / rel)]
        sys.modules[name] = pkg

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                     bias=True, drop=0.0, use_conv=False):
            super().__init__()
            out_features, hidden_features = out_features or in_features, hidden_features or in_features
            bias, drop = to_2tuple(bias), to_2tuple(drop)
            self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
            self.act = act_layer()
            self.drop1 = nn.Dropout(drop[0])
            self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
            self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])
            self.drop2 = nn.Dropout(drop[1])

        def forward(self, x):
            return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            t = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                t.div_(keep)
            return x * t

    class LayerScale(nn.Module):
        def __init__(self, dim, init_values=1e-5, inplace=False):
            super().__init__()
            self.gamma = nn.Parameter(init_values * torch.ones(dim))

        def forward(self, x):
            return x * self.gamma

    class SwiGLU(nn.Module):
        pass

    timm, layers = types.ModuleType("timm"), types.ModuleType("timm.layers")
    models, vt = types.ModuleType("timm.models"), types.ModuleType("timm.models.vision_transformer")
    layers.Mlp, layers.SwiGLU, layers.DropPath, layers.to_2tuple = Mlp, SwiGLU, DropPath, to_2tuple
    layers.use_fused_attn = lambda: True  # timm default unless TIMM_FUSED_ATTN=0 -> F.scaled_dot_product_attention
    vt.LayerScale = LayerScale
    timm.layers, timm.models, models.vision_transformer = layers, models, vt
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = type("DictConfig", (dict,), {})
    oc.OmegaConf = type("OmegaConf", (), {})
    for name, mod in (("timm", timm), ("timm.layers", layers), ("timm.models", models),
                      ("timm.models.vision_transformer", vt), ("omegaconf", oc)):
        sys.modules.setdefault(name, mod)


def param_groups_weight_decay(model: nn.Module, weight_decay: float) -> list[dict]:
    """timm.optim.param_groups_weight_decay (1.0.15): no decay for 1-D parameters and biases (cinema/mae/pretrain.py:365)."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if p.ndim <= 1 or name.endswith(".bias") else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def build_reference_model(kw: dict, grad_ckpt: bool, seed: int = 0) -> nn.Module:
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_reference.py` where /root/reference exists")
    _install_shims()
    from cinema.mae.mae import CineMA  # type: ignore  # the reference class, unmodified

    torch.manual_seed(seed)
    model = CineMA(**kw)
    model.set_grad_ckpt(grad_ckpt)
    return model


def _make_step(model: nn.Module, device: torch.device, lr: float, betas, weight_decay: float, clip_grad: float, ratio: float):
    _install_shims()
    from cinema.optim import GradScaler  # type: ignore  # reference class (cinema/optim.py:173-226)

    optimizer = torch.optim.AdamW(param_groups_weight_decay(model, weight_decay), lr=lr, betas=tuple(betas))
    scaler = GradScaler()
    on_cuda = device.type == "cuda"

    def step(batch: dict) -> float:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=on_cuda):
            loss, _, _, metrics = model({k: v.to(device) for k, v in batch.items()}, ratio)
        metrics = {k: v.item() for k, v in metrics.items()}  # cinema/mae/pretrain.py:253
        if torch.isnan(loss).any():
            return float("nan")
        scaler(loss=loss, optimizer=optimizer, clip_grad=clip_grad, parameters=model.parameters(), update_grad=True)
        optimizer.zero_grad()
        if on_cuda:
            torch.cuda.synchronize()
        return metrics["loss"] if "loss" in metrics else float(loss)

    return step


def stock_gpu_step_time(kw: dict, batch_host: dict, device: torch.device, steps: int, warmup: int, grad_ckpt: bool,
                        lr: float = 1e-3, betas=(0.9, 0.95), weight_decay: float = 0.05, clip_grad: float = 5.0,
                        ratio: float = 0.75, ddp: bool = False, local_rank: int = 0) -> dict:
    """ms per optimisation step of the reference loop on ``device`` (CUDA events around ``steps`` steps, pinned host batch
    uploaded every step like the reference's ``v.to(device)``)."""
    torch.backends.cudnn.benchmark = True  # cinema/device.py:63
    model = build_reference_model(kw, grad_ckpt).to(device).train()
    if ddp:  # cinema/device.py:86-104 (setup_ddp_model)
        from torch.nn.parallel import DistributedDataParallel

        model = DistributedDataParallel(model, device_ids=[local_rank], find_unused_parameters=False)
    torch.cuda.reset_peak_memory_stats(device)
    step = _make_step(model, device, lr, betas, weight_decay, clip_grad, ratio)
    for _ in range(warmup):
        step(batch_host)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    last = 0.0
    for _ in range(steps):
        last = step(batch_host)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    b = next(iter(batch_host.values())).shape[0]
    peak_gb = torch.cuda.max_memory_allocated(device) / 2 ** 30
    del model, step
    torch.cuda.empty_cache()
    return {"value": b / (ms / 1e3), "unit": "frame-set volumes/s", "ms_per_step": ms, "grad_ckpt": grad_ckpt, "kind": "_ref",
            "batch_per_gpu": b, "steps": steps, "warmup": warmup, "final_loss": last, "peak_mem_gib": round(peak_gb, 2),
            "how": "unmodified reference CineMA (baseline/_ref) under torch.autocast(bf16) + SDPA, cinema.optim.GradScaler "
                   "(clip 5.0) + torch.optim.AdamW, the loop of cinema/mae/pretrain.py:242-272, same B200, same process"}


def cpu_step_time(kw: dict, batch_host: dict, steps: int, warmup: int, grad_ckpt: bool = False) -> dict:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_reference_model(kw, grad_ckpt).train()
    step = _make_step(model, torch.device("cpu"), 1e-3, (0.9, 0.95), 0.05, 5.0, 0.75)
    for _ in range(warmup):
        step(batch_host)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(batch_host)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    b = next(iter(batch_host.values())).shape[0]
    return {"value": b / dt, "unit": "frame-set volumes/s", "cores": cores, "kind": "reference", "ms_per_step": dt * 1e3,
            "sample": f"{steps} optimisation steps (forward + backward + clip + AdamW) of {b} frame-sets through the unmodified "
                      f"reference modules (baseline/_ref; fp32: the reference disables autocast on CPU; torch {torch.__version__}, "
                      f"{cores} threads), {dt:.2f} s per step"}
