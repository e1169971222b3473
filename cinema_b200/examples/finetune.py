"""Fine-tuning steps for the classification / regression (``ConvViT``) and segmentation (``ConvUNetR``) models, the
counterparts of the loops in ``cinema/examples/train/{classification,regression,segmentation}.py``.

The models run their encoder on the B200 kernels inside autograd, so the reference's recipe carries over unchanged:
``param_groups_lr_decay`` groups -> ``torch.optim.AdamW`` -> cosine schedule with per-group ``lr_scale`` -> loss -> backward ->
``clip_grad_norm_`` -> step.  No GradScaler: the kernels accumulate in fp32 and bf16 has fp32's exponent range
(cinema/optim.py:173-218 only scales for fp16)."""

from __future__ import annotations

from typing import Callable, Iterable

import torch
import torch.nn.functional as F  # noqa: N812
from torch import nn

from cinema_b200.convvit import param_groups_lr_decay
from cinema_b200.train import adjust_learning_rate


def build_optimizer(model: nn.Module, lr: float, weight_decay: float, layer_decay: float, betas=(0.9, 0.999)) -> torch.optim.Optimizer:
    """AdamW over layer-wise lr-decay groups (cinema/examples/train/classification.py:199-214)."""
    groups = param_groups_lr_decay(model, no_weight_decay_list=[], weight_decay=weight_decay, layer_decay=layer_decay)
    return torch.optim.AdamW(groups, lr=lr, betas=betas)


def classification_loss(logits: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    return F.cross_entropy(logits, label.long(), label_smoothing=0.1)  # cinema/examples/train/classification.py:239-243


def regression_loss(preds: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    return F.mse_loss(preds.squeeze(-1), label.to(preds.dtype))  # cinema/examples/train/regression.py


def segmentation_loss(logits: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """Cross-entropy (``ignore_index=-1``) + MONAI ``DiceLoss(include_background=False, softmax=True)``, the loss of
    cinema/segmentation/train.py:77-103 and cinema/examples/train/segmentation.py:238-242, through the fused forward / backward
    kernels (``cinema_b200.segmentation.loss``).  ``label``: (B, *spatial) or (B, 1, *spatial) integers, -1 = unlabelled."""
    from cinema_b200.segmentation.loss import segmentation_loss as fused

    if label.dim() == logits.dim() - 1:
        label = label.unsqueeze(1)
    return fused(logits, label)[0]


def finetune_one_epoch(model: nn.Module, batches: Iterable[tuple[dict[str, torch.Tensor], torch.Tensor]],
                       optimizer: torch.optim.Optimizer, loss_fn: Callable, *, epoch: int, n_batches: int, n_epochs: int,
                       n_warmup_epochs: float, lr: float, min_lr: float, clip_grad: float | None = None,
                       view: str | None = None) -> list[float]:
    """One epoch of (image_dict, label) batches.  ``view``: for models that return a dict of per-view outputs
    (ConvUNetR) the view whose logits the loss is taken on."""
    model.train()
    losses = []
    for i, (image_dict, label) in enumerate(batches):
        adjust_learning_rate(optimizer, step=i / max(n_batches, 1) + epoch, warmup_steps=n_warmup_epochs, max_n_steps=n_epochs,
                             lr=lr, min_lr=min_lr)
        out = model(image_dict)
        if isinstance(out, dict):
            out = out[view if view is not None else next(iter(out))]
        loss = loss_fn(out, label)
        optimizer.zero_grad(set_to_none=False)  # gradients live in the flat arena: clear in place
        loss.backward()
        if clip_grad is not None and clip_grad > 0:
            torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], clip_grad)
        optimizer.step()
        losses.append(float(loss.detach()))
    return losses
