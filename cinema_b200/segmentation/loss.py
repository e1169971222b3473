"""Segmentation loss of the reference's fine-tuning scripts on the B200 path (``cinema/segmentation/train.py:77-103``):

    ce   = F.cross_entropy(logits, labels.squeeze(1), ignore_index=-1)
    dice = monai.losses.DiceLoss(include_background=False, to_onehot_y=False, softmax=True)(logits, one_hot(labels.clamp(min=0)))
    loss = dice + ce

as ONE forward pass + ONE backward pass over the logits (``csrc/segloss.cu``) instead of the ~12 elementwise / reduction
kernels of the stock path.  ``segmentation_loss`` keeps the reference's signature and metric names.  MONAI is not installed
here: the Dice term follows MONAI's published definition (softmax over the channel axis, per sample and foreground class
``1 - (2 sum(p y) + 1e-5) / (sum(p) + sum(y) + 1e-5)``, mean over batch and classes), restated in
``segmentation_loss_restated`` below, which the kernels are tested against -- against a MONAI run this is parity unpinned.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from cinema_b200 import _C


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        logits = logits.contiguous()
        labels = labels.contiguous()
        out, coef = _C.seg_loss_fwd(logits, labels)
        ctx.save_for_backward(logits, labels, coef)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):  # type: ignore[override]
        logits, labels, coef = ctx.saved_tensors
        # out = [loss, ce, dice]: only out[0] (= ce + dice, what the training loop minimises) is differentiable here --
        # ``segmentation_loss`` hands the other two out detached, so their incoming gradient is zero by construction (it is
        # not inspected: that would be a device-to-host sync in every backward)
        return _C.seg_loss_bwd(logits, labels, coef, grad_out.to(torch.float32)[:1].contiguous()), None


def segmentation_loss(logits: torch.Tensor, labels: torch.Tensor) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
    """``_segmentation_loss`` of cinema/segmentation/train.py:77-103.

    Args:
        logits: (batch, n_classes, ...) fp32 or bf16, 2 <= n_classes <= 8.
        labels: (batch, 1, ...) integer labels (int64 / int32 / int16 / uint8); -1 marks unlabelled voxels.

    Returns:
        loss: 0-d tensor (differentiable w.r.t. ``logits``); metrics: ``cross_entropy``, ``mean_dice_loss``, ``loss`` (detached).
    """
    if labels.dim() != logits.dim() or labels.shape[1] != 1 or labels.shape[0] != logits.shape[0] or labels.shape[2:] != logits.shape[2:]:
        raise ValueError(f"labels must be (batch, 1, *spatial) matching the logits {tuple(logits.shape)}, got {tuple(labels.shape)}")
    if labels.dtype not in _C._LABEL_DT:
        labels = labels.long()  # the reference's labels.long()
    if logits.dtype not in (torch.float32, torch.bfloat16):
        logits = logits.float()
    out = _SegLossFn.apply(logits, labels)
    det = out.detach()
    return out[0], {"cross_entropy": det[1], "mean_dice_loss": det[2], "loss": det[0]}


def segmentation_loss_restated(logits: torch.Tensor, labels: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Plain-torch fp32 restatement of the same loss (test reference; MONAI ``DiceLoss`` spelled out): -> (loss, ce, dice)."""
    lg = logits.float()
    lab = labels.long()
    n_cls = lg.shape[1]
    ce = F.cross_entropy(lg, lab.squeeze(1), ignore_index=-1)
    onehot = F.one_hot(lab.clamp(min=0).squeeze(1), n_cls).movedim(-1, 1).to(lg.dtype)
    prob = lg.softmax(dim=1)
    dims = tuple(range(2, lg.dim()))
    inter = (prob[:, 1:] * onehot[:, 1:]).sum(dims)
    denom = prob[:, 1:].sum(dims) + onehot[:, 1:].sum(dims)
    dice = (1.0 - (2.0 * inter + 1e-5) / (denom + 1e-5)).mean()
    return dice + ce, ce, dice
