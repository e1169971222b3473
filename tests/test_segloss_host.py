"""Segmentation loss (cinema_b200/segmentation/loss.py; reference cinema/segmentation/train.py:77-103), host logic on the CPU
through the emulated kernels: signature / metrics of the reference, the ignore index, label dtypes, shape checks, and the
restatement itself against its definition (F.cross_entropy + MONAI's DiceLoss formula spelled out by hand)."""

import pytest
import torch
import torch.nn.functional as F

from cinema_b200.segmentation.loss import segmentation_loss, segmentation_loss_restated


def _case(shape=(2, 4, 12, 10, 3), seed=0, ignore=True):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(shape, generator=g) * 2.0
    labels = torch.randint(0, shape[1], (shape[0], 1, *shape[2:]), generator=g)
    if ignore:
        labels[torch.rand(labels.shape, generator=g) < 0.2] = -1
    return logits, labels


def test_restatement_is_ce_plus_foreground_dice_by_hand():
    logits, labels = _case()
    loss, ce, dice = segmentation_loss_restated(logits, labels)
    assert torch.allclose(ce, F.cross_entropy(logits, labels.squeeze(1), ignore_index=-1))
    p = logits.softmax(1)
    y = labels.clamp(min=0).squeeze(1)
    acc = 0.0
    for b in range(logits.shape[0]):
        for c in range(1, logits.shape[1]):  # include_background=False
            t = (y[b] == c).float()
            acc += 1.0 - (2.0 * (p[b, c] * t).sum() + 1e-5) / (p[b, c].sum() + t.sum() + 1e-5)
    assert torch.allclose(dice, acc / (logits.shape[0] * (logits.shape[1] - 1)), atol=1e-6)
    assert torch.allclose(loss, ce + dice)


@pytest.mark.parametrize("label_dtype", [torch.int64, torch.int32, torch.int16, torch.int8])
def test_segmentation_loss_forward_backward_and_metrics(emulated_kernels, label_dtype):
    logits, labels = _case(shape=(2, 4, 16, 12), seed=3)
    logits.requires_grad_(True)
    loss, metrics = segmentation_loss(logits, labels.to(label_dtype))  # int8 is converted with labels.long(), as the reference does
    ref = logits.detach().clone().requires_grad_(True)
    ref_loss, ref_ce, ref_dice = segmentation_loss_restated(ref, labels)
    assert set(metrics) == {"cross_entropy", "mean_dice_loss", "loss"}  # cinema/segmentation/train.py:102
    assert torch.allclose(loss, ref_loss) and torch.allclose(metrics["cross_entropy"], ref_ce)
    assert torch.allclose(metrics["mean_dice_loss"], ref_dice) and not metrics["loss"].requires_grad
    (3.0 * loss).backward()
    (3.0 * ref_loss).backward()
    assert torch.allclose(logits.grad, ref.grad, atol=1e-7)


def test_segmentation_loss_rejects_bad_shapes(emulated_kernels):
    logits, labels = _case(shape=(2, 4, 8, 8), ignore=False)
    with pytest.raises(ValueError):
        segmentation_loss(logits, labels.squeeze(1))
    with pytest.raises(ValueError):
        segmentation_loss(logits, labels[:, :, :4])
