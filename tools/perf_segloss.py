"""Fused segmentation loss (cb_seg_loss_fwd / _bwd) against the stock torch ops of cinema/segmentation/train.py:77-103 on the
ACDC shape of BASELINE.json config 4 (B = 4, 4 classes, 192 x 192 x 16), forward + backward.   python tools/perf_segloss.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200.segmentation.loss import segmentation_loss, segmentation_loss_restated  # noqa: E402

DEV = "cuda"
for dt in (torch.float32, torch.bfloat16):
    logits = (torch.randn(4, 4, 192, 192, 16, device=DEV) * 2).to(dt).requires_grad_(True)
    labels = torch.randint(-1, 4, (4, 1, 192, 192, 16), device=DEV)

    def ours():
        logits.grad = None
        loss, _ = segmentation_loss(logits, labels)
        loss.backward()

    def stock():
        logits.grad = None
        loss, _, _ = segmentation_loss_restated(logits, labels)
        loss.backward()

    for name, fn in (("fused kernels", ours), ("torch ops (restatement of the stock path)", stock)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        n = logits.numel()
        by = n * logits.element_size() * 3 + labels.numel() * 8 * 2  # logits read twice + gradient written, labels read twice
        t = sorted(ts)[3]
        print(f"{str(dt):16s} {name:44s} {t * 1e3:8.1f} us   {by / t / 1e6:7.0f} GB/s algorithmic")
