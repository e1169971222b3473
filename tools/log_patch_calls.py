"""Print the shapes / strides of every gather_patches / scatter_patches call of one MAE step at the bench configuration."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import model_kwargs, synthetic_batch  # noqa: E402
from cinema_b200 import CineMA, _C  # noqa: E402
from cinema_b200.train import MAETrainer  # noqa: E402

kw = model_kwargs("base", (192, 192, 16), (192, 192))
dev = torch.device("cuda")
model = CineMA(**kw).to(dev).train()
tr = MAETrainer(model, use_cuda_graph=False)
batch = {k: v.to(dev) for k, v in synthetic_batch(kw, 16, 0, False).items()}
tr.step(batch)
og, osc = _C.gather_patches, _C.scatter_patches


def g(src, grid, patch, idx, chan_last, out):
    print("gather ", tuple(src.shape), src.stride(), src.dtype, "grid", tuple(grid), "patch", tuple(patch), "idx", None if idx is None else tuple(idx.shape),
          "chan_last", chan_last, "out", tuple(out.shape))
    return og(src, grid, patch, idx, chan_last, out)


def s(rows, dst, grid, patch, idx, chan_last, accumulate=False):
    print("scatter", tuple(dst.shape), dst.stride(), dst.dtype, "grid", tuple(grid), "patch", tuple(patch), "idx", None if idx is None else tuple(idx.shape),
          "chan_last", chan_last, "rows", tuple(rows.shape), rows.dtype, "acc", accumulate)
    return osc(rows, dst, grid, patch, idx, chan_last, accumulate)


_C.gather_patches, _C.scatter_patches = g, s
tr.eager_step(batch)
torch.cuda.synchronize()
