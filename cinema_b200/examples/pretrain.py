"""MAE pre-training loop, the counterpart of ``cinema/examples/train/pretrain.py:139-225`` on the B200 path.

    python -m cinema_b200.examples.pretrain <config.yaml> [--shards DIR]      # one process per GPU: launch with torchrun

``config`` has the reference example's keys (``cinema/examples/train/pretrain.yaml``: seed, logging.dir, data.{sax,lax}.
patch_size / in_chans, train.{lr, min_lr, n_warmup_epochs, n_epochs, batch_size, enc_mask_ratio, clip_grad, weight_decay,
betas}, model.{size, patch_size, scale_factor, enc_conv_chans, enc_conv_n_blocks}) plus ``data.shard_dir`` (a directory
written by ``cinema_b200.data.write_shards``; the reference reads NIfTI files found through manifest CSVs instead).

What maps to what:

    reference                                              here
    -----------------------------------------------------  ----------------------------------------------------------
    DataLoader(UKBDataset, RandomSampler, pin_memory)      CineShardDataset -> ShardSampler -> FrameBatcher -> DeviceFeeder
    adjust_learning_rate(optimizer, i / len + epoch, ...)  trainer.set_lr(cosine_lr(i / len + epoch, ...))
    autocast forward, GradScaler backward, clip, AdamW     trainer.step(batch)  (one fused call, CUDA-graph replay)
    isnan(loss) -> skip the batch                          the AdamW kernel skips the update on a non-finite gradient norm
    save_file(model.state_dict(), ckpt_{epoch}.safetensors) the same file (same keys), plus trainer state for resuming
"""

from __future__ import annotations

import argparse
import logging
import os
from pathlib import Path

import torch

from cinema_b200 import data as D
from cinema_b200.mae import _Cfg, get_model
from cinema_b200.train import MAETrainer, cosine_lr

logger = logging.getLogger(__name__)


def get_feeder(config, device, rank: int = 0, world: int = 1) -> D.DeviceFeeder:
    """Shards -> device batches (cinema/examples/train/pretrain.py:60-126 builds the DataLoader here)."""
    c = _Cfg(config)
    views = list(c.model.views) if not isinstance(c.model.views, str) else [c.model.views]
    sizes = {v: tuple(c.data.sax.patch_size if v == "sax" else c.data.lax.patch_size) for v in views}
    ds = D.CineShardDataset(c.data.shard_dir, views=views)
    sampler = D.ShardSampler(len(ds), rank=rank, world=world, seed=c.seed)
    tr = config.get("transform") if isinstance(config, dict) else getattr(config, "transform", None)
    zoom_prob = float((tr.get("prob", 0.0) if isinstance(tr, dict) else getattr(tr, "prob", 0.0)) if tr is not None else 0.0)
    batcher = D.FrameBatcher(ds, sampler, batch_size=c.train.batch_size, image_size_dict=sizes, seed=c.seed + rank,
                             zoom_prob=zoom_prob)  # RandZoomd(prob=config.transform.prob), cinema/mae/pretrain.py:163-183
    return D.DeviceFeeder(batcher, device)


def pretrain_one_epoch(trainer: MAETrainer, feeder: D.DeviceFeeder, config, epoch: int, log_every: int = 50) -> list[float]:
    """One pass over the feeder (cinema/examples/train/pretrain.py:139-188).  The loss is read back every ``log_every``
    steps only: nothing else in the loop synchronises with the device."""
    c = _Cfg(config)
    feeder.batcher.sampler.set_epoch(epoch)
    n = max(len(feeder.batcher), 1)
    losses = []
    for i, batch in enumerate(feeder):
        lr = cosine_lr(step=i / n + epoch, warmup_steps=c.train.n_warmup_epochs, max_n_steps=c.train.n_epochs, lr=c.train.lr,
                       min_lr=c.train.min_lr)
        trainer.set_lr(lr)
        loss = trainer.step(batch)
        if i % log_every == 0 or i == n - 1:
            losses.append(float(loss))
            logger.info("epoch %d step %d/%d: loss %.5f lr %.3e grad_norm %.4f", epoch, i, n, losses[-1], lr,
                        float(trainer.opt.grad_norm(1.0 / (trainer.world * trainer.n_accum))))
    return losses


def run(config, resume: str | os.PathLike | None = None) -> MAETrainer:
    """Build model / trainer / feeder from ``config`` and train ``config.train.n_epochs`` epochs, writing
    ``<logging.dir>/ckpt/ckpt_{epoch}.safetensors`` (model, reference keys) and ``trainer_{epoch}.pt`` (model + optimiser)."""
    from safetensors.torch import save_file

    c = _Cfg(config)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if torch.cuda.is_available():
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        torch.cuda.set_device(device)
        if world > 1 and not torch.distributed.is_initialized():
            torch.distributed.init_process_group("nccl", device_id=device)
    else:
        device = torch.device("cpu")
    torch.manual_seed(c.seed + rank)  # cinema/mae/pretrain.py:309-310
    ckpt_dir = Path(c.logging.dir) / "ckpt"
    ckpt_dir.mkdir(parents=True, exist_ok=True)
    model = get_model(config).to(device)
    model.train(True)
    trainer = MAETrainer(model, lr=c.train.lr, betas=tuple(c.train.betas), weight_decay=c.train.weight_decay,
                         clip_grad=c.train.clip_grad if c.train.clip_grad > 0 else None, enc_mask_ratio=c.train.enc_mask_ratio)
    start = 0
    if resume is not None:
        state = torch.load(resume, map_location=device)
        trainer.load_state_dict(state)
        start = int(state.get("epoch", -1)) + 1
    feeder = get_feeder(config, device, rank, world)
    for epoch in range(start, c.train.n_epochs):
        pretrain_one_epoch(trainer, feeder, config, epoch)
        if rank == 0:
            save_file({k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}, str(ckpt_dir / f"ckpt_{epoch}.safetensors"))
            torch.save({**trainer.state_dict(), "epoch": epoch}, ckpt_dir / f"trainer_{epoch}.pt")
            logger.info("Saved checkpoint of epoch %d at %s.", epoch, ckpt_dir)
    return trainer


def main() -> None:
    import yaml

    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("config")
    ap.add_argument("--shards", default=None, help="overrides data.shard_dir")
    ap.add_argument("--resume", default=None, help="trainer_{epoch}.pt to continue from")
    a = ap.parse_args()
    logging.basicConfig(level=logging.INFO)
    with open(a.config, encoding="utf-8") as f:
        config = yaml.safe_load(f)
    if a.shards:
        config["data"]["shard_dir"] = a.shards
    run(config, resume=a.resume)


if __name__ == "__main__":
    main()
