"""World-size-2 data-parallel step over ``gloo`` on CPU (kernels emulated, tests/emu_c.py): the flat gradient
all-reduce + fused optimiser must (a) leave both ranks with identical parameters and (b) equal the single-process
step on the concatenated batch -- the semantics of DDP's gradient averaging (cinema/device.py:101-103)."""

import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden" / "mae_small_4view.pt"


def _patch():
    sys.path.insert(0, str(ROOT))
    from cinema_b200 import _C, engine
    from tests import emu_c

    for name in emu_c.ALL:
        setattr(_C, name, getattr(emu_c, name))
    engine.check_head_dim = lambda d: None


def _one_step(images, masks, world):
    from cinema_b200 import CineMA
    from cinema_b200.train import MAETrainer

    g = torch.load(GOLDEN)
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.train()
    tr = MAETrainer(model, lr=1e-3, use_cuda_graph=False, overlap_allreduce=True)  # bucketed path (no-op at world 1)
    # fixed masks so that both layouts see the same problem
    orig, orig_ts = model.forward, model.train_step
    model.forward = lambda image_dict, ratio: orig(image_dict, ratio, enc_mask_dict=masks)
    model.train_step = lambda image_dict, ratio: orig_ts(image_dict, ratio, enc_mask_dict=masks)
    loss = tr.step(images)
    return float(loss), tr.arena.flat32.clone(), float(tr.opt.grad_norm(1.0 / world)), tr.arena.gflat.clone() / world


def _worker(rank, world, port, out):
    _patch()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = torch.load(GOLDEN)
    images = {k: v[rank:rank + 1] for k, v in g["images"].items()}
    masks = {k: v[rank:rank + 1] for k, v in g["masks"].items()}
    loss, flat, gn, grad = _one_step(images, masks, world)
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"loss": loss, "flat": flat, "same": bool(torch.equal(gathered[0], gathered[1])), "gn": gn, "grad": grad}, out)
    dist.destroy_process_group()


def test_two_rank_step_equals_full_batch_step(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "r0.pt"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    res = torch.load(out)
    assert res["same"], "ranks diverged after the all-reduced update"
    _patch()
    g = torch.load(GOLDEN)
    loss, flat, gn, grad = _one_step(g["images"], g["masks"], 1)
    # batch of 2 in one process == 2 ranks x 1 sample: same averaged gradient (fp32 summation order differs) ...
    assert abs(gn - res["gn"]) < 1e-2 * gn
    assert float((grad - res["grad"]).norm() / grad.norm()) < 1e-2  # bf16 roundings differ (per-rank loss scale)
    # ... and the same clipped AdamW update (the first Adam step is sign-like, so compare norm-wise)
    rel = float((flat - res["flat"]).norm() / (flat - g_flat_initial(g)).norm())
    assert rel < 0.05


def g_flat_initial(g):
    _patch()
    from cinema_b200 import CineMA
    from cinema_b200.arena import ensure_arena

    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    return ensure_arena(model).flat32.clone()


def test_allreduce_buckets_partition_the_arena():
    """The overlapped all-reduce sends three buckets (decoder subtree, upper encoder half, rest): their flat ranges must
    be disjoint and cover every trainable parameter exactly once."""
    _patch()
    from cinema_b200 import CineMA
    from cinema_b200.arena import ensure_arena
    from cinema_b200.mae import grad_stages

    g = torch.load(GOLDEN)
    model = CineMA(**g["kw"])
    arena = ensure_arena(model)
    stages = grad_stages(model)
    covered = {id(p) for ps in stages.values() for p in ps}
    ranges = [r for ps in stages.values() for r in arena.ranges_of(ps)]
    ranges += arena.ranges_of([p for p in arena.params if id(p) not in covered])
    hit = torch.zeros(arena.numel, dtype=torch.int32)
    for s, e in ranges:
        hit[s:e] += 1
    assert int(hit.max()) == 1, "buckets overlap"
    for p in arena.params:
        off = arena._off[id(p)]
        assert bool((hit[off:off + p.numel()] == (1 if p.requires_grad else 0)).all()), arena._names[id(p)]
    # a module subtree is a handful of contiguous runs (one per optimiser category, registration-order layout)
    assert len(arena.ranges_of(stages["decoder"])) <= 3


# ------------------------------------------------------------------------------------------
# single-process trainer logic (kernels emulated): gradient accumulation, checkpoint round trip
# ------------------------------------------------------------------------------------------
def _trainer(n_accum=1, lr=1e-3):
    from cinema_b200 import CineMA
    from cinema_b200.train import MAETrainer

    g = torch.load(GOLDEN)
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.train()
    tr = MAETrainer(model, lr=lr, use_cuda_graph=False, n_accum_steps=n_accum)
    orig, orig_ts = model.forward, model.train_step
    state = {"masks": None}
    model.forward = lambda image_dict, ratio: orig(image_dict, ratio, enc_mask_dict=state["masks"])
    model.train_step = lambda image_dict, ratio: orig_ts(image_dict, ratio, enc_mask_dict=state["masks"])
    return g, tr, state


def test_gradient_accumulation_equals_full_batch_step():
    """Two micro-steps of one sample each (n_accum_steps = 2) = one step on both samples: same mean gradient, same update
    (cinema/mae/pretrain.py:258-267), and no optimiser update on the first micro-step."""
    _patch()
    g, tr_full, st_full = _trainer(1)
    st_full["masks"] = g["masks"]
    tr_full.step(g["images"])
    assert tr_full.updated
    _, tr_acc, st_acc = _trainer(2)
    before = tr_acc.arena.flat32.clone()
    for i in range(2):
        st_acc["masks"] = {k: v[i:i + 1] for k, v in g["masks"].items()}
        tr_acc.step({k: v[i:i + 1] for k, v in g["images"].items()})
        if i == 0:
            assert not tr_acc.updated and torch.equal(tr_acc.arena.flat32, before) and tr_acc.opt.t == 0
    assert tr_acc.updated and tr_acc.opt.t == 1
    # per-sample losses are means over that sample's masked patches -> the full-batch gradient is the mean of the two
    gfull, gacc = tr_full.arena.gflat, tr_acc.arena.gflat / 2
    assert float((gfull - gacc).norm() / gfull.norm()) < 2e-2  # bf16 rounding of different batch shapes
    assert abs(float(tr_full.opt.grad_norm(1.0)) - float(tr_acc.opt.grad_norm(0.5))) < 2e-2 * float(tr_full.opt.grad_norm(1.0))
    d_full, d_acc = tr_full.arena.flat32 - before, tr_acc.arena.flat32 - before
    assert float((d_full - d_acc).norm() / d_full.norm()) < 5e-2
    # the next update starts from a cleared accumulator
    st_acc["masks"] = {k: v[:1] for k, v in g["masks"].items()}
    tr_acc.step({k: v[:1] for k, v in g["images"].items()})
    one = tr_acc.arena.gflat.clone()
    _, tr_one, st_one = _trainer(1)
    st_one["masks"] = st_acc["masks"]
    tr_one.model.load_state_dict(tr_acc.model.state_dict())
    tr_one.step({k: v[:1] for k, v in g["images"].items()})
    assert float((one - tr_one.arena.gflat).norm() / one.norm()) < 1e-5


def test_trainer_checkpoint_round_trip():
    """state_dict() -> fresh trainer -> load_state_dict(): the continued run is identical to the uninterrupted one."""
    _patch()
    g, tr_a, st_a = _trainer(1)
    st_a["masks"] = g["masks"]
    for _ in range(2):
        tr_a.step(g["images"])
    ckpt = tr_a.state_dict()
    assert set(ckpt) == {"model", "optimizer"} and ckpt["optimizer"]["step"] == 2
    assert set(ckpt["optimizer"]["exp_avg"]) == {n for n, p in tr_a.model.named_parameters() if p.requires_grad}
    tr_a.step(g["images"])
    _, tr_b, st_b = _trainer(1)
    st_b["masks"] = g["masks"]
    tr_b.load_state_dict(ckpt)
    assert tr_b.opt.t == 2
    tr_b.step(g["images"])
    assert torch.equal(tr_a.arena.flat32, tr_b.arena.flat32)
    assert torch.equal(tr_a.opt.m, tr_b.opt.m) and torch.equal(tr_a.opt.v, tr_b.opt.v)


def test_get_n_accum_steps_and_cosine_lr():
    import math

    import pytest

    from cinema_b200.train import cosine_lr, get_n_accum_steps

    assert get_n_accum_steps(256, 16, 8) == 2 and get_n_accum_steps(128, 16, 8) == 1
    with pytest.raises(ValueError):
        get_n_accum_steps(64, 16, 8)
    with pytest.raises(ValueError):
        get_n_accum_steps(200, 16, 8)
    # cinema/optim.py:21-52: linear warm-up, half-cycle cosine to min_lr
    assert cosine_lr(0, 10, 100, 1e-3, 1e-5) == 0.0
    assert abs(cosine_lr(5, 10, 100, 1e-3, 1e-5) - 5e-4) < 1e-12
    assert abs(cosine_lr(10, 10, 100, 1e-3, 1e-5) - 1e-3) < 1e-12
    assert abs(cosine_lr(55, 10, 100, 1e-3, 1e-5) - (1e-5 + (1e-3 - 1e-5) * 0.5 * (1 + math.cos(math.pi * 0.5)))) < 1e-12
    assert abs(cosine_lr(100, 10, 100, 1e-3, 1e-5) - 1e-5) < 1e-12


def _helper_worker(rank, world, port, out):
    _patch()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from cinema_b200 import CineMA
    from cinema_b200.train import allreduce_gradients

    g = torch.load(GOLDEN)
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    model.train()
    images = {k: v[rank:rank + 1] for k, v in g["images"].items()}
    masks = {k: v[rank:rank + 1] for k, v in g["masks"].items()}
    loss, _, _, _ = model(images, g["ratio"], enc_mask_dict=masks)
    loss.backward()
    local = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    allreduce_gradients(model)
    mean = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        want = {k: (gathered[0][k] + gathered[1][k]) / 2 for k in local}
        worst = max(float((mean[k] - want[k]).abs().max()) for k in want)
        torch.save({"worst": worst, "n": len(want)}, out)
    dist.destroy_process_group()


def test_allreduce_gradients_helper(tmp_path):
    """Eager data-parallel loops: ``allreduce_gradients`` after ``backward`` leaves every ``p.grad`` equal to
    the mean of the ranks' local gradients -- DDP's semantics without the DDP wrapper."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "h.pt"
    mp.spawn(_helper_worker, args=(2, port, str(out)), nprocs=2, join=True)
    res = torch.load(out)
    assert res["n"] > 100 and res["worst"] < 1e-6, res
