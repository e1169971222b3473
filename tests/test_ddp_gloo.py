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
    orig = model.forward
    model.forward = lambda image_dict, ratio: orig(image_dict, ratio, enc_mask_dict=masks)
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
