"""Input pipeline (cinema_b200/data.py; reference cinema/mae/pretrain.py:88-200,322-338), CPU only: shard round trip,
sampler == torch's DistributedSampler index for index, frame selection / end padding of the batcher, ScaleIntensity
semantics, and the double-buffered feeder's ordering."""

import numpy as np
import pytest
import torch
from torch.utils.data import DistributedSampler

from cinema_b200 import data as D

VIEWS = ("sax", "lax_2c", "lax_3c", "lax_4c")


def _subjects(n, t=6, seed=0):
    rng = np.random.default_rng(seed)
    for i in range(n):
        nt = t if i != 1 else 3  # one short study: exercises the middle-frame fallback
        sax = rng.integers(-50, 2000, size=(20, 24, 5, nt)).astype(np.int16)
        lax = {v: rng.integers(0, 256, size=(28, 30, nt)).astype(np.uint8) for v in VIEWS[1:]}
        for k in range(nt):  # tag every frame with its index so that tests can read back which frame was chosen
            sax[0, 0, 0, k] = 1000 + k
            for v in lax:
                lax[v][0, 0, k] = 100 + k
        yield f"s{i:03d}", {"sax": sax, **lax}


@pytest.fixture(scope="module")
def shard_root(tmp_path_factory):
    root = tmp_path_factory.mktemp("shards")
    D.write_shards(root, _subjects(7))
    return root


def test_shard_round_trip(shard_root):
    ds = D.CineShardDataset(shard_root)
    assert len(ds) == 7 and ds.views == list(VIEWS)
    ref = dict(_subjects(7))
    for i in (0, 1, 6):
        for v in VIEWS:
            src = ref[f"s{i:03d}"][v]
            nt = src.shape[-1]
            assert ds.n_frames(i, v) == nt
            for t in (0, nt - 1):
                frame, mn, mx = ds.frame(i, v, t)
                assert frame.dtype == src.dtype and np.array_equal(frame, src[..., t])
                assert mn == float(src[..., t].min()) and mx == float(src[..., t].max())
    # t beyond the study -> the middle frame (cinema/mae/pretrain.py:107-108)
    frame, _, _ = ds.frame(1, "sax", 40)
    assert np.array_equal(frame, ref["s001"]["sax"][..., 3 // 2])
    assert len(D.CineShardDataset(shard_root, views=["sax"], max_n_samples=3)) == 3
    with pytest.raises(ValueError):
        D.CineShardDataset(shard_root, views=["sax", "lax_5c"])


@pytest.mark.parametrize(("n", "world"), [(7, 2), (16, 4), (5, 8), (100, 3), (1, 2)])
def test_sampler_equals_torch_distributed_sampler(n, world):
    data = list(range(n))
    for seed in (0, 3):
        for epoch in (0, 1, 5):
            for rank in range(world):
                ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=seed)
                ref.set_epoch(epoch)
                mine = D.ShardSampler(n, rank=rank, world=world, seed=seed)
                mine.set_epoch(epoch)
                assert list(mine) == list(ref) and len(mine) == len(ref)
    ref = DistributedSampler(data, num_replicas=world, rank=0, shuffle=False)
    assert list(D.ShardSampler(n, 0, world, shuffle=False)) == list(ref)
    one = D.ShardSampler(n, 0, 1, seed=1)
    assert sorted(one) == data  # single process: a permutation of everything
    with pytest.raises(ValueError):
        D.ShardSampler(n, rank=world, world=world)


def test_batcher_frames_padding_and_drop_last(shard_root):
    ds = D.CineShardDataset(shard_root)
    sizes = {"sax": (24, 24, 8), "lax_2c": (32, 32), "lax_3c": (32, 32), "lax_4c": (32, 32)}
    sampler = D.ShardSampler(len(ds), seed=2)
    bat = D.FrameBatcher(ds, sampler, batch_size=3, image_size_dict=sizes, n_frames=6, seed=5, pin_memory=False)
    assert len(bat) == 2  # 7 // 3, drop_last
    order = sampler.indices()
    ref = dict(_subjects(7))
    n_batches = 0
    for k, raw in enumerate(bat):
        n_batches += 1
        assert raw.images["sax"].shape == (3, 1, 24, 24, 8) and raw.images["sax"].dtype == torch.int16
        assert raw.images["lax_3c"].shape == (3, 1, 32, 32) and raw.images["lax_3c"].dtype == torch.uint8
        for j in range(3):
            idx = order[k * 3 + j]
            nt = ref[f"s{idx:03d}"]["sax"].shape[-1]
            t = int(raw.images["sax"][j, 0, 0, 0, 0]) - 1000  # the frame tag
            assert 0 <= t < nt
            for v in VIEWS[1:]:  # the SAME frame index for every view of a sample
                assert int(raw.images[v][j, 0, 0, 0]) - 100 == t
            src = ref[f"s{idx:03d}"]["sax"][..., t]
            img = raw.images["sax"][j, 0].numpy()
            assert np.array_equal(img[:20, :24, :5], src)
            assert float(raw.lo["sax"][j]) == float(src.min()) and float(raw.hi["sax"][j]) == float(src.max())
            # SpatialPad(method="end"): everything outside the frame holds the value that scales to 0
            pad = np.ones(img.shape, dtype=bool)
            pad[:20, :24, :5] = False
            assert (img[pad] == src.min()).all()
            scaled = D.scale_intensity(raw.images["sax"][j:j + 1], raw.lo["sax"][j:j + 1], raw.hi["sax"][j:j + 1])[0, 0].numpy()
            assert (scaled[pad] == 0).all() and scaled.min() == 0.0 and abs(scaled.max() - 1.0) < 1e-6
    assert n_batches == 2
    with pytest.raises(ValueError):  # nothing in the reference pipeline crops: a frame larger than the patch is an error
        next(iter(D.FrameBatcher(ds, sampler, 2, {**sizes, "sax": (16, 24, 8)}, pin_memory=False)))


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32])
def test_scale_intensity_matches_monai_definition(dtype):
    """ScaleIntensity(minv=0, maxv=1): (img - min) / (max - min); constant image -> img * minv = 0 (MONAI rescale_array)."""
    rng = np.random.default_rng(0)
    x = rng.integers(0, 200, size=(3, 1, 6, 7, 4)).astype(dtype)
    x[2] = 17  # constant sample
    lo = torch.tensor([float(v.min()) for v in x])
    hi = torch.tensor([float(v.max()) for v in x])
    got = D.scale_intensity(torch.from_numpy(x), lo, hi)
    assert got.dtype == torch.float32
    for i in range(2):
        ref = (x[i].astype(np.float64) - x[i].min()) / (x[i].max() - x[i].min())
        np.testing.assert_allclose(got[i].numpy(), ref, rtol=1e-6, atol=1e-7)
    assert (got[2] == 0).all()
    out = torch.empty(x.shape)
    assert D.scale_intensity(torch.from_numpy(x), lo, hi, out=out) is out and torch.equal(out, got)


def test_device_feeder_yields_every_batch_in_order(shard_root):
    ds = D.CineShardDataset(shard_root)
    sizes = {"sax": (24, 24, 8), "lax_2c": (32, 32), "lax_3c": (32, 32), "lax_4c": (32, 32)}
    mk = lambda: D.FrameBatcher(ds, D.ShardSampler(len(ds), seed=4), 2, sizes, n_frames=6, seed=9, pin_memory=False)  # noqa: E731
    direct = [{v: D.scale_intensity(r.images[v], r.lo[v], r.hi[v]).clone() for v in r.images} for r in mk()]
    fed = [{v: t.clone() for v, t in b.items()} for b in D.DeviceFeeder(mk(), "cpu")]
    assert len(direct) == len(fed) == 3
    for a, b in zip(direct, fed):
        for v in a:
            assert b[v].dtype == torch.float32 and torch.equal(a[v], b[v])
            assert float(b[v].min()) == 0.0 and float(b[v].max()) <= 1.0 + 1e-6


def test_zoom_identity_and_keep_size_restatement():
    """``zoom_intensity`` (the torch restatement of RandZoom(keep_size) -> ScaleIntensity -> SpatialPad): factor 1 equals
    plain ScaleIntensity of the frame, zoom-out centre-pads with zeros, zoom-in keeps the centre, the model-size padding
    stays 0, and every sample lands in [0, 1] with both ends reached."""
    g = torch.Generator().manual_seed(0)
    raw = torch.zeros(3, 1, 24, 20, 6, dtype=torch.int16)
    ext = torch.tensor([[20, 18, 5], [24, 20, 6], [16, 16, 4]], dtype=torch.int32)
    for j in range(3):
        e = ext[j].tolist()
        raw[j, 0, :e[0], :e[1], :e[2]] = torch.randint(5, 900, tuple(e), generator=g, dtype=torch.int16)
    zoom = torch.tensor([1.0, 0.93, 1.1], dtype=torch.float32)
    out = D.zoom_intensity(raw, ext, zoom)
    f0 = raw[0, 0, :20, :18, :5].float()
    assert torch.equal(out[0, 0, :20, :18, :5], (f0 - f0.min()) / (f0.max() - f0.min()))
    for j in range(3):
        e = ext[j].tolist()
        inside = out[j, 0, :e[0], :e[1], :e[2]]
        assert float(inside.min()) == 0.0 and abs(float(inside.max()) - 1.0) < 1e-6
        assert float(out[j].sum()) == pytest.approx(float(inside.sum()))  # nothing outside the frame
    # zoom 0.93 of a 24 x 20 x 6 frame: 22 x 18 x 5 voxels centred -> pads (1, 1), (1, 1), (0, 1): the border planes hold the
    # zoomed frame's zero padding, which is also its minimum (raw values are >= 5), i.e. ScaleIntensity's 0
    z = out[1, 0]
    assert float(z[0].abs().max()) == 0.0 and float(z[23].abs().max()) == 0.0 and float(z[:, :, 5].abs().max()) == 0.0
    assert float(z[:, 0].abs().max()) == 0.0 and float(z[:, 19].abs().max()) == 0.0
    assert float(z[1:23, 1:19, :5].min()) > 0.0


def test_batcher_draws_zoom_factors_and_feeder_applies_them(shard_root):
    ds = D.CineShardDataset(shard_root)
    sizes = {"sax": (24, 24, 8), "lax_2c": (32, 32), "lax_3c": (32, 32), "lax_4c": (32, 32)}
    mk = lambda p: D.FrameBatcher(ds, D.ShardSampler(len(ds), seed=4), 2, sizes, n_frames=6, seed=9, pin_memory=False, zoom_prob=p)  # noqa: E731
    raws = [(r.zoom["sax"].clone(), r.zoom["lax_2c"].clone(), r.zoom["lax_3c"].clone(), r.extent["sax"].clone(),
             r.extent["lax_4c"].clone()) for r in mk(1.0)]
    assert len(raws) == 3
    for zs, z2, z3, es, el in raws:
        assert torch.equal(z2, z3)                                     # one draw for the LAX views of a sample
        assert ((zs >= 0.9) & (zs <= 1.1)).all() and not torch.equal(zs, z2)
        assert es.tolist() == [[20, 24, 5]] * 2 and el.tolist() == [[28, 30, 1]] * 2
    assert all(float(r.zoom["sax"].min()) == 1.0 == float(r.zoom["sax"].max()) for r in mk(1e-9))  # prob ~ 0: never zoomed
    assert next(iter(mk(0.0))).zoom is None
    fed = [{v: t.clone() for v, t in b.items()} for b in D.DeviceFeeder(mk(1.0), "cpu")]  # (the feeder recycles its slots)
    direct = [{v: D.zoom_intensity(r.images[v], r.extent[v], r.zoom[v]).clone() for v in r.images} for r in mk(1.0)]
    for a, b in zip(direct, fed):
        for v in a:
            assert torch.equal(a[v], b[v]) and float(b[v].min()) == 0.0 and float(b[v].max()) <= 1.0 + 1e-6


def test_pipeline_feeds_the_model(shard_root, emulated_kernels, golden_dir):
    """The feeder's batches go straight into the model (host logic through the emulated kernels)."""
    from cinema_b200 import CineMA

    g = torch.load(golden_dir / "mae_small_4view.pt")
    ds = D.CineShardDataset(shard_root)
    sizes = g["kw"]["image_size_dict"]  # SAX (64, 64, 2), LAX (64, 64): frames (20, 24, 5) do not fit the SAX depth
    with pytest.raises(ValueError):
        next(iter(D.FrameBatcher(ds, D.ShardSampler(len(ds)), 2, sizes, pin_memory=False)))
    lax_only = D.CineShardDataset(shard_root, views=["lax_2c", "lax_4c"])
    feeder = D.DeviceFeeder(D.FrameBatcher(lax_only, D.ShardSampler(len(lax_only)), 2, sizes, n_frames=6, seed=0, pin_memory=False), "cpu")
    model = CineMA(**g["kw"])
    model.load_state_dict(g["state_dict"])
    batch = next(iter(feeder))
    loss, preds, _, _ = model(batch, 0.75)
    assert torch.isfinite(loss) and set(preds) == {"lax_2c", "lax_4c"}


def test_example_pretrain_loop_trains_checkpoints_and_resumes(tmp_path, emulated_kernels):
    """cinema_b200/examples/pretrain.py (counterpart of cinema/examples/train/pretrain.py): shards -> feeder -> trainer with
    the cosine schedule, per-epoch safetensors checkpoint in the reference's key schema, resume from the trainer state."""
    from safetensors.torch import load_file

    from cinema_b200.examples import pretrain as ex

    rng = np.random.default_rng(1)
    shards = tmp_path / "shards"
    D.write_shards(shards, [(f"s{i}", {"sax": rng.integers(0, 900, size=(30, 32, 4, 5)).astype(np.int16),
                                       "lax_4c": rng.integers(0, 256, size=(32, 28, 5)).astype(np.uint8)}) for i in range(4)])
    config = {
        "seed": 0, "grad_ckpt": False, "logging": {"dir": str(tmp_path / "run")},
        "data": {"shard_dir": str(shards), "sax": {"patch_size": [32, 32, 4], "in_chans": 1},
                 "lax": {"patch_size": [32, 32], "in_chans": 1}},
        "train": {"clip_grad": 5.0, "weight_decay": 0.05, "betas": [0.9, 0.95], "lr": 1e-3, "min_lr": 1e-6, "n_warmup_epochs": 1,
                  "n_epochs": 2, "batch_size": 2, "enc_mask_ratio": 0.75},
        "model": {"size": "tiny", "views": ["sax", "lax_4c"], "patch_size": [4, 4, 1], "scale_factor": [2, 2, 1],
                  "enc_conv_chans": [8, 16], "enc_conv_n_blocks": 1},
    }
    # the reference's get_model always builds the four UK Biobank views; the loader feeds the subset named in model.views
    trainer = ex.run(config)
    assert trainer.opt.t == 4  # 2 epochs x (4 subjects // batch 2)
    ckpt = tmp_path / "run" / "ckpt"
    sd = load_file(str(ckpt / "ckpt_1.safetensors"))
    assert set(sd.keys()) == set(trainer.model.state_dict().keys())  # (safetensors stores keys sorted)
    assert all(torch.equal(sd[k], v.cpu()) for k, v in trainer.model.state_dict().items())
    # resume after epoch 0 and run epoch 1 again: same optimiser step count, finite weights, a different (continued) state
    resumed = ex.run(config, resume=ckpt / "trainer_0.pt")
    assert resumed.opt.t == 4
    assert all(torch.isfinite(p).all() for p in resumed.model.parameters())
    first = load_file(str(ckpt / "ckpt_0.safetensors"))
    moved = sum(float((sd[k] - first[k]).abs().sum()) for k in sd if sd[k].is_floating_point())
    assert moved > 0
