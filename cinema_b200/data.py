"""Input pipeline for MAE pre-training at B200 step rates (reference:
``cinema/mae/pretrain.py:88-200,322-338``).

The reference decodes one time frame per view from gzip NIfTI with SimpleITK in 16 DataLoader workers per GPU, runs the
MONAI transforms (RandZoom, ScaleIntensity, SpatialPad) on the CPU in fp32 and ships fp32 batches.  At ~560 frame-sets per
second and GPU that is 25 MB/ms of fp32 host traffic after a gzip decode per sample -- the loader, not the step, becomes
the limit.  Here the cine studies are decoded ONCE into frame-major integer shards; per step the host only picks frames,
copies them (raw dtype, 1-2 bytes per voxel instead of 4) into a pinned staging batch and uploads it on a copy stream
behind the running step; intensity scaling happens on the device.

    shard format   ``<root>/index.json`` + one ``.npy`` per (subject, view), shape (T, *spatial), integer or float dtype,
                   memory-mappable; the index stores per-frame min / max so that ScaleIntensity needs no host pass
    sampler        ``ShardSampler``: the index sequence of ``torch.utils.data.DistributedSampler(shuffle=True)`` /
                   a seeded ``RandomSampler``, bit for bit (same permutation, padding and rank striding)
    batcher        ``FrameBatcher``: random frame per sample (``rng.integers(n_frames)``, middle frame if the study is
                   shorter, cinema/mae/pretrain.py:107-108,146), SpatialPad(method="end") to the model's patch size with the
                   value that scales to 0, ``drop_last`` batches
    device side    ``scale_intensity``: MONAI ``ScaleIntensity(minv=0, maxv=1)`` = (x - min) / (max - min), constant images
                   map to 0; ``DeviceFeeder`` double-buffers raw batches through pinned memory and a copy stream
    augmentation   ``RandZoomd`` (prob ``config.transform.prob``, zoom in [0.9, 1.1], keep_size; trilinear for SAX, bicubic
                   for LAX, one draw shared by the LAX views; cinema/mae/pretrain.py:163-183): the batcher only DRAWS the
                   per-sample factors, the resampling runs on the device fused with ScaleIntensity and the end padding
                   (``cb_zoom_intensity``: zoom the unpadded frame about its centre, min / max of the zoomed frame, scale,
                   zeros outside the frame).  MONAI is not installed here, so this is pinned to MONAI's published ``Zoom``
                   algorithm restated over ``torch.nn.functional.interpolate`` (``zoom_intensity`` below), not to a MONAI run;
                   the reference's ``lazy=True`` composition resamples once through an affine grid instead and differs at
                   the borders.

Not reproduced: the NIfTI decoder itself (SimpleITK is not available here; ``write_shards`` takes arrays).
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Iterable, Iterator

import numpy as np
import torch

UKB_N_FRAMES = 50  # cinema/__init__.py: every UK Biobank cine has 50 frames


# ------------------------------------------------------------------------------------------
# shards
# ------------------------------------------------------------------------------------------
def write_shards(root: str | Path, subjects: Iterable[tuple[str, dict[str, np.ndarray]]]) -> Path:
    """Decode-once storage.  ``subjects`` yields (id, {view: array (*spatial, T)}) in the reference's axis order
    (x, y[, z], t) (cinema/mae/pretrain.py:88-118); frames are stored time-major so that one frame is one contiguous read."""
    root = Path(root)
    root.mkdir(parents=True, exist_ok=True)
    index = {"version": 1, "subjects": []}
    for sid, views in subjects:
        entry = {"id": str(sid), "views": {}}
        for view, arr in views.items():
            frames = np.ascontiguousarray(np.moveaxis(np.asarray(arr), -1, 0))
            name = f"{sid}_{view}.npy"
            np.save(root / name, frames)
            flat = frames.reshape(frames.shape[0], -1)
            entry["views"][view] = {"file": name, "shape": list(frames.shape), "dtype": str(frames.dtype),
                                    "frame_min": [float(x) for x in flat.min(axis=1)],
                                    "frame_max": [float(x) for x in flat.max(axis=1)]}
        index["subjects"].append(entry)
    with open(root / "index.json", "w", encoding="utf-8") as f:
        json.dump(index, f)
    return root / "index.json"


class CineShardDataset:
    """Memory-mapped frame access over ``write_shards`` output."""

    def __init__(self, root: str | Path, views: Iterable[str] | None = None, max_n_samples: int = 0) -> None:
        self.root = Path(root)
        with open(self.root / "index.json", encoding="utf-8") as f:
            index = json.load(f)
        if index.get("version") != 1:
            raise ValueError(f"unsupported shard index version {index.get('version')}")
        self.subjects = index["subjects"]
        if max_n_samples > 0:  # cinema/mae/pretrain.py:325-328
            self.subjects = self.subjects[:min(max_n_samples, len(self.subjects))]
        self.views = list(views) if views is not None else list(self.subjects[0]["views"].keys())
        for s in self.subjects:
            missing = [v for v in self.views if v not in s["views"]]
            if missing:
                raise ValueError(f"subject {s['id']} has no view(s) {missing}")
        self._maps: dict[tuple[int, str], np.ndarray] = {}

    def __len__(self) -> int:
        return len(self.subjects)

    def n_frames(self, index: int, view: str) -> int:
        return int(self.subjects[index]["views"][view]["shape"][0])

    def frame(self, index: int, view: str, t: int) -> tuple[np.ndarray, float, float]:
        """-> (frame (*spatial) view into the memory map, its min, its max); ``t`` beyond the study falls back to the
        middle frame like the reference loader (cinema/mae/pretrain.py:107-108)."""
        meta = self.subjects[index]["views"][view]
        n = meta["shape"][0]
        if t >= n:
            t = n // 2
        key = (index, view)
        arr = self._maps.get(key)
        if arr is None:
            arr = np.load(self.root / meta["file"], mmap_mode="r")
            if len(self._maps) >= 256:  # every map holds a file descriptor: stay far below the usual 1024 limit
                self._maps.clear()
            self._maps[key] = arr
        return arr[t], meta["frame_min"][t], meta["frame_max"][t]


# ------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------
class ShardSampler:
    """Per-rank index stream with the semantics of the samplers the reference uses (cinema/mae/pretrain.py:329-332):
    world > 1 -> ``DistributedSampler(dataset, num_replicas, rank, shuffle=True)`` (seed + epoch permutation, padded to a
    multiple of the world size by wrapping around, strided by rank; identical sequence); world == 1 -> a seeded random
    permutation per epoch."""

    def __init__(self, n: int, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True) -> None:
        if not 0 <= rank < world:
            raise ValueError(f"Invalid rank {rank}, rank should be in the interval [0, {world - 1}]")
        self.n, self.rank, self.world, self.seed, self.shuffle = n, rank, world, seed, shuffle
        self.epoch = 0
        self.num_samples = -(-n // world)
        self.total_size = self.num_samples * world

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self) -> int:
        return self.num_samples

    def indices(self) -> list[int]:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g).tolist()
        else:
            order = list(range(self.n))
        pad = self.total_size - len(order)
        if pad > 0:
            order += (order * (-(-pad // max(len(order), 1))))[:pad]
        return order[self.rank:self.total_size:self.world]

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices())


# ------------------------------------------------------------------------------------------
# host batches
# ------------------------------------------------------------------------------------------
class RawBatch:
    """One batch in the shards' dtype: ``images[view]`` (B, 1, *patch_size) pinned host tensors, ``lo`` / ``hi`` (B,) fp32
    per view = ScaleIntensity's min / max of each sample's frame.  With the zoom augmentation on: ``extent[view]`` (B, 3)
    int32 = the frame's own size inside the padded sample (unused axes 1) and ``zoom[view]`` (B,) fp32 = this sample's
    RandZoom factor (1 = not zoomed)."""

    def __init__(self, images: dict[str, torch.Tensor], lo: dict[str, torch.Tensor], hi: dict[str, torch.Tensor],
                 extent: dict[str, torch.Tensor] | None = None, zoom: dict[str, torch.Tensor] | None = None) -> None:
        self.images, self.lo, self.hi, self.extent, self.zoom = images, lo, hi, extent, zoom

    def nbytes(self) -> int:
        groups = [self.images, self.lo, self.hi] + [d for d in (self.extent, self.zoom) if d is not None]
        return sum(t.numel() * t.element_size() for d in groups for t in d.values())


class FrameBatcher:
    """dataset + sampler -> ``RawBatch`` stream (``drop_last``, one random frame per sample, padded at the END of every
    axis to the model's patch size with the value that ScaleIntensity maps to 0; a frame larger than the patch size is
    rejected, as nothing in the reference pipeline crops)."""

    def __init__(self, dataset: CineShardDataset, sampler: ShardSampler, batch_size: int,
                 image_size_dict: dict[str, tuple[int, ...]], n_frames: int = UKB_N_FRAMES, seed: int | None = None,
                 pin_memory: bool | None = None, n_buffers: int = 2, zoom_prob: float = 0.0,
                 zoom_range: tuple[float, float] = (0.9, 1.1)) -> None:
        """``zoom_prob`` > 0 switches the RandZoom augmentation on (``config.transform.prob``; MONAI's default range
        0.9 - 1.1): per sample one factor for SAX and one shared by the LAX views (two ``RandZoomd`` calls in the
        reference's Compose), drawn here and applied on the device by ``DeviceFeeder``."""
        self.ds, self.sampler, self.b = dataset, sampler, batch_size
        self.zoom_prob, self.zoom_range = float(zoom_prob), (float(zoom_range[0]), float(zoom_range[1]))
        self.sizes = {v: tuple(image_size_dict[v]) for v in dataset.views}
        self.n_frames = n_frames
        self.rng = np.random.default_rng(seed)  # cinema/mae/pretrain.py:134 (unseeded there)
        pin = torch.cuda.is_available() if pin_memory is None else pin_memory
        self._buffers = []
        for _ in range(n_buffers):
            images, lo, hi = {}, {}, {}
            for v in dataset.views:
                dt = torch.from_numpy(np.empty(0, dtype=np.dtype(dataset.subjects[0]["views"][v]["dtype"]))).dtype
                images[v] = torch.empty((batch_size, 1, *self.sizes[v]), dtype=dt)
                lo[v], hi[v] = torch.empty(batch_size), torch.empty(batch_size)
                if pin:
                    images[v], lo[v], hi[v] = images[v].pin_memory(), lo[v].pin_memory(), hi[v].pin_memory()
            extent = zoom = None
            if self.zoom_prob > 0.0:
                extent = {v: torch.ones((batch_size, 3), dtype=torch.int32) for v in dataset.views}
                zoom = {v: torch.ones(batch_size, dtype=torch.float32) for v in dataset.views}
                if pin:
                    extent = {v: t.pin_memory() for v, t in extent.items()}
                    zoom = {v: t.pin_memory() for v, t in zoom.items()}
            self._buffers.append(RawBatch(images, lo, hi, extent, zoom))
        self._next = 0

    def __len__(self) -> int:
        return len(self.sampler) // self.b

    def _fill(self, out: RawBatch, indices: list[int]) -> RawBatch:
        for j, idx in enumerate(indices):
            t = int(self.rng.integers(self.n_frames))  # ONE frame index per sample, shared by its views
            z3 = z2 = 1.0
            if self.zoom_prob > 0.0:  # RandZoomd(keys="sax") then RandZoomd(keys=lax views): one Bernoulli + one factor each
                if self.rng.random() < self.zoom_prob:
                    z3 = float(np.float32(self.rng.uniform(*self.zoom_range)))
                if self.rng.random() < self.zoom_prob:
                    z2 = float(np.float32(self.rng.uniform(*self.zoom_range)))
            for v in self.ds.views:
                frame, mn, mx = self.ds.frame(idx, v, t)
                size = self.sizes[v]
                if frame.ndim != len(size) or any(a > s for a, s in zip(frame.shape, size)):
                    raise ValueError(f"frame of view {v} has shape {frame.shape}, larger than the patch size {size}")
                dst = out.images[v][j, 0].numpy()
                if frame.shape != size:
                    dst[...] = np.asarray(mn).astype(dst.dtype)  # scales to 0: SpatialPad runs after ScaleIntensity
                dst[tuple(slice(0, a) for a in frame.shape)] = frame
                out.lo[v][j], out.hi[v][j] = mn, mx
                if out.zoom is not None:
                    out.zoom[v][j] = z3 if frame.ndim == 3 else z2
                    out.extent[v][j] = torch.tensor(list(frame.shape) + [1] * (3 - frame.ndim), dtype=torch.int32)
        return out

    def __iter__(self) -> Iterator[RawBatch]:
        idx = self.sampler.indices()
        for k in range(len(idx) // self.b):  # drop_last=True (cinema/mae/pretrain.py:335)
            buf = self._buffers[self._next]
            self._next = (self._next + 1) % len(self._buffers)
            yield self._fill(buf, idx[k * self.b:(k + 1) * self.b])


# ------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------
def scale_intensity(raw: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """MONAI ``ScaleIntensity(minv=0.0, maxv=1.0)`` per sample: (x - min) / (max - min) in fp32; a constant image
    (max == min) maps to 0 (``rescale_array`` returns ``arr * minv``).  ``raw`` (B, ...) any dtype, ``lo`` / ``hi`` (B,)."""
    shape = (raw.shape[0],) + (1,) * (raw.dim() - 1)
    lo32 = lo.to(device=raw.device, dtype=torch.float32).view(shape)
    span = hi.to(device=raw.device, dtype=torch.float32).view(shape) - lo32
    inv = torch.where(span > 0, 1.0 / span.clamp_min(torch.finfo(torch.float32).tiny), torch.zeros_like(span))
    res = torch.sub(raw.to(torch.float32), lo32, out=out)
    return res.mul_(inv)


def zoom_intensity(raw: torch.Tensor, extent: torch.Tensor, zoom: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """RandZoom(keep_size) -> ScaleIntensity -> SpatialPad("end") restated in torch: the CPU feeder's path and the reference
    the CUDA kernel (``cb_zoom_intensity``) is tested against.  MONAI ``Zoom`` (monai/transforms/spatial/functional.py
    ``zoom``): ``interpolate(frame, scale_factor=z, mode, align_corners=False)`` (trilinear for 3-D frames, bicubic for
    2-D), then per axis pad ``[half, diff - half]`` zeros or slice ``[half, half + size)`` with ``half = |diff| // 2`` back
    to the frame's own size; ScaleIntensity over the zoomed frame; zeros outside the frame.
    raw (B, 1, *size), extent (B, 3) int, zoom (B,) fp32."""
    import torch.nn.functional as F

    nd = raw.dim() - 2
    res = torch.zeros(raw.shape, dtype=torch.float32, device=raw.device) if out is None else out.zero_()
    for j in range(raw.shape[0]):
        e = [int(x) for x in extent[j, :nd]]
        z = float(zoom[j])
        box = tuple(slice(0, a) for a in e)
        img = raw[j, 0][box].to(torch.float32)
        if z != 1.0:
            zoomed = F.interpolate(img[None, None], scale_factor=[z] * nd, mode="trilinear" if nd == 3 else "bicubic",
                                   align_corners=False)[0, 0]
            pads, cut = [], []
            for od, zd in zip(e, zoomed.shape):
                diff, half = od - zd, abs(od - zd) // 2
                pads.append((half, diff - half) if diff > 0 else (0, 0))
                cut.append(slice(half, half + od) if diff < 0 else slice(None))
            zoomed = F.pad(zoomed, [p for pr in reversed(pads) for p in pr])[tuple(cut)]
        else:
            zoomed = img
        lo, hi = zoomed.min(), zoomed.max()
        res[j, 0][box] = (zoomed - lo) / (hi - lo) if float(hi) > float(lo) else torch.zeros_like(zoomed)
    return res


class DeviceFeeder:
    """Raw batches -> fp32 device batches for ``MAETrainer.step``: every ``RawBatch`` is uploaded on a copy stream into
    one of two device staging sets (overlapping the step in flight), then scaled on the compute stream.  Iterating yields
    ``{view: (B, 1, *size) fp32 device tensor}``; the tensors of a yielded batch stay valid until the batch after next."""

    def __init__(self, batcher: FrameBatcher, device: torch.device | str) -> None:
        self.batcher = batcher
        self.dev = torch.device(device)
        self.cuda = self.dev.type == "cuda"
        self._copy = torch.cuda.Stream(device=self.dev) if self.cuda else None
        self._slots: list[dict] = []
        self.h2d_bytes_per_batch = 0

    def _slot(self, k: int, raw: RawBatch) -> dict:
        while len(self._slots) <= k:
            self._slots.append({
                "img": {v: torch.empty(t.shape, dtype=t.dtype, device=self.dev) for v, t in raw.images.items()},
                "lo": {v: torch.empty(t.shape, device=self.dev) for v, t in raw.lo.items()},
                "hi": {v: torch.empty(t.shape, device=self.dev) for v, t in raw.hi.items()},
                "out": {v: torch.empty(t.shape, dtype=torch.float32, device=self.dev) for v, t in raw.images.items()},
                "free": torch.cuda.Event() if self.cuda else None,
            })
            if raw.zoom is not None:
                self._slots[-1]["extent"] = {v: torch.empty(t.shape, dtype=t.dtype, device=self.dev) for v, t in raw.extent.items()}
                self._slots[-1]["zoom"] = {v: torch.empty(t.shape, dtype=t.dtype, device=self.dev) for v, t in raw.zoom.items()}
                self._slots[-1]["keys"] = {v: torch.empty(2 * t.shape[0], dtype=torch.int32, device=self.dev) for v, t in raw.zoom.items()}
        return self._slots[k]

    @staticmethod
    def _groups(raw: RawBatch):
        g = [("img", raw.images), ("lo", raw.lo), ("hi", raw.hi)]
        if raw.zoom is not None:
            g += [("extent", raw.extent), ("zoom", raw.zoom)]
        return g

    def _upload(self, k: int, raw: RawBatch):
        slot = self._slot(k, raw)
        self.h2d_bytes_per_batch = raw.nbytes()
        slot["zoomed"] = raw.zoom is not None
        if not self.cuda:
            for name, src in self._groups(raw):
                for v, t in src.items():
                    slot[name][v].copy_(t)
            return slot, None
        ready = torch.cuda.Event()
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(slot["free"])  # the scale kernels that last read this slot have run
            for name, src in self._groups(raw):
                for v, t in src.items():
                    slot[name][v].copy_(t, non_blocking=True)
            ready.record(self._copy)
        return slot, ready

    def _finish(self, slot: dict, ready) -> dict[str, torch.Tensor]:
        if ready is not None:
            torch.cuda.current_stream(self.dev).wait_event(ready)
        if self.cuda:  # one fused raw-dtype -> fp32 ScaleIntensity pass per view (csrc/elementwise.cu)
            from cinema_b200 import _C

            out = {}
            for v, img in slot["img"].items():
                if slot["zoomed"]:  # RandZoom + ScaleIntensity + end padding in two passes (cb_zoom_intensity); no torch path
                    out[v] = _C.zoom_intensity(img, slot["extent"][v], slot["zoom"][v], slot["out"][v], slot["keys"][v])
                elif img.dtype in _C._RAW_DT and (img[0].numel() * img.element_size()) % 16 == 0:
                    out[v] = _C.scale_intensity(img, slot["lo"][v], slot["hi"][v], slot["out"][v])
                else:  # odd sample sizes / dtypes keep the two-kernel torch form
                    out[v] = scale_intensity(img, slot["lo"][v], slot["hi"][v], out=slot["out"][v])
        elif slot["zoomed"]:
            out = {v: zoom_intensity(slot["img"][v], slot["extent"][v], slot["zoom"][v], out=slot["out"][v]) for v in slot["img"]}
        else:
            out = {v: scale_intensity(slot["img"][v], slot["lo"][v], slot["hi"][v], out=slot["out"][v]) for v in slot["img"]}
        if self.cuda:
            slot["free"].record(torch.cuda.current_stream(self.dev))
        return out

    def __iter__(self) -> Iterator[dict[str, torch.Tensor]]:
        it = iter(self.batcher)
        pending = None
        k = 0
        for raw in it:
            cur = self._upload(k % 2, raw)
            if self.cuda:
                self._copy.synchronize()  # the batcher may refill this pinned buffer two batches later: the copy must have run
            k += 1
            if pending is not None:
                yield self._finish(*pending)
            pending = cur
        if pending is not None:
            yield self._finish(*pending)
