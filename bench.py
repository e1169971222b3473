"""Benchmark of the MAE pre-training hot path (BASELINE.json: "MAE pretrain volumes/sec (ViT-B, mask 0.75)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--lax 192|256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one optimisation step over one synthetic batch of B frame-sets per GPU (one frame-set =
one "volume": 1 SAX 192x192x16 + 3 LAX 192x192, BASELINE.json configs[2] at one time frame):
host->device copy, mask draw, conv stem, ViT encoder on the 25 % visible tokens, cross-attention decoder,
masked-pixel MSE, full backward, gradient all-reduce (N > 1), global-norm clip and AdamW.

One JSON line is printed by rank 0:
  value      frame-set volumes/s, whole job, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e        the same through the public API with pinned HOST buffers: H2D copy of every batch and a D2H read of
             every step's loss inside the timed region
  roofline   the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time, measured live in an
             instrumented step, against the measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (unmodified reference modules from baseline/_ref, fp32, torch threads = host
             cores; the oracle port when that install did not travel) on a bounded sample
  stock_gpu_baseline  the unmodified reference modules on the SAME GPU in the same process (bf16 autocast + SDPA +
             GradScaler / clip / AdamW, the reference's loop; grad_ckpt on and off; DDP when N > 1): the number to beat
``--impl reference`` times the CPU path alone (the reference is pure Python + torch).
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

VIEWS = ("sax", "lax_2c", "lax_3c", "lax_4c")


# ------------------------------------------------------------------------------------------
# workload description and algorithmic FLOPs (SURVEY.md section 8d / BASELINE.md section 3)
# ------------------------------------------------------------------------------------------
def model_kwargs(size: str, sax, lax) -> dict:
    dims = {"base": (768, 12, 12), "large": (1024, 24, 16)}[size]
    return dict(
        image_size_dict={"sax": tuple(sax), **{v: tuple(lax) for v in VIEWS[1:]}},
        in_chans_dict={v: 1 for v in VIEWS},
        enc_patch_size_dict={"sax": (4, 4, 1), **{v: (4, 4) for v in VIEWS[1:]}},
        enc_scale_factor_dict={"sax": (2, 2, 1), **{v: (2, 2) for v in VIEWS[1:]}},
        enc_conv_chans=[64, 128], enc_conv_n_blocks=2,
        enc_embed_dim=dims[0], enc_depth=dims[1], enc_n_heads=dims[2], dec_embed_dim=512, dec_depth=8, dec_n_heads=16,
    )


def train_gflop_per_sample(kw: dict, ratio: float = 0.75) -> dict:
    """Forward multiply-add = 2 FLOPs; train = 3 x forward; recompute not credited."""
    d, le, dd, ld = kw["enc_embed_dim"], kw["enc_depth"], kw["dec_embed_dim"], kw["dec_depth"]
    n_tok, n_keep, conv = [], [], 0.0
    for v, size in kw["image_size_dict"].items():
        nd = len(size)
        eff = [a * b * b for a, b in zip(kw["enc_patch_size_dict"][v], kw["enc_scale_factor_dict"][v])]
        grid = [s // e for s, e in zip(size, eff)]
        n = math.prod(grid)
        n_tok.append(n)
        n_keep.append(int(n * (1 - ratio)))
        pos, cin, ps = list(size), 1, kw["enc_patch_size_dict"][v]
        for lvl, ch in enumerate(kw["enc_conv_chans"]):
            pos = [s // p for s, p in zip(pos, ps)]
            p = math.prod(pos)
            conv += 2 * p * ch * cin * math.prod(ps)                      # strided patch conv
            conv += kw["enc_conv_n_blocks"] * (2 * p * ch * ch * 10 + 2 * p * ch * 5 ** nd)  # 1x1 x2, mlp 4x x2, dw 5^nd
            k = [s // g for s, g in zip(pos, grid)]
            conv += 2 * n * d * ch * math.prod(k)                         # fusion conv over all tokens
            cin, ps = ch, kw["enc_scale_factor_dict"][v]
    n_enc = 1 + sum(n_keep)
    n_k = sum(n_keep)
    n_q = 1 + sum(n_tok) - n_k
    enc = le * (24 * n_enc * d * d + 4 * n_enc * n_enc * d)
    dec = ld * (2 * dd * dd * (10 * n_q + 2 * n_k) + 4 * n_q * n_k * dd)
    other = sum(2 * n * (512 * d + d * d) for n in n_tok) + 2 * n_enc * d * dd + 2 * (n_q - 1) * dd * 256
    fwd = enc + dec + other + conv
    return dict(fwd_gflop=fwd / 1e9, train_gflop=3 * fwd / 1e9, enc_tokens=n_enc, q_tokens=n_q, kv_tokens=n_k,
                attn_fwd_gflop=(le * 4 * n_enc * n_enc * d + ld * 4 * n_q * n_k * dd) / 1e9)


def measured_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def measured_write_gbs(dev) -> float:
    """HBM WRITE bandwidth, measured live: a 1 GiB fill (no reads), best of 6.  The copy figure of MEASURED_PEAKS.json counts
    read + write bytes; this pool's B200s fill at ~7.1 TB/s, so the written-bytes roof below has not bound any GEMM shape so
    far (profiles/r02_gemm_epilogue.md records the hypothesis it was added to test)."""
    buf = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    best = 0.0
    for i in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        buf.fill_(float(i))
        e.record()
        torch.cuda.synchronize()
        if i:
            best = max(best, buf.numel() * 4 / (s.elapsed_time(e) * 1e-3) / 1e9)
    del buf
    return round(best, 1)


def synthetic_batch(kw: dict, b: int, seed: int, pin: bool) -> dict:
    """Images in [0, 1) like ScaleIntensityd (cinema/mae/pretrain.py:184)."""
    g = torch.Generator().manual_seed(seed)
    out = {v: torch.rand(b, 1, *s, generator=g) for v, s in kw["image_size_dict"].items()}
    return {k: t.pin_memory() for k, t in out.items()} if pin else out


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference path) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_throughput(kw: dict, sample_b: int, steps: int, warmup: int) -> dict:
    """The reference's own CPU path: the unmodified reference modules from ``baseline/_ref`` when the install travelled
    with the repo (``kind: reference``), else the oracle port (``kind: port``)."""
    from baseline import stock

    if stock.available():
        return stock.cpu_step_time(kw, synthetic_batch(kw, sample_b, seed=0, pin=False), steps, warmup)
    from oracle import cinema_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = oracle.MAEConfig(**kw)
    sd = oracle.init_state_dict(cfg, seed=0)
    batch = synthetic_batch(kw, sample_b, seed=0, pin=False)
    torch.manual_seed(0)
    masks = {v: oracle.random_patch_mask(sample_b, cfg.n_patches(v), 0.75) for v in kw["image_size_dict"]}
    for _ in range(warmup):
        oracle.train_step_cpu(sd, cfg, batch, masks)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.train_step_cpu(sd, cfg, batch, masks)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": sample_b / dt, "unit": "frame-set volumes/s", "cores": cores, "kind": "port",
            "sample": f"{steps} forward+backward steps of {sample_b} frame-sets (same views / sizes / ViT-B, fp32, "
                      f"torch {torch.__version__} CPU, {cores} threads), {dt:.2f} s per step", "ms_per_step": dt * 1e3}


# ------------------------------------------------------------------------------------------
# per-kernel instrumentation (one eager step, CUDA events around every C-ABI launch)
# ------------------------------------------------------------------------------------------
def instrumented_step(trainer, batch_dev: dict) -> dict:
    from cinema_b200 import _C

    records: list[tuple[str, float, torch.cuda.Event, torch.cuda.Event]] = []
    shapes: list = []
    calls: list = []  # every C-ABI launch of the step (arguments kept alive) for the queued replay below
    originals = {}

    def flops_of(name, args, kwargs) -> float:
        if name == "gemm":
            a, b = args[0], args[1]
            m, k = (a.shape[1], a.shape[0]) if kwargs.get("a_mn") else (a.shape[0], a.shape[1])
            n = b.shape[1] if kwargs.get("b_mn") else b.shape[0]
            return 2.0 * m * n * k
        if name in ("attention_fwd", "attention_bwd"):
            q, k = args[0], args[1]
            bb, nq, h, d = q.shape
            return (4.0 if name == "attention_fwd" else 8.0) * bb * h * nq * k.shape[1] * d
        return 0.0

    def wrap(name):
        fn = getattr(_C, name)
        originals[name] = fn

        def timed(*args, **kwargs):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*args, **kwargs)
            e.record()
            tag = name
            if name == "gemm":
                a_, b_ = args[0], args[1]
                m_, k_ = (a_.shape[1], a_.shape[0]) if kwargs.get("a_mn") else (a_.shape[0], a_.shape[1])
                n_ = b_.shape[1] if kwargs.get("b_mn") else b_.shape[0]
                o_ = args[2]
                shapes.append((f"M{m_} N{n_} K{k_} a_mn={int(bool(kwargs.get('a_mn')))} b_mn={int(bool(kwargs.get('b_mn')))} "
                               f"epi={kwargs.get('epilogue', 0)} res={int(kwargs.get('residual') is not None)} "
                               f"o32={int(o_ is not None and o_.dtype == torch.float32)} acc={int(bool(kwargs.get('accumulate')))}", s, e,
                               2.0 * m_ * n_ * k_))
            records.append((tag, flops_of(name, args, kwargs), s, e))
            calls.append((name, fn, args, kwargs, flops_of(name, args, kwargs), shapes[-1][0] if name == "gemm" else name))
            return r

        setattr(_C, name, timed)

    from cinema_b200 import engine

    names = ["gemm", "colsum", "attention_fwd", "attention_bwd", "layernorm_fwd", "layernorm_bwd", "cast_bf16",
             "mask_to_index", "gather_rows", "scatter_rows", "embed_rows", "colsum_seg", "scale_cast", "mae_loss_finalize",
             "gather_patches", "scatter_patches", "masked_mse_fwd", "sumsq", "adamw_flat", "dwconv_tokens",
             "dwconv_tokens_wgrad", "expand_token_index"]
    for n in names:
        wrap(n)
    engine.VIEW_STREAMS_ENABLED = False  # one stream: every kernel's event pair brackets that kernel alone
    pdl_was = _C.set_pdl(False)          # and no prologue overlap with the predecessor
    try:
        s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        # park the GPU so that the host (~0.2 ms of Python / ctypes per launch) runs ahead: the launches then execute
        # back to back and each event pair measures device time only
        torch.cuda._sleep(int(6e8))
        # ... and bring the clocks back up behind the sleep (an idle-then-burst start measured the first ~10 ms of
        # kernels 30 % slow), still ahead of the step's launches in the queue
        wa = torch.randn(8192, 8192, device=next(iter(batch_dev.values())).device, dtype=torch.bfloat16)
        for _ in range(24):
            torch.matmul(wa, wa)
        s_all.record()
        trainer.eager_step(batch_dev)
        e_all.record()
        torch.cuda.synchronize()
    finally:
        engine.VIEW_STREAMS_ENABLED = True
        _C.set_pdl(pdl_was)
        for n, fn in originals.items():
            setattr(_C, n, fn)
    # Replay of ALL the step's launches in their original order (same arguments, same buffers), queued behind a GPU-side
    # sleep and a short clock warm-up: the eager step above is host-bound (two event records and a Python wrapper per
    # launch), which distorts the device time of whatever runs while the queue is draining -- round 1 reported the
    # HBM-bound kernels (LayerNorm, gathers) from the eager events, 1.4 - 1.7 x their ncu durations.  The replay runs
    # after the timed region and the loss read-back; its in-place side effects (gradient accumulation, a second AdamW
    # update) touch nothing that is reported.
    torch.cuda.synchronize()
    torch.cuda._sleep(int(3e8))
    replay = []
    for name, fn, args, kwargs, fl, tag in calls:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(*args, **kwargs)
        e.record()
        replay.append((name, tag, fl, s, e))
    torch.cuda.synchronize()
    rep: dict[str, list[float]] = {}
    by_shape = {}
    for name, tag, fl, s, e in replay:
        a = rep.setdefault(name, [0.0, 0.0, 0])
        t = s.elapsed_time(e)
        a[0] += t
        a[1] += fl
        a[2] += 1
        if name == "gemm":
            b = by_shape.setdefault(tag, [0.0, 0.0, 0])
            b[0] += t
            b[1] += fl
            b[2] += 1
    calls.clear()
    agg: dict[str, list[float]] = {}
    for name, fl, s, e in records:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += s.elapsed_time(e)
        a[1] += fl
        a[2] += 1
    total = s_all.elapsed_time(e_all)
    own = sum(a[0] for a in agg.values())
    eager_own = own
    for name, v in rep.items():  # take the replayed (queue-fed) device times
        agg[name] = v
    own = sum(a[0] for a in agg.values())
    peaks, _ = measured_peaks()
    write_gbs = measured_write_gbs(next(iter(batch_dev.values())).device)

    def shape_roof(tag: str, ms: float, launches: int, flops: float) -> dict:
        """Which roof bounds this GEMM shape?  Algorithmic bytes: A + B (bf16) + the stored output (+ the fp32 residual it
        reads, + the bf16 GELU' operand, + the second bf16 output of the GELU flavour); split-K accumulation reads and
        writes the fp32 output.  Three candidate times: tensor (sustained bf16 peak), all bytes at the copy bandwidth, and
        the WRITTEN bytes at the write bandwidth measured above (a fill)."""
        f = dict(kv.split("=") for kv in tag.split()[3:])
        m, n, k = (int(x[1:]) for x in tag.split()[:3])
        o32, res, epi, acc = int(f["o32"]), int(f["res"]), int(f["epi"]), int(f["acc"])
        by = 2.0 * (m * k + n * k) + m * n * (4 if o32 else 2) * (2 if acc else 1) + res * 4.0 * m * n
        by += 2.0 * m * n if epi in (1, 2) else 0.0
        wr = m * n * (4.0 if o32 else 2.0) + (2.0 * m * n if epi == 1 else 0.0)
        t_hbm = by / (peaks["hbm_gbs"] * 1e9)
        t_wr = wr / (write_gbs * 1e9) if write_gbs > 0 else 0.0
        t_tc = (flops / launches) / (peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) * 1e12)
        t = ms / launches / 1e3
        bound, t_b = max((("tensor", t_tc), ("hbm", t_hbm), ("hbm-write", t_wr)), key=lambda kv: kv[1])
        return {"bound": bound, "frac_of_bound": round(t_b / t, 3), "gb_per_s": round(by / t / 1e9, 0),
                "written_gb_per_s": round(wr / t / 1e9, 0)}

    top_shapes = [{"gemm": k, "launches": v[2], "ms": round(v[0], 3), "tflops": round(v[1] / v[0] / 1e9, 1),
                   **shape_roof(k, v[0], v[2], v[1])}
                  for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:24]]
    return {"step_ms": total, "own_kernels_ms": own, "own_kernels_eager_events_ms": eager_own, "hbm_write_gbs": write_gbs,
            "note": "one eager step on a single stream records every C-ABI launch; the per-kernel times are CUDA events around "
                    "each launch of a REPLAY of those calls in order, queued behind a GPU-side sleep (device time only, no "
                    "host gaps); step_ms is the host-bound eager step and is not a throughput figure",
            "gemm_top_shapes": top_shapes,
            "kernels": {k: {"ms": round(v[0], 3), "launches": v[2], "tflops": round(v[1] / v[0] / 1e9, 1) if v[0] > 0 and v[1] else None}
                        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}}


# ------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="frame-sets per GPU per step (reference: batch_size_per_device 16)")
    ap.add_argument("--size", default="base", choices=["base", "large"])
    ap.add_argument("--lax", type=int, default=192, help="LAX edge: 192 (BASELINE.json wording) or 256 (reference default)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the timed host-buffer pass")
    ap.add_argument("--no-stock-gpu", action="store_true", help="skip the stock torch arm (unmodified reference on the same GPU)")
    ap.add_argument("--stock-steps", type=int, default=5)
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kw = model_kwargs(args.size, (192, 192, 16), (args.lax, args.lax))
    work = train_gflop_per_sample(kw)
    workload = (f"4-view cine frame-set: SAX 192x192x16 + 3 LAX {args.lax}x{args.lax}, ViT-{args.size[0].upper()} MAE, mask 0.75, "
                f"enc/q/kv tokens {work['enc_tokens']}/{work['q_tokens']}/{work['kv_tokens']}")
    config = {"workload": workload, "batch_per_gpu": args.batch, "global_batch": args.batch * world, "parallelism": f"dp{world}",
              "optimizer": "AdamW lr 1e-3 betas (0.9, 0.95) wd 0.05, clip 5.0", "train_gflop_per_volume": round(work["train_gflop"], 1),
              "cache": "per-step working set (activations > 5 GB) far exceeds the 126 MB L2; no flush needed",
              "cine32_volumes": "divide value by 32 for 32-frame cine volumes/s"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample_b = 2 if args.steps + args.warmup <= 30 else 1
        cpu = cpu_reference_throughput(kw, sample_b, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "mae_pretrain_volumes_per_sec", "value": cpu["value"], "unit": "frame-set volumes/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {**config, "batch_per_gpu": sample_b, "global_batch": sample_b, "parallelism": "cpu"},
                "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cpu["value"], "unit": "frame-set volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    import torch.distributed as dist

    from cinema_b200 import CineMA, _C
    from cinema_b200.train import MAETrainer

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _C.lib()
    torch.backends.cudnn.benchmark = True  # cinema/device.py:63
    torch.manual_seed(0 + rank)            # cinema/mae/pretrain.py:309-310
    model = CineMA(**kw).to(dev)
    model.train()
    trainer = MAETrainer(model, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05, clip_grad=5.0, enc_mask_ratio=0.75,
                         use_cuda_graph=not args.no_graph)
    host = synthetic_batch(kw, args.batch, seed=rank, pin=True)
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps: int) -> float:
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident() -> None:
        trainer.step(resident)

    def step_e2e() -> None:
        # one H2D upload of a pinned host batch and one D2H read of the loss per step, both inside the timed region; the
        # upload of the NEXT step's batch is issued right after this step's launch so that it overlaps the compute
        # (trainer.prefetch: copy stream + staging buffer), as a pinned-memory data loader would
        loss = trainer.step(host)
        trainer.prefetch(host)
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host reads this step's loss

    for _ in range(warmup):
        step_e2e()
    with ClockSampler(local_rank) as clocks:
        ms = timed(step_resident, args.steps)
        if not args.skip_e2e:
            trainer.prefetch(host)  # the first timed step's upload; every timed step issues exactly one more
        ms_e2e = timed(step_e2e, args.steps) if not args.skip_e2e else float("nan")
    final_loss = float(loss_host[0])
    n_vol = args.batch * world * args.steps
    value = n_vol / (ms / 1e3)
    e2e_value = n_vol / (ms_e2e / 1e3)

    peaks, peak_kind = measured_peaks()
    prof = None
    if not args.no_kernel_profile:
        prof = instrumented_step(trainer, resident)
    line = {
        "metric": "mae_pretrain_volumes_per_sec", "value": round(value, 2), "unit": "frame-set volumes/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
        "e2e": {"value": round(e2e_value, 2), "unit": "frame-set volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": trainer.launches_per_step * args.steps, "gpu_launches_per_step": trainer.launches_per_step,
        "cuda_graph": trainer.use_graph, "final_loss": final_loss, "clocks": clocks.summary(),
    }
    sustained = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    step_tflops = value / world * work["train_gflop"] / 1e3
    line["step_roofline"] = {"bound": "tensor", "achieved": round(step_tflops, 1), "peak": sustained, "unit": "TFLOP/s",
                             "frac": round(step_tflops / sustained, 4), "note": f"whole step per GPU, algorithmic train FLOPs, {peak_kind} sustained peak"}
    if prof is not None:
        g = prof["kernels"].get("gemm")
        if g and g["tflops"]:
            traffic = None
            tp = ROOT / "profiles" / "gemm_traffic.json"  # written by tools/summarize_launches.py from an ncu capture
            if tp.exists():
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            line["roofline"] = {"bound": "tensor", "achieved": g["tflops"], "peak": sustained, "unit": "TFLOP/s",
                                "frac": round(g["tflops"] / sustained, 4), "traffic": traffic,
                                "kernel": "gemm_bf16_kernel (tcgen05 GEMM, all launches of one step: algorithmic 2MNK FLOPs "
                                          "of a launch / its CUDA-event duration, averaged over the step's launches)",
                                "launches_per_step": g["launches"], "avg_launch_us": round(1e3 * g["ms"] / g["launches"], 2),
                                "avg_gflop_per_launch": round(g["tflops"] * g["ms"] / g["launches"], 2),
                                "share_of_step": round(g["ms"] / (ms / args.steps), 3), "peak_kind": f"{peak_kind} sustained"}
            a_f, a_b = prof["kernels"].get("attention_fwd"), prof["kernels"].get("attention_bwd")
            if a_f and a_b:
                line["attention_roofline"] = {"bound": "tensor", "fwd_tflops": a_f["tflops"], "bwd_tflops": a_b["tflops"],
                                              "peak": sustained, "unit": "TFLOP/s",
                                              "frac": round((a_f["tflops"] * a_f["ms"] + a_b["tflops"] * a_b["ms"]) /
                                                            (a_f["ms"] + a_b["ms"]) / sustained, 4),
                                              "share_of_step": round((a_f["ms"] + a_b["ms"]) / (ms / args.steps), 3),
                                              "note": "tcgen05 flash attention (head_dim 64 encoder / 32 decoder); "
                                                      "4 BHNqNkd fwd, 8 BHNqNkd bwd credited"}
        line["kernel_profile"] = prof
    if not args.no_stock_gpu:
        # the number to beat (north_star): the UNMODIFIED reference modules on this same GPU, same batch, same process,
        # after our own timing is finished; all ranks take part when N > 1 (DistributedDataParallel, cinema/device.py:86-104)
        from baseline import stock

        if stock.available():
            del trainer, model
            torch.cuda.empty_cache()
            arms = []
            for ckpt in (True, False):  # reference default (cinema/mae/config.yaml:3) first, then the faster setting
                r = stock.stock_gpu_step_time(kw, host, dev, steps=args.stock_steps, warmup=3, grad_ckpt=ckpt,
                                              ddp=world > 1, local_rank=local_rank)
                t = torch.tensor([r["ms_per_step"]], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                r["ms_per_step"] = round(float(t), 3)
                r["value"] = round(args.batch * world / (float(t) / 1e3), 2)
                arms.append(r)
            best = max(arms, key=lambda r: r["value"])
            line["stock_gpu_baseline"] = {**best, "n_gpus": world, "arms": [
                {k: a[k] for k in ("grad_ckpt", "value", "ms_per_step", "peak_mem_gib")} for a in arms],
                "speedup_e2e": round(e2e_value / best["value"], 2) if e2e_value == e2e_value else None,
                "speedup_vs_reference_default_grad_ckpt": round(e2e_value / arms[0]["value"], 2) if e2e_value == e2e_value else None}
        else:
            line["stock_gpu_baseline"] = {"unavailable": "baseline/_ref not installed (python baseline/install_reference.py)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_throughput(kw, 2, 2, 1)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
