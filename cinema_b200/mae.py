"""CineMA masked autoencoder with the reference's ``nn.Module`` API (cinema/mae/mae.py) on the B200 path.

``CineMA(...)`` takes the reference's constructor arguments, exposes the same attributes and the same
``state_dict`` keys, and ``forward(image_dict, enc_mask_ratio)`` returns the same
``(loss, pred_dict, enc_mask_dict, metrics)``.  Underneath, one ``torch.autograd.Function`` runs the
whole token path -- visible-patch embedding, ViT encoder, multi-scale fusion, decoder embedding,
cross-/self-attention decoder, prediction heads, masked-pixel MSE -- as an explicit sequence of
sm_100a kernel launches with a hand-written backward, and accumulates parameter gradients in the
flat fp32 gradient arena (cinema_b200/arena.py).

What is evaluated differently from the reference, with identical results:
  * patch embedding (``patch_embed.proj``, ``linear``) and the fusion convs run on the 25 % visible tokens
    only -- the reference computes all tokens and gathers afterwards (cinema/mae/mae.py:548-550,
    cinema/convvit.py:284-288); both are row-wise maps, so gather-then-compute is exact;
  * boolean-mask gathers become index gathers from one mask -> index kernel (no ``nonzero`` host syncs);
  * the kv projections of all cross-attention decoder blocks read the same un-normalised visible
    tokens (cinema/vit.py:589,773), so they are one GEMM with N = depth * 2 * D;
  * the per-view ``isfinite`` test on the loss (cinema/mae/mae.py:604) is evaluated on the device.
"""

from __future__ import annotations

import math

import torch
from torch import nn

from cinema_b200 import _C, engine, stem
from cinema_b200.arena import ensure_arena
from cinema_b200.conv import Linear
from cinema_b200.convvit import DownsampleEncoder, MultiScaleFusion
from cinema_b200.vit import Mlp, ViTDecoder, ViTEncoder, get_pos_embed, get_tokens, get_vit_config, init_weights

F32, BF16 = torch.float32, torch.bfloat16


# ------------------------------------------------------------------------------------------
# mask, decoder embedding, loss (public helpers of the reference module)
# ------------------------------------------------------------------------------------------
def get_batch_random_patch_mask(batch_size: int, n_patches: int, mask_ratio: float, device: torch.device) -> torch.Tensor:
    """(B, n) bool, 1 = removed, exactly int(n * (1 - ratio)) zeros per row (cinema/mae/mae.py:30-65).
    Kept on torch's RNG and sort so that a seed gives the same masks as the reference."""
    if mask_ratio < 0:
        raise ValueError(f"mask_ratio must be positive, got {mask_ratio}.")
    if mask_ratio == 0:
        return torch.zeros((batch_size, n_patches), dtype=torch.bool, device=device)
    n_keep = int(n_patches * (1 - mask_ratio))
    noise = torch.rand(batch_size, n_patches, device=device)
    ids_shuffle = torch.argsort(noise, dim=1)
    # the reference sorts a second time (ids_restore = argsort(ids_shuffle)) and gathers a [0..0 1..1] row through it;
    # the argsort of a permutation is its inverse, so that is exactly "clear the n_keep first entries of ids_shuffle":
    # one scatter instead of a second radix sort + gather, bit-identical masks
    mask = torch.ones((batch_size, n_patches), dtype=torch.bool, device=device)
    return mask.scatter_(1, ids_shuffle[:, :n_keep], False)


def get_decoder_patch_size(image_size, n_conv_layers, enc_patch_size, enc_scale_factor) -> tuple[int, ...]:
    """Product of the stem strides = pixels per ViT token (cinema/mae/mae.py:207-228)."""
    out = tuple(enc_patch_size)
    for _ in range(n_conv_layers):
        out = tuple(a * b for a, b in zip(out, enc_scale_factor))
    if len(out) != len(image_size):
        raise ValueError(f"patch size {out} does not match image size {image_size}")
    return out


class _MaskedMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, pred, enc_mask, norm_target, eps):
        b, n, e = target.shape
        n_drop = pred.shape[1]
        _, _, slot = _C.mask_to_index(enc_mask.contiguous(), n - n_drop)
        acc = torch.zeros(8, dtype=F32, device=pred.device)
        acc[3:5] = float("-inf")
        p32 = pred.detach().to(F32).contiguous()
        diff = torch.empty_like(p32)
        _C.masked_mse_fwd(target.detach().to(F32).contiguous().view(b, 1, n, e), (1, e), enc_mask.contiguous(), slot, p32,
                          bool(norm_target), eps, acc, diff)
        ctx.save_for_backward(diff)
        ctx.count = p32.numel()
        ctx.dt = pred.dtype
        return acc

    @staticmethod
    def backward(ctx, g):
        (diff,) = ctx.saved_tensors
        return None, (diff * (2.0 / ctx.count) * g[0]).to(ctx.dt), None, None, None


def mse_loss(target: torch.Tensor, pred: torch.Tensor, enc_mask: torch.Tensor, norm_target: bool, eps: float = 1e-6):
    """Masked-patch MSE and target statistics (cinema/mae/mae.py:107-152) through the fused loss kernel.
    target (B, n_patches, E), pred (B, n_masked, E), enc_mask (B, n_patches) bool with 1 = masked."""
    b, n, _ = target.shape
    acc = _MaskedMSEFn.apply(target, pred, enc_mask, norm_target, eps)
    loss = acc[0] / max(pred.numel(), 1) if pred.numel() > 0 else acc[0] * float("nan")
    metrics = {"target_mean": acc[1].detach() / (b * n), "target_std": acc[2].detach() / (b * n), "mse_loss": loss}
    if norm_target and pred.shape[1] > 0:
        metrics["normed_target_max"] = acc[3].detach()
        metrics["pred_max"] = acc[4].detach()
    return loss, metrics


class DecoderEmbedding(nn.Module):
    """Decoder positional embedding and mask token of one view (cinema/mae/mae.py:155-204).  Parameters only:
    the add / gather is one kernel inside the fused path."""

    def __init__(self, enc_grid_size: tuple[int, ...], dec_embed_dim: int, add_embed_token: bool) -> None:
        super().__init__()
        self.pos_embed = get_pos_embed(embed_dim=dec_embed_dim, grid_size=enc_grid_size)
        self.embed_token = get_tokens(embed_dim=dec_embed_dim, n_tokens=1) if add_embed_token else None
        self.mask_token = get_tokens(embed_dim=dec_embed_dim, n_tokens=1)


# ------------------------------------------------------------------------------------------
# the fused token path
# ------------------------------------------------------------------------------------------
def _rows(t: torch.Tensor, start: int, count: int) -> torch.Tensor:
    return t[start:start + count]


def model_grid(model, view, source_levels):
    """ViT token grid of this call for ``view`` (stored on the source list by the callers)."""
    return source_levels.token_grid


class _Sources(list):
    """Per-view list of (src, grid, idx) gather sources plus the ViT token grid they describe."""

    token_grid: tuple[int, ...] = ()


class _Encoded:
    """Saved state of the encoder half (embedding -> ViT encoder -> fusion), shared by MAE and feature paths."""


def _lin_of(arena, mod, train):
    return engine.linw(arena, mod.weight, mod.bias, train)


def _encode(model, arena, views, sources, keep, n_keeps, b, train, want_fused32, fuse=True):
    """Visible-token embedding, ViT encoder and multi-scale fusion (``fuse=False``: stop after the encoder's final
    LayerNorm and return (None, {"enc": (B, 1 + sum n_keep, D) fp32}, state) -- models without a fusion stage).

    sources[v]: per stem level a triple (src, grid, idx) such that ``gather_patches(src, grid, patch, idx, ...)`` yields
    one row per visible token -- either a dense (B, C, *spatial) map with the token grid and ``keep`` ids, or the
    token-major output of the native stem viewed as (T, C, *f) with a unit grid; the last entry feeds the patch
    embedding (for a model without stem it is the image).  Returns (F16, fused32 | None, state)."""
    dev = keep[0].device
    d = model.encoder.cls_token.shape[-1]
    n = 1 + sum(n_keeps)
    st = _Encoded()
    st.n, st.d = n, d
    x0 = torch.empty((b, n, d), dtype=F32, device=dev)
    _C.embed_rows(None, 0, model.encoder.cls_token.data.view(-1), None, None, b, 1, out=x0, out_off=0)
    st.embed = []
    off = 1
    offs = []
    vs = engine.ViewStreams(dev)
    for i, v in enumerate(views):
        down = model.enc_down_dict[v]
        nk = n_keeps[i]
        offs.append(off)
        if nk == 0:  # a view without any visible token contributes nothing to the encoder (tiny grids / high ratios)
            st.embed.append(None)
            continue
        ps = tuple(down.patch_sizes[-1])
        src, sgrid, sidx = sources[i][-1]
        e_in = src.shape[1] * math.prod(ps)
        w_pe = _lin_of(arena, down.patch_embed.proj, train)
        w_li = _lin_of(arena, down.linear, train)
        with vs.view(i):  # views write disjoint token rows of x0
            p16 = torch.empty((b * nk, e_in), dtype=BF16, device=dev)
            _C.gather_patches(src, sgrid, ps, sidx, True, p16)
            t1 = engine.linear_fwd(p16, w_pe)
            t2 = engine.linear_fwd(t1, w_li, out_dtype=F32)
            pos = down.interpolate_pos_encoding(model_grid(model, v, sources[i])).data.reshape(-1, d).contiguous()
            _C.embed_rows(t2.view(b, nk, d), 0, None, pos, keep[i], b, nk, out=x0, out_off=off)
        st.embed.append((p16, t1, w_pe, w_li, ps))
        off += nk
    vs.join()
    st.offs = offs

    st.enc_w = [engine.blockw(arena, blk, train) for blk in model.encoder.blocks]
    cur = x0.view(b * n, d)
    st.enc_saved = []
    for w in st.enc_w:
        cur, sv = engine.block_fwd(cur, w, b, None, train)
        st.enc_saved.append(sv)
    st.enc_norm = engine.normw(arena, model.encoder.norm, train)
    _, enc32, st.enc_mean, st.enc_rstd = engine.ln_fwd(cur, st.enc_norm, want16=False, want32=True, stats=train)
    st.enc_last = cur if train else None
    enc3 = enc32.view(b, n, d)
    if not fuse:
        return None, {"enc": enc3}, st

    total = b + b * sum(n_keeps)
    f16 = torch.empty((total, d), dtype=BF16, device=dev)
    fused32 = {} if want_fused32 else None
    _C.embed_rows(enc3, 0, None, None, None, b, 1, out16=f16[:b].view(b, 1, d))
    if want_fused32:
        fused32["cls"] = enc3[:, :1]
    st.fusion = []
    foff = b
    foffs = []
    idx_rows = [engine.arange_idx(b, offs[i], n_keeps[i], dev) for i in range(len(views))]
    for i, v in enumerate(views):
        fus = model.enc_fusion_dict[v]
        nk = n_keeps[i]
        foffs.append(foff)
        if nk == 0:
            st.fusion.append(None)
            if want_fused32:
                fused32[v] = torch.empty((b, 0, d), dtype=F32, device=dev)
            continue
        nw = engine.normw(arena, fus.norm, train)
        wcs = [_lin_of(arena, conv, train) for conv in fus.down_convs]
        with vs.view(i):  # views write disjoint row segments of f16
            xv = torch.empty((b, nk, d), dtype=F32, device=dev)
            _C.gather_rows(enc3, idx_rows[i], xv)
            cur_v = xv.view(b * nk, d)
            lv = []
            for lvl, conv in enumerate(fus.down_convs):
                k = tuple(conv.kernel_size)
                skip, sgrid, sidx = sources[i][lvl]
                pf = torch.empty((b * nk, skip.shape[1] * math.prod(k)), dtype=BF16, device=dev)
                _C.gather_patches(skip, sgrid, k, sidx, False, pf)
                cur_v = engine.linear_fwd(pf, wcs[lvl], out_dtype=F32, residual=cur_v)
                lv.append((pf, wcs[lvl], k))
            y16 = _rows(f16, foff, b * nk)
            _, y32, mean, rstd = engine.ln_fwd(cur_v, nw, want32=want_fused32, stats=train, y16=y16)
        if want_fused32:
            fused32[v] = y32.view(b, nk, d)
        st.fusion.append((lv, nw, cur_v if train else None, mean, rstd))
        foff += b * nk
    vs.join()
    st.foffs = foffs
    return f16, fused32, st


def _encode_bwd(model, arena, views, st, d_f32, n_keeps, b, targets, fuse=True):
    """Backward of :func:`_encode` given d F (fp32, (B + B * sum n_keep, D)).  targets[v][lvl] = (dst, grid, idx) names the
    pre-zeroed gradient buffer of source level lvl (same addressing as the source) or is None when no gradient is needed;
    contributions are accumulated into it.  ``fuse=False``: ``d_f32`` is the gradient of the encoder output
    (B, 1 + sum n_keep, D) and only the last entry of targets[v] (the embedding source) is used."""
    dev = d_f32.device
    n, d = st.n, st.d
    denc = torch.empty((b, n, d), dtype=F32, device=dev) if fuse else d_f32.contiguous()
    if fuse:
        _C.scatter_rows(d_f32[:b].view(b, 1, d), engine.arange_idx(b, 0, 1, dev), denc)
    idx_rows = [engine.arange_idx(b, st.offs[i], n_keeps[i], dev) for i in range(len(views))]
    vs = engine.ViewStreams(dev)
    for i, _ in enumerate(views if fuse else []):
        nk = n_keeps[i]
        if st.fusion[i] is None:
            continue
        lv, nw, cur_v, mean, rstd = st.fusion[i]
        dyv = _rows(d_f32, st.foffs[i], b * nk)
        with vs.view(i):  # disjoint rows of denc, per-view parameters and gradient targets
            dcur32, dcur16 = engine.ln_bwd(dyv, cur_v, mean, rstd, nw)
            _C.scatter_rows(dcur32.view(b, nk, d), idx_rows[i], denc)
            for lvl, (pf, wc, k) in enumerate(lv):
                tgt = targets[i][lvl]
                dpf = engine.linear_bwd(dcur16, pf, wc, need_dx=tgt is not None)
                if tgt is not None:
                    _C.scatter_patches(dpf, tgt[0], tgt[1], k, tgt[2], False, accumulate=True)
    vs.join()
    ew = st.enc_w
    stage_points = set(encoder_stage_points(len(ew)))
    gb_of = lambda j: engine.out_bias_of(ew, j, d)  # noqa: E731
    dx32, dx16 = engine.ln_bwd(denc.view(b * n, d), st.enc_last, st.enc_mean, st.enc_rstd, st.enc_norm,
                               dxsum=gb_of(len(ew) - 1))
    for j in range(len(ew) - 1, -1, -1):
        dx32, dx16 = engine.block_bwd(dx32, dx16, ew[j], b, st.enc_saved[j], None, None,
                                      fc2_bias_done=gb_of(j) is not None, out_bias=gb_of(j - 1))
        st.enc_saved[j] = None
        if j in stage_points:
            _grad_stage_done(model, f"encoder_{j}")  # blocks[j:] (and, at the first point, encoder.norm + fusion) are final
    dx0 = dx32.view(b, n, d)
    cls = model.encoder.cls_token
    if cls.requires_grad:
        _C.colsum_seg(dx0, 0, 1, arena.grad_view(cls).view(-1))
    for i, _ in enumerate(views):
        nk = n_keeps[i]
        if st.embed[i] is None:
            continue
        p16, t1, w_pe, w_li, ps = st.embed[i]
        with vs.view(i):  # left open on purpose: the caller continues with the stem backward of view i, then joins
            dt2 = torch.empty((b, nk, d), dtype=BF16, device=dev)
            _C.embed_rows(dx0, st.offs[i], None, None, None, b, nk, out16=dt2)
            dt1 = engine.linear_bwd(dt2.view(b * nk, d), t1, w_li)
            tgt = targets[i][-1] if targets[i] else None
            dp = engine.linear_bwd(dt1, p16, w_pe, need_dx=tgt is not None)
            if tgt is not None:
                _C.scatter_patches(dp, tgt[0], tgt[1], ps, tgt[2], True, accumulate=True)
    return vs



def _grad_stage_done(model, stage: str) -> None:
    """Tell a data-parallel trainer that the gradients of a module subtree are final, so that their all-reduce can
    start while the rest of the backward still runs (``MAETrainer`` installs ``model._grad_stage_hook``)."""
    hook = getattr(model, "_grad_stage_hook", None)
    if hook is not None:
        hook(stage)


def encoder_stage_points(depth: int) -> list[int]:
    """Encoder block indices (descending) after whose backward a gradient stage is reported: quarters of the stack.
    (Finer cuts -- down to block 1, so that only block 0 + stems + embeddings remain for the bucket that follows the
    backward -- measured the same at 2 x B200: 26.16 / 26.23 / 26.30 / 26.31 ms per step for the points {6} / {9,6,3} /
    {9,6,3,1} / {8,4,1}, 26.54 ms without any overlap, profiles/r02_scaling.md.)"""
    import os

    env = os.environ.get("CB_STAGE_POINTS")  # e.g. "6" or "9,6,3": override for scaling experiments
    if env is not None:
        return sorted({int(x) for x in env.split(",") if x.strip() and 0 < int(x) < depth}, reverse=True)
    return sorted({depth * 3 // 4, depth // 2, depth // 4} - {0}, reverse=True)


def grad_stages(model) -> dict[str, list[nn.Parameter]]:
    """Parameter sets whose gradients become final at the points reported by :func:`_grad_stage_done`, in the order the
    backward reaches them: "decoder" (decoder-side subtree), then "encoder_k" = encoder blocks [k, previous point) -- the
    first one also carries ``encoder.norm`` and the fusion modules, whose backward has run by then.  Everything else
    (lowest blocks, cls token, stems, embeddings) is final only when the backward ends."""
    stages = {"decoder": [*model.dec_linear.parameters(), *model.dec_embed_dict.parameters(), *model.decoder.parameters(),
                          *model.pred_head_dict.parameters()]}
    blocks = list(model.encoder.blocks)
    prev = len(blocks)
    for k in encoder_stage_points(len(blocks)):
        ps = [p for blk in blocks[k:prev] for p in blk.parameters()]
        if prev == len(blocks):
            ps += [*model.encoder.norm.parameters(), *model.enc_fusion_dict.parameters()]
        stages[f"encoder_{k}"] = ps
        prev = k
    return stages


def _needs(flags, counts):
    out, pos = [], 0
    for c in counts:
        out.append([bool(f) for f in flags[pos:pos + c]])
        pos += c
    return out


def _build_sources(model, arena, views, imgs32, skips, keep, masks, slot, grids, n_keeps, b, train):
    """Gather sources of every view (see :func:`_encode`): the native visible-only stem when the view's stem is the
    reference default (layer norm), else the dense feature maps computed by the caller (cuDNN path)."""
    sources, stems = [], []
    vs = engine.ViewStreams(imgs32[0].device)
    for i, v in enumerate(views):
        down = model.enc_down_dict[v]
        src = _Sources()
        src.token_grid = grids[i]
        if n_keeps[i] == 0:  # nothing of this view is visible: no stem, no sources
            stems.append(None)
            sources.append(src)
            continue
        if len(down.conv_blocks) == 0:
            src.append((imgs32[i], grids[i], keep[i]))
            stems.append(None)
        elif skips[i]:
            for sk in skips[i]:
                src.append((sk, grids[i], keep[i]))
            stems.append(None)
        else:
            levels = stem.stem_weights(arena, down, train)
            mask_i = masks[i].contiguous()
            with vs.view(i):
                geo = stem.Geometry(keep[i], mask_i, slot[i], grids[i], b, n_keeps[i])
                outs, saved = stem.stem_fwd(levels, imgs32[i], geo, train)
            t = b * n_keeps[i]
            unit = (1,) * len(grids[i])
            for lw, x in zip(levels, outs):
                src.append((stem.level_view(x, t, lw.f), unit, None))
            stems.append((levels, geo, saved))
        sources.append(src)
    vs.join()
    return sources, stems


def _build_targets(views, sources, stems, skips, needs):
    """Gradient buffers matching the gather sources: fp32 token-major buffers for the native stem (always needed: the
    stem parameters are upstream), zero-filled dense maps for cuDNN skips that require grad."""
    targets, dskips = [], []
    for i, _ in enumerate(views):
        per_view, grads = [], []
        if stems[i] is not None:
            levels, geo, _ = stems[i]
            t = geo.b * geo.nk
            unit = (1,) * len(geo.grid_tok)
            for lw, (src, _, _) in zip(levels, sources[i]):
                buf = torch.zeros((t * math.prod(lw.f), lw.chans), dtype=F32, device=src.device)
                per_view.append((stem.level_view(buf, t, lw.f), unit, None, buf))
        elif skips[i] and len(sources[i]) == 0:  # dense maps of a view without visible tokens: no gradient
            per_view = [None] * len(skips[i])
            grads = [None] * len(skips[i])
        elif skips[i]:
            for lvl, sk in enumerate(skips[i]):
                if needs[i][lvl]:
                    g = torch.zeros_like(sk)
                    grads.append(g)
                    per_view.append((g, sources[i][lvl][1], sources[i][lvl][2]))
                else:
                    grads.append(None)
                    per_view.append(None)
        else:
            per_view.append(None)  # the image needs no gradient
        targets.append(per_view)
        dskips.append(grads)
    return targets, dskips


def _native_grad_buffers(per_view_targets):
    return [(t[3],) for t in per_view_targets]


class _MAEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, views, images, masks, n_keeps, anchor, *skips_flat):  # noqa: ARG004
        train = any(ctx.needs_input_grad)
        arena = ensure_arena(model)
        arena.refresh_shadow()
        if train:
            arena.prepare_grads()
        nv = len(views)
        counts = model._dense_levels  # per view: number of dense (cuDNN) skip maps passed in, 0 for the native stem
        skips, pos = [], 0
        for c in counts:
            skips.append(list(skips_flat[pos:pos + c]))
            pos += c
        b = images[0].shape[0]
        dev = images[0].device
        imgs32 = [im.detach().to(F32).contiguous() for im in images]
        skips = [[s.detach() for s in sk] for sk in skips]
        grids, keep, drop, slot, n_masks = [], [], [], [], []
        for i, v in enumerate(views):
            down = model.enc_down_dict[v]
            grid = tuple(s // p for s, p in zip(images[i].shape[2:], down.eff_patch_size))
            grids.append(grid)
            kd = _C.mask_to_index(masks[i].contiguous(), n_keeps[i])
            keep.append(kd[0]), drop.append(kd[1]), slot.append(kd[2])
            n_masks.append(math.prod(grid) - n_keeps[i])

        sources, stems = _build_sources(model, arena, views, imgs32, skips, keep, masks, slot, grids, n_keeps, b, train)
        f16, _, st = _encode(model, arena, views, sources, keep, n_keeps, b, train, False)
        d = st.d
        dd = model.dec_linear.weight.shape[0]
        w_dl = _lin_of(arena, model.dec_linear, train)
        y = engine.linear_fwd(f16, w_dl, out_dtype=F32)  # rows: [cls (B) | view 0 (B * n_keep) | ...]

        nk_tot, nm_tot = sum(n_keeps), sum(n_masks)
        cross = model.cross_attn
        nq = 1 + nm_tot if cross else 1 + nk_tot + nm_tot
        xq = torch.empty((b, nq, dd), dtype=F32, device=dev)
        xk16 = torch.empty((b, nk_tot, dd), dtype=BF16, device=dev) if cross else None
        _C.embed_rows(y[:b].view(b, 1, dd), 0, None, None, None, b, 1, out=xq, out_off=0)
        koffs, qoffs = [], []
        koff, moff = 0, 0
        for i, v in enumerate(views):
            emb = model.dec_embed_dict[v]
            if tuple(emb.pos_embed.shape[1:2]) != (math.prod(grids[i]),):
                raise ValueError(f"decoder positional table of view {v} does not match the token grid {grids[i]}")
            dpos = emb.pos_embed.data.view(-1, dd)
            nk, nm = n_keeps[i], n_masks[i]
            yv = _rows(y, st.foffs[i], b * nk).view(b, nk, dd)
            q0 = (1 + moff) if cross else (1 + nk_tot + moff)
            if nk > 0 and cross:
                _C.embed_rows(yv, 0, None, dpos, keep[i], b, nk, out16=xk16, out_off=koff)
            elif nk > 0:
                _C.embed_rows(yv, 0, None, dpos, keep[i], b, nk, out=xq, out_off=1 + koff)
            if nm > 0:
                _C.embed_rows(None, 0, emb.mask_token.data.view(-1), dpos, drop[i], b, nm, out=xq, out_off=q0)
            koffs.append(koff), qoffs.append(q0)
            koff += nk
            moff += nm

        dec_w = [engine.blockw(arena, blk, train) for blk in model.decoder.blocks]
        depth = len(dec_w)
        kv_all = kv_w = None
        kvs = [None] * depth
        if cross:
            h = dec_w[0].n_heads
            hd = dd // h
            xk2 = xk16.view(b * nk_tot, dd)
            kv_w = engine.linw_fused(arena, [blk.attn.kv.weight for blk in model.decoder.blocks],
                                     [blk.attn.kv.bias for blk in model.decoder.blocks], train)
            if nk_tot == 0:  # no visible token anywhere: attention over no keys returns 0 (engine.attn_fwd)
                kv_w = None
                kv_all = [torch.empty((b, 0, 2, h, hd), dtype=BF16, device=dev) for _ in dec_w]
                kvs = [(t[:, :, 0], t[:, :, 1]) for t in kv_all]
            elif kv_w is not None:
                kv_all = engine.linear_fwd(xk2, kv_w).view(b, nk_tot, depth, 2, h, hd)
                kvs = [(kv_all[:, :, j, 0], kv_all[:, :, j, 1]) for j in range(depth)]
            else:
                kv_all = [engine.linear_fwd(xk2, w.kv).view(b, nk_tot, 2, h, hd) for w in dec_w]
                kvs = [(t[:, :, 0], t[:, :, 1]) for t in kv_all]
        cur = xq.view(b * nq, dd)
        dec_saved = []
        for j, w in enumerate(dec_w):
            cur, sv = engine.block_fwd(cur, w, b, kvs[j], train)
            dec_saved.append(sv)
        dec_norm = engine.normw(arena, model.decoder.norm, train)
        dec16, _, dmean, drstd = engine.ln_fwd(cur, dec_norm, stats=train)
        dec3 = dec16.view(b, nq, dd)

        acc = torch.zeros((nv, 8), dtype=F32, device=dev)
        acc[:, 3:5] = float("-inf")
        preds, diffs, heads, dvs = [], [], [], []
        sq_counts, patch_counts = [], []
        q_rows = [engine.arange_idx(b, qoffs[i], n_masks[i], dev) if n_masks[i] > 0 else None for i in range(nv)]
        masks_c = [m.contiguous() for m in masks]
        vs = engine.ViewStreams(dev)
        for i, v in enumerate(views):
            nm = n_masks[i]
            w_ph = _lin_of(arena, model.pred_head_dict[v], train)
            e = w_ph.n
            with vs.view(i):  # per-view head, loss accumulators acc[i]
                dv16 = torch.empty((b, nm, dd), dtype=BF16, device=dev)
                pred = torch.empty((b, nm, e), dtype=F32, device=dev)
                diff = torch.empty((b, nm, e), dtype=F32, device=dev) if train else None
                if nm > 0:
                    _C.gather_rows(dec3, q_rows[i], dv16)
                    engine.linear_fwd(dv16.view(b * nm, dd), w_ph, out=pred.view(b * nm, e))
                _C.masked_mse_fwd(imgs32[i], tuple(model.dec_patch_size_dict[v]), masks_c[i], slot[i], pred,
                                  bool(model.norm_target), 1e-6, acc[i], diff)
            preds.append(pred), diffs.append(diff), heads.append(w_ph), dvs.append(dv16)
            sq_counts.append(b * nm * e)
            patch_counts.append(b * math.prod(grids[i]))
        vs.join()
        out = torch.empty(1 + 5 * nv, dtype=F32, device=dev)
        scales = torch.empty(nv, dtype=F32, device=dev)
        _C.mae_loss_finalize(acc, sq_counts, patch_counts, out, scales)

        if train:
            ctx.state = dict(model=model, arena=arena, views=views, st=st, f16=f16, w_dl=w_dl, keep=keep, n_keeps=n_keeps,
                             n_masks=n_masks, grids=grids, b=b, skips=skips, stems=stems, sources=sources, cross=cross,
                             nq=nq, nk_tot=nk_tot,
                             xk16=xk16, kv_all=kv_all, kv_w=kv_w, kvs=kvs, dec_w=dec_w, dec_saved=dec_saved,
                             dec_norm=dec_norm, dec_last=cur, dmean=dmean, drstd=drstd, diffs=diffs, heads=heads,
                             dvs=dvs, scales=scales, koffs=koffs, qoffs=qoffs, dd=dd, d=d,
                             needs=_needs(ctx.needs_input_grad[6:], counts))
        ctx.mark_non_differentiable(out, *preds)
        return (out[0], out, *preds)

    @staticmethod
    def backward(ctx, g_loss, _g_out, *_g_preds):
        s = ctx.state
        ctx.state = None
        model, arena, views, st = s["model"], s["arena"], s["views"], s["st"]
        b, dd, d, nq, nk_tot = s["b"], s["dd"], s["d"], s["nq"], s["nk_tot"]
        n_keeps, n_masks, keep = s["n_keeps"], s["n_masks"], s["keep"]
        dev = g_loss.device
        sc = (s["scales"] * g_loss.to(F32)).contiguous()
        ddec = torch.zeros((b, nq, dd), dtype=BF16, device=dev)
        q_rows = [engine.arange_idx(b, s["qoffs"][i], n_masks[i], dev) if n_masks[i] > 0 else None for i in range(len(views))]
        vs = engine.ViewStreams(dev)
        for i, _ in enumerate(views):
            nm = n_masks[i]
            if nm == 0:
                continue
            diff = s["diffs"][i]
            with vs.view(i):  # disjoint query rows of ddec
                dpred = torch.empty(diff.shape, dtype=BF16, device=dev)
                _C.scale_cast(diff, dpred, sc[i:i + 1])
                ddv = engine.linear_bwd(dpred.view(b * nm, -1), s["dvs"][i].view(b * nm, dd), s["heads"][i])
                _C.scatter_rows(ddv.view(b, nm, dd), q_rows[i], ddec)
        vs.join()
        dec_w = s["dec_w"]
        dgb_of = lambda j: engine.out_bias_of(dec_w, j, dd)  # noqa: E731
        dx32, dx16 = engine.ln_bwd(ddec.view(b * nq, dd), s["dec_last"], s["dmean"], s["drstd"], s["dec_norm"],
                                   dxsum=dgb_of(len(dec_w) - 1))
        cross = s["cross"]
        dec_w, kvs, kv_all = s["dec_w"], s["kvs"], s["kv_all"]
        depth = len(dec_w)
        dkv_all = None
        if cross:
            if s["kv_w"] is not None:
                dkv_all = torch.empty_like(kv_all)
                dkvs = [(dkv_all[:, :, j, 0], dkv_all[:, :, j, 1]) for j in range(depth)]
            else:
                dkv_all = [torch.empty_like(t) for t in kv_all]
                dkvs = [(t[:, :, 0], t[:, :, 1]) for t in dkv_all]
        for j in range(depth - 1, -1, -1):
            dx32, dx16 = engine.block_bwd(dx32, dx16, dec_w[j], b, s["dec_saved"][j], kvs[j], dkvs[j] if cross else None,
                                          fc2_bias_done=dgb_of(j) is not None, out_bias=dgb_of(j - 1))
            s["dec_saved"][j] = None
        dxq = dx32.view(b, nq, dd)

        total = b + b * nk_tot
        dy16 = torch.empty((total, dd), dtype=BF16, device=dev)
        _C.embed_rows(dxq, 0, None, None, None, b, 1, out16=dy16[:b].view(b, 1, dd))
        if cross and nk_tot == 0:
            dxk3 = None  # no keys, nothing flows back into the (empty) visible-token rows
        elif cross:
            xk2 = s["xk16"].view(b * nk_tot, dd)
            if s["kv_w"] is not None:
                dxk = engine.linear_bwd(dkv_all.view(b * nk_tot, -1), xk2, s["kv_w"])
            else:
                acc32 = torch.zeros((b * nk_tot, dd), dtype=F32, device=dev)
                for j, w in enumerate(dec_w):
                    dkv2 = dkv_all[j].view(b * nk_tot, 2 * dd)
                    if w.kv.gw is not None:
                        _C.gemm(dkv2, xk2, w.kv.gw, a_mn=True, b_mn=True, accumulate=True)
                    if w.kv.gb is not None:
                        _C.colsum(dkv2, w.kv.gb)
                    _C.gemm(dkv2, w.kv.w16, acc32, b_mn=True, accumulate=True)
                dxk = torch.empty((b * nk_tot, dd), dtype=BF16, device=dev)
                _C.cast_bf16(acc32, dxk)
            dxk3 = dxk.view(b, nk_tot, dd)
        for i, v in enumerate(views):
            nk, nm = n_keeps[i], n_masks[i]
            seg = _rows(dy16, st.foffs[i], b * nk).view(b, nk, dd)
            if nk > 0 and cross:
                _C.gather_rows(dxk3, engine.arange_idx(b, s["koffs"][i], nk, dev), seg)
            elif nk > 0:
                _C.embed_rows(dxq, 1 + s["koffs"][i], None, None, None, b, nk, out16=seg)
            mt = model.dec_embed_dict[v].mask_token
            if mt.requires_grad and nm > 0:
                _C.colsum_seg(dxq, s["qoffs"][i], nm, arena.grad_view(mt).view(-1))
        d_f32 = engine.linear_bwd(dy16, s["f16"], s["w_dl"], dx_dtype=F32)
        _grad_stage_done(model, "decoder")  # dec_linear, decoder, decoder embeddings and prediction heads are final
        targets, dskips = _build_targets(views, s["sources"], s["stems"], s["skips"], s["needs"])
        vs = _encode_bwd(model, arena, views, st, d_f32, n_keeps, b, targets)
        for i, stem_state in enumerate(s["stems"]):
            if stem_state is not None:
                levels, geo, saved = stem_state
                with vs.view(i):  # same side stream as the embedding backward of this view
                    stem.stem_bwd(levels, geo, saved, [t[0] for t in _native_grad_buffers(targets[i])])
        vs.join()
        flat = [g for per_view in dskips for g in per_view]
        return (None, None, None, None, None, None, *flat)


class _DirectCtx:
    """Stand-in for the autograd context when ``_MAEFn.forward`` / ``backward`` are called directly (train_step)."""

    needs_input_grad = (False, False, False, False, False, True)  # the parameter anchor

    def __init__(self) -> None:
        self.state = None

    def mark_non_differentiable(self, *_):
        pass


# ------------------------------------------------------------------------------------------
# the module
# ------------------------------------------------------------------------------------------
def _uniform_mask_count(m: torch.Tensor, view: str) -> int:
    """Number of removed patches per sample of a user-supplied mask.  Every sample must remove the same number: the
    reference's ``x[~mask].reshape(batch, n_keep, -1)`` (cinema/mae/mae.py:550) raises otherwise, and so does this (one
    host read; masks drawn by ``forward`` itself have static counts and never come through here)."""
    counts = m.sum(1)
    n = int(counts[0])
    if not bool((counts == n).all()):
        raise ValueError(f"enc_mask_dict[{view!r}] removes a different number of patches per sample: {counts.tolist()}")
    return n


class CineMA(nn.Module):
    """Cine masked autoencoder (cinema/mae/mae.py:285-642) -- same constructor, attributes and state dict."""

    def __init__(self, image_size_dict, in_chans_dict, enc_patch_size_dict, enc_scale_factor_dict, enc_conv_chans,
                 enc_conv_n_blocks, enc_embed_dim, enc_depth, enc_n_heads, dec_embed_dim, dec_depth, dec_n_heads,
                 mlp_ratio=4, qkv_bias=True, norm_target=False, cross_attn=True, norm_layer=nn.LayerNorm, norm_eps=1e-5,
                 rotary=False, act_layer=nn.GELU, mlp_layer=Mlp, drop_path=0.0, norm="layer") -> None:
        super().__init__()
        self.grad_ckpt = False
        self.native_stem = True  # False: run the conv stem densely through cuDNN (reference evaluation order)
        self._dense_levels: list[int] = []
        self.norm_target = norm_target
        self.views = list(image_size_dict.keys())
        self.enc_down_dict = nn.ModuleDict({
            v: DownsampleEncoder(image_size=image_size_dict[v], in_chans=in_chans_dict[v], patch_size=enc_patch_size_dict[v],
                                 scale_factor=enc_scale_factor_dict[v], conv_chans=enc_conv_chans,
                                 conv_n_blocks=enc_conv_n_blocks, embed_dim=enc_embed_dim, norm=norm)
            for v in self.views
        })
        self.enc_fusion_dict = nn.ModuleDict({
            v: MultiScaleFusion(image_size=image_size_dict[v], patch_size=enc_patch_size_dict[v],
                                scale_factor=enc_scale_factor_dict[v], conv_chans=enc_conv_chans, embed_dim=enc_embed_dim,
                                norm_layer=norm_layer, norm_eps=norm_eps)
            for v in self.views
        })
        self.encoder = ViTEncoder(embed_dim=enc_embed_dim, depth=enc_depth, n_heads=enc_n_heads, mlp_ratio=mlp_ratio,
                                  qkv_bias=qkv_bias, norm_layer=norm_layer, norm_eps=norm_eps, rotary=rotary,
                                  act_layer=act_layer, mlp_layer=mlp_layer, drop_path=drop_path)
        self.dec_linear = Linear(enc_embed_dim, dec_embed_dim)
        self.dec_embed_dict = nn.ModuleDict({
            v: DecoderEmbedding(enc_grid_size=self.enc_down_dict[v].patch_embed.grid_size, dec_embed_dim=dec_embed_dim,
                                add_embed_token=False)
            for v in self.views
        })
        self.cross_attn = cross_attn
        self.decoder = ViTDecoder(embed_dim=dec_embed_dim, depth=dec_depth, n_heads=dec_n_heads, mlp_ratio=mlp_ratio,
                                  qkv_bias=qkv_bias, norm_layer=norm_layer, norm_eps=norm_eps, rotary=rotary,
                                  act_layer=act_layer, mlp_layer=mlp_layer, drop_path=drop_path)
        self.dec_patch_size_dict = {
            v: get_decoder_patch_size(image_size=image_size_dict[v], n_conv_layers=len(enc_conv_chans),
                                      enc_patch_size=enc_patch_size_dict[v], enc_scale_factor=enc_scale_factor_dict[v])
            for v in self.views
        }
        self.pred_head_dict = nn.ModuleDict({
            v: Linear(dec_embed_dim, math.prod(ps) * in_chans_dict[v]) for v, ps in self.dec_patch_size_dict.items()
        })
        self.apply(init_weights)

    # -------------------------------------------------------------- plumbing
    def _arena_groups(self):
        """Cross-attention decoder: lay the kv projections of all blocks back to back (one GEMM, N = depth * 2D)."""
        if not self.cross_attn:
            return []
        blocks = list(self.decoder.blocks)
        if any(blk.attn.kv.bias is None for blk in blocks):
            return []
        return [[blk.attn.kv.weight for blk in blocks], [blk.attn.kv.bias for blk in blocks]]

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        """API compatibility (cinema/mae/mae.py:444-455); nothing is recomputed on the B200 path."""
        self.grad_ckpt = enable
        for v in self.views:
            self.enc_down_dict[v].set_grad_ckpt(enable)
            self.enc_fusion_dict[v].set_grad_ckpt(enable)
        self.encoder.set_grad_ckpt(enable)
        self.dec_linear.set_grad_ckpt(enable)
        self.decoder.set_grad_ckpt(enable)
        for v in self.views:
            self.pred_head_dict[v].set_grad_ckpt(enable)

    def _check_views(self, image_dict) -> list[str]:
        views = list(image_dict.keys())
        if any(v not in self.views for v in views):
            raise ValueError(f"views {views} must be in self.input_keys {self.views}.")
        return views

    def _stem(self, views, image_dict, masks):
        """Dense (cuDNN, bf16 autocast) stem for the views whose stem the native visible-only path does not cover
        (non-default ``norm``); an empty list for the others.  -> per-view lists of feature maps."""
        first = image_dict[views[0]]
        out = []
        with torch.autocast(device_type="cuda", dtype=BF16, enabled=first.is_cuda):
            for i, v in enumerate(views):
                down = self.enc_down_dict[v]
                if self.native_stem and stem.supported(down):
                    out.append([])
                else:
                    out.append(down.conv_stem(image_dict[v], None if masks is None else masks[i]))
        self._dense_levels = [len(x) for x in out]
        return out

    # -------------------------------------------------------------- forward paths
    def feature_forward(self, image_dict: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        """No masking, all tokens: {"cls": (B, 1, D), view: (B, n_patches, D)} (cinema/mae/mae.py:457-502).
        Inference path (no gradient)."""
        views = self._check_views(image_dict)
        with torch.no_grad():
            arena = ensure_arena(self)
            arena.refresh_shadow()
            skips = self._stem(views, image_dict, None)
            b = image_dict[views[0]].shape[0]
            dev = image_dict[views[0]].device
            grids, keep, n_keeps, masks, slots = [], [], [], [], []
            for i, v in enumerate(views):
                down = self.enc_down_dict[v]
                grid = tuple(s // p for s, p in zip(image_dict[v].shape[2:], down.eff_patch_size))
                n = math.prod(grid)
                grids.append(grid)
                n_keeps.append(n)
                keep.append(engine.arange_idx(b, 0, n, dev))
                slots.append(keep[-1])  # every token is visible: slot == token id
                masks.append(torch.zeros((b, n), dtype=torch.bool, device=dev))
            imgs32 = [image_dict[v].to(F32).contiguous() for v in views]
            sources, _ = _build_sources(self, arena, views, imgs32, skips, keep, masks, slots, grids, n_keeps, b, False)
            _, fused, _ = _encode(self, arena, views, sources, keep, n_keeps, b, False, True)
        return fused

    def forward(self, image_dict: dict[str, torch.Tensor], enc_mask_ratio: float,
                enc_mask_dict: dict[str, torch.Tensor] | None = None):
        """-> (loss, pred_dict, enc_mask_dict, metrics), as cinema/mae/mae.py:504-612.

        ``enc_mask_dict`` (extension, default ``None``): use the given (B, n_patches) boolean masks instead of
        drawing them, e.g. to compare against another implementation on identical masks."""
        views = self._check_views(image_dict)
        first = image_dict[views[0]]
        b, dev = first.shape[0], first.device
        masks, n_keeps = [], []
        for v in views:
            down = self.enc_down_dict[v]
            grid = tuple(s // p for s, p in zip(image_dict[v].shape[2:], down.eff_patch_size))
            n_patches = math.prod(grid)
            if enc_mask_dict is not None:
                m = enc_mask_dict[v].to(device=dev, dtype=torch.bool)
                n_keeps.append(n_patches - _uniform_mask_count(m, v))
            else:
                m = get_batch_random_patch_mask(b, n_patches, enc_mask_ratio, dev)
                n_keeps.append(n_patches if enc_mask_ratio == 0 else int(n_patches * (1 - enc_mask_ratio)))
            masks.append(m)
        skips = self._stem(views, image_dict, masks)
        flat = [s for per_view in skips for s in per_view]
        anchor = next((p for p in self.parameters() if p.requires_grad), None)
        if anchor is None:
            anchor = first.new_zeros(())
        loss, out, *preds = _MAEFn.apply(self, views, [image_dict[v] for v in views], masks, n_keeps, anchor, *flat)
        metrics = {}
        for i, v in enumerate(views):
            metrics[f"{v}_target_mean"] = out[2 + 5 * i]
            metrics[f"{v}_target_std"] = out[3 + 5 * i]
            metrics[f"{v}_mse_loss"] = out[1 + 5 * i]
            if self.norm_target and preds[i].shape[1] > 0:
                metrics[f"{v}_normed_target_max"] = out[4 + 5 * i]
                metrics[f"{v}_pred_max"] = out[5 + 5 * i]
        metrics["loss"] = loss
        return loss, dict(zip(views, preds)), dict(zip(views, masks)), metrics

    def direct_step_supported(self) -> bool:
        """Can :meth:`train_step` be used (every stem runs natively, at least one trainable parameter)?"""
        return (self.native_stem and all(stem.supported(self.enc_down_dict[v]) for v in self.views)
                and any(p.requires_grad for p in self.parameters()))

    @torch.no_grad()
    def train_step(self, image_dict: dict[str, torch.Tensor], enc_mask_ratio: float,
                   enc_mask_dict: dict[str, torch.Tensor] | None = None) -> torch.Tensor:
        """forward + backward of ``forward(image_dict, enc_mask_ratio)[0]`` without the autograd engine: the fused
        forward and its hand-written backward are called back to back on the calling thread (same kernels, same RNG
        draws, gradients accumulated into the ``.grad`` arena views).  Returns the detached loss.  This is what
        ``MAETrainer`` runs: with no engine thread in between, the step can be captured as SEVERAL CUDA graphs split at
        the gradient stages (``_grad_stage_done``), which is how the gradient all-reduce overlaps the backward."""
        if not self.direct_step_supported():
            raise RuntimeError("train_step needs the native stem for every view and a trainable parameter")
        views = self._check_views(image_dict)
        first = image_dict[views[0]]
        b, dev = first.shape[0], first.device
        masks, n_keeps = [], []
        for v in views:
            down = self.enc_down_dict[v]
            grid = tuple(s // p for s, p in zip(image_dict[v].shape[2:], down.eff_patch_size))
            n_patches = math.prod(grid)
            if enc_mask_dict is not None:  # given masks (tests / comparisons), as in forward()
                masks.append(enc_mask_dict[v].to(device=dev, dtype=torch.bool))
                n_keeps.append(n_patches - _uniform_mask_count(masks[-1], v))
            else:
                masks.append(get_batch_random_patch_mask(b, n_patches, enc_mask_ratio, dev))
                n_keeps.append(n_patches if enc_mask_ratio == 0 else int(n_patches * (1 - enc_mask_ratio)))
        self._dense_levels = [0] * len(views)
        ctx = _DirectCtx()
        loss, *_ = _MAEFn.forward(ctx, self, views, [image_dict[v] for v in views], masks, n_keeps, None)
        _MAEFn.backward(ctx, torch.ones((), dtype=F32, device=dev), None)
        return loss

    @classmethod
    def from_pretrained(cls, config: dict | None = None, state_dict: dict | None = None, **kwargs) -> "CineMA":
        """Build from a reference-format config (``model`` section of cinema/mae/config.yaml) and a state dict in the
        reference key schema.  The reference downloads both from the Hugging Face hub (cinema/mae/mae.py:614-642);
        this image has no network, so they are passed in (or loaded from local files via ``config_path`` /
        ``weights_path``)."""
        if config is None:
            import yaml

            with open(kwargs["config_path"]) as f:
                config = yaml.safe_load(f)
        model = get_model(config)
        if state_dict is None and "weights_path" in kwargs:
            from safetensors.torch import load_file

            state_dict = load_file(kwargs["weights_path"])
        if state_dict is not None:
            model.load_state_dict(state_dict)
        return model


class _Cfg:
    """Attribute / key access over nested dicts or OmegaConf-like objects."""

    def __init__(self, node) -> None:
        self._node = node

    def __getattr__(self, key):
        node = self._node
        val = node[key] if isinstance(node, dict) else getattr(node, key)
        return _Cfg(val) if isinstance(val, dict) or (hasattr(val, "keys") and not isinstance(val, (list, tuple))) else val


def get_model(config) -> CineMA:
    """Config -> CineMA, as cinema/mae/mae.py:231-282: four views (SAX 3-D, LAX 2C/3C/4C 2-D) with sizes from
    ``config.data``, stem geometry from ``config.model`` and ViT dims from ``get_vit_config(config.model.size)``.
    ``config`` may be a nested dict (``yaml.safe_load`` of the reference's config.yaml) or an OmegaConf object."""
    c = _Cfg(config)
    lax = ("lax_2c", "lax_3c", "lax_4c")
    sax_size, lax_size = tuple(c.data.sax.patch_size), tuple(c.data.lax.patch_size)
    patch, scale = tuple(c.model.patch_size), tuple(c.model.scale_factor)
    model = CineMA(
        image_size_dict={"sax": sax_size, **{v: lax_size for v in lax}},
        in_chans_dict={"sax": c.data.sax.in_chans, **{v: c.data.lax.in_chans for v in lax}},
        enc_patch_size_dict={"sax": patch, **{v: patch[:2] for v in lax}},
        enc_scale_factor_dict={"sax": scale, **{v: scale[:2] for v in lax}},
        enc_conv_chans=list(c.model.enc_conv_chans), enc_conv_n_blocks=c.model.enc_conv_n_blocks,
        **get_vit_config(c.model.size),
    )
    model.set_grad_ckpt(c.grad_ckpt)
    return model
