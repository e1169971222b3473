"""CPU emulation of the ``cinema_b200._C`` entry points.  TEST INFRASTRUCTURE ONLY.

The product has no CPU path: ``cinema_b200`` fails loudly without the CUDA library and a GPU.
To exercise the *host logic* (layout bookkeeping, index arithmetic, the hand-written backward
chains, the parameter arena) in the GPU-less build container, the ``emulated_kernels`` fixture of
``tests/conftest.py`` monkey-patches every ``_C`` wrapper with the plain-torch restatement below,
which follows the semantics documented in ``include/cinema_b200.h`` (including the bf16 rounding
points).  The GPU tests (``-m gpu``) never use this module.
"""

from __future__ import annotations

import math

import torch

BF16 = torch.bfloat16
EPI_NONE, EPI_GELU, EPI_GELU_BWD = 0, 1, 2


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


def _gelu_grad(x):
    cdf = 0.5 * (1.0 + torch.erf(x * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * x * x)
    return cdf + x * pdf


def gemm(a, b, out, *, a_mn=False, b_mn=False, accumulate=False, out2=None, bias=None, residual=None, aux=None,
         epilogue=EPI_NONE, alpha=1.0, split_k=0, block_n=0, colsum=None, row_scale=None, rows_per_group=0):
    assert a.dtype == BF16 and b.dtype == BF16
    am = a.float().t() if a_mn else a.float()
    bm = b.float() if b_mn else b.float().t()  # (K, N)
    acc = (am.double() @ bm.double()).float() * alpha
    if bias is not None:
        acc = acc + bias.float()
    if epilogue == EPI_GELU:
        pre = acc.to(BF16)
        if out is not None:
            out.copy_(pre)
        out2.copy_(_gelu(pre.float()).to(BF16))
        return
    if epilogue == EPI_GELU_BWD:
        acc = acc * _gelu_grad(aux.float())
    if row_scale is not None:
        assert epilogue == EPI_NONE and not accumulate and rows_per_group > 0
        acc = acc * row_scale.float().repeat_interleave(rows_per_group)[:acc.shape[0], None]
    if residual is not None:
        acc = acc + residual
    if accumulate:
        assert out.dtype == torch.float32
        out.add_(acc)
    else:
        out.copy_(acc.to(out.dtype))
    if colsum is not None:
        assert out.dtype == BF16 and epilogue != EPI_GELU
        colsum.add_(acc.to(BF16).float().sum(0))
    if out2 is not None:
        out2.copy_(acc.to(BF16))


def conv_gemm(x_rows, w_taps, out, row_off, *, bias=None, residual=None, block_n=0):
    """out[r] = sum_t x_rows[r + off_t] @ W_t^T: rows outside the matrix read as zero (what the TMA producer delivers)."""
    r, c_in = x_rows.shape
    x = x_rows.float()
    acc = torch.zeros((r, w_taps.shape[0]), dtype=torch.float64)
    for t, off in enumerate(row_off):
        sh = torch.zeros_like(x)
        lo, hi = max(0, -off), min(r, r - off)
        if hi > lo:
            sh[lo:hi] = x[lo + off:hi + off]
        acc += sh.double() @ w_taps[:, t * c_in:(t + 1) * c_in].double().t()
    acc = acc.float()
    if bias is not None:
        acc = acc + bias.float()
    if residual is not None:
        acc = acc + residual
    out.copy_(acc.to(out.dtype))


def colsum(x, out):
    out.add_(x.float().sum(0))


def attention_fwd(q, k, v, o, lse, scale):
    b, nq, h, d = q.shape
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    l = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - l[..., None])
    o.copy_(torch.einsum("bhqk,bkhd->bqhd", p.to(BF16).float(), v.float()).to(BF16))
    lse.view(b, h, nq).copy_(l)


def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dq_acc, scale, dq_colsum=None, dk_colsum=None, dv_colsum=None):
    b, nq, h, d = q.shape
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    p = torch.exp(s - lse.view(b, h, nq)[..., None])
    dlt = (o.float() * do.float()).sum(-1).permute(0, 2, 1)  # (b, h, q)
    dp = torch.einsum("bqhd,bkhd->bhqk", do.float(), v.float())
    ds = (p * (dp - dlt[..., None]) * scale).to(BF16).float()
    dv.copy_(torch.einsum("bhqk,bqhd->bkhd", p.to(BF16).float(), do.float()).to(BF16))
    dk.copy_(torch.einsum("bhqk,bqhd->bkhd", ds, q.float()).to(BF16))
    dq.copy_(torch.einsum("bhqk,bkhd->bqhd", ds, k.float()).to(BF16))
    for cs, t in ((dq_colsum, dq), (dk_colsum, dk), (dv_colsum, dv)):
        if cs is not None:
            cs += t.float().sum((0, 1)).reshape(-1)


def layernorm_fwd(x, gamma, beta, eps, y16=None, y32=None, mean=None, rstd=None, act=False):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    r = torch.rsqrt(var + eps)
    y = (x - mu) * r * gamma + beta
    if act:
        y = _gelu(y)
    if y32 is not None:
        y32.copy_(y)
    if y16 is not None:
        y16.copy_(y.to(BF16))
    if mean is not None:
        mean.copy_(mu.squeeze(-1))
    if rstd is not None:
        rstd.copy_(r.squeeze(-1))


def layernorm_bwd(dy, x, mean, rstd, gamma, dres=None, dx32=None, dx16=None, dgamma=None, dbeta=None, beta_act=None,
                  dxsum=None):
    d = dy.float()
    xh = (x - mean[:, None]) * rstd[:, None]
    if beta_act is not None:
        d = d * _gelu_grad(xh * gamma + beta_act)
    dyg = d * gamma
    m1 = dyg.mean(-1, keepdim=True)
    m2 = (dyg * xh).mean(-1, keepdim=True)
    o = rstd[:, None] * (dyg - m1 - xh * m2)
    if dres is not None:
        o = o + dres
    if dgamma is not None:
        dgamma.add_((d * xh).sum(0))
    if dbeta is not None:
        dbeta.add_(d.sum(0))
    if dx32 is not None:
        dx32.copy_(o)
    if dx16 is not None:
        dx16.copy_(o.to(BF16))
    if dxsum is not None:
        dxsum.add_(o.to(BF16).float().sum(0))


def cast_bf16(src, dst):
    dst.copy_(src.to(BF16))


def mask_to_index(mask, n_keep):
    b, n = mask.shape
    order = torch.arange(n).expand(b, n)
    keep = order[~mask].reshape(b, n_keep).int()
    drop = order[mask].reshape(b, n - n_keep).int()
    slot = torch.empty(b, n, dtype=torch.int32)
    pos_k = torch.arange(n_keep, dtype=torch.int32).expand(b, n_keep)
    pos_d = torch.arange(n - n_keep, dtype=torch.int32).expand(b, n - n_keep)
    slot.scatter_(1, keep.long(), pos_k)
    slot.scatter_(1, drop.long(), pos_d)
    return keep.contiguous(), drop.contiguous(), slot


def gather_rows(src, idx, out, out_off=0):
    b, k = idx.shape
    s = src if src.dim() == 3 else src[None]
    if s.shape[0] == 1 and b > 1:
        s = s.expand(b, -1, -1)
    out[:, out_off:out_off + k] = torch.gather(s, 1, idx.long()[..., None].expand(-1, -1, s.shape[-1]))


def scatter_rows(src, idx, dst, src_off=0):
    b, k = idx.shape
    dst.scatter_(1, idx.long()[..., None].expand(-1, -1, src.shape[-1]), src[:, src_off:src_off + k])


def embed_rows(a, a_off, row, table, idx, b, k, out=None, out16=None, out_off=0):
    ref = out if out is not None else out16
    d = ref.shape[-1]
    r = torch.zeros(b, k, d)
    if table is not None:
        r = r + table.reshape(-1, d)[idx.long()]
    if a is not None:
        r = r + a[:, a_off:a_off + k]
    if row is not None:
        r = r + row.reshape(d)
    if out is not None:
        out[:, out_off:out_off + k] = r
    if out16 is not None:
        out16[:, out_off:out_off + k] = r.to(BF16)


def colsum_seg(x, off, k, out):
    out.add_(x[:, off:off + k].sum(dim=(0, 1)))


def scale_cast(src, dst, scale_dev=None, scale=1.0, group=0):
    if group:
        f = scale_dev.float().repeat_interleave(group).view(src.shape)
        dst.copy_((src * (scale * f)).to(BF16))
        return
    s = scale * (float(scale_dev.reshape(-1)[0]) if scale_dev is not None else 1.0)
    dst.copy_((src * s).to(BF16))


def mae_loss_finalize(acc, sq_count, patch_count, out, scales):
    v = acc.shape[0]
    mses = []
    for i in range(v):
        mse = acc[i, 0] / sq_count[i] if sq_count[i] > 0 else torch.tensor(float("nan"))
        out[1 + 5 * i] = mse
        out[2 + 5 * i] = acc[i, 1] / patch_count[i]
        out[3 + 5 * i] = acc[i, 2] / patch_count[i]
        out[4 + 5 * i] = acc[i, 3]
        out[5 + 5 * i] = acc[i, 4]
        mses.append(mse)
    fin = [bool(torch.isfinite(m)) for m in mses]
    n_fin = sum(fin)
    out[0] = sum(m for m, f in zip(mses, fin) if f) / n_fin if n_fin else float("nan")
    for i in range(v):
        scales[i] = 2.0 / (sq_count[i] * n_fin) if fin[i] and n_fin else 0.0


def _patchify(image, patch):
    n = len(patch)
    b, c, *sp = image.shape
    grid = [s // p for s, p in zip(sp, patch)]
    split = []
    for g, p in zip(grid, patch):
        split += [g, p]
    x = image.reshape(b, c, *split)
    x = x.permute(0, *[2 + 2 * i for i in range(n)], *[3 + 2 * i for i in range(n)], 1)
    return x.reshape(b, math.prod(grid), math.prod(patch) * c)


def _unpatchify(x, patch, grid, c):
    n = len(patch)
    b = x.shape[0]
    x = x.reshape(b, *grid, *patch, c)
    order = [0, 2 * n + 1]
    for i in range(n):
        order += [1 + i, 1 + n + i]
    return x.permute(*order).reshape(b, c, *[g * p for g, p in zip(grid, patch)])


def patchify(src, dst, b, c, spatial, patch, inverse):
    grid = [s // p for s, p in zip(spatial, patch)]
    if not inverse:
        dst.copy_(_patchify(src.reshape(b, c, *spatial), patch).reshape(dst.shape))
    else:
        e = math.prod(patch) * c
        dst.copy_(_unpatchify(src.reshape(b, math.prod(grid), e), patch, grid, c).reshape(dst.shape))


def _tokens(src, grid, patch, chan_last):
    """all tokens of a (B, C, *spatial) tensor: (B, n_tok, E) in either element order."""
    b, c = src.shape[:2]
    tok = _patchify(src, patch)  # (..., off, c) order
    if not chan_last:
        p = math.prod(patch)
        tok = tok.reshape(b, -1, p, c).transpose(2, 3).reshape(b, -1, p * c)
    return tok


def gather_patches(src, grid, patch, idx, chan_last, out):
    tok = _tokens(src.float(), grid, patch, chan_last)
    if idx is not None:
        tok = torch.gather(tok, 1, idx.long()[..., None].expand(-1, -1, tok.shape[-1]))
    out.copy_(tok.reshape(out.shape).to(out.dtype))


def scatter_patches(rows, dst, grid, patch, idx, chan_last, accumulate=False):
    if accumulate:
        tmp = torch.zeros_like(dst)
        scatter_patches(rows, tmp, grid, patch, idx, chan_last)
        dst.add_(tmp)
        return
    b, c = dst.shape[:2]
    n_tok = math.prod(grid)
    p = math.prod(patch)
    e = p * c
    r = rows.float().reshape(b, -1, e)
    if idx is not None:
        full = _tokens(dst.float(), grid, patch, chan_last).clone()
        full.scatter_(1, idx.long()[..., None].expand(-1, -1, e), r)
    else:
        full = r
    if not chan_last:
        full = full.reshape(b, n_tok, c, p).transpose(2, 3).reshape(b, n_tok, e)
    dst.copy_(_unpatchify(full, patch, grid, c).to(dst.dtype))


def masked_mse_fwd(image, patch, mask, slot, pred, norm_target, eps, acc, diff):
    b = image.shape[0]
    tgt = _patchify(image, patch)
    mean = tgt.mean(-1, keepdim=True)
    std = tgt.var(-1, keepdim=True) ** 0.5
    acc[1] += mean.sum()
    acc[2] += std.sum()
    if norm_target:
        tgt = (tgt - mean) / (std + eps)
    t = tgt[mask].reshape(pred.shape)
    d = pred - t
    acc[0] += (d * d).sum()
    if norm_target and t.numel() > 0:
        acc[3] = torch.maximum(acc[3], t.max())
        acc[4] = torch.maximum(acc[4], pred.max())
    if diff is not None:
        diff.copy_(d)


def rope_apply(x, cos, sin, transpose=False):
    n = x.shape[1]
    ro = cos.shape[1] * 2
    c = torch.cat([cos[:n], cos[:n]], -1)[:, None, :]
    s = torch.cat([sin[:n], sin[:n]], -1)[:, None, :] * (-1.0 if transpose else 1.0)
    head = x[..., :ro].float()
    half = ro // 2
    rot = torch.cat([-head[..., half:], head[..., :half]], -1)
    return torch.cat([(head * c + rot * s).to(x.dtype), x[..., ro:]], -1)


def expand_token_index(keep, grid_tok, f):
    b, nk = keep.shape
    nd = len(f)
    t = keep.long()
    tg = []
    for a in range(nd - 1, -1, -1):
        tg.insert(0, t % grid_tok[a])
        t = t // grid_tok[a]
    p = math.prod(f)
    pid = torch.arange(p)
    pa = []
    for a in range(nd - 1, -1, -1):
        pa.insert(0, pid % f[a])
        pid = pid // f[a]
    idx = torch.zeros(b, nk, p, dtype=torch.long)
    for a in range(nd):
        idx = idx * (grid_tok[a] * f[a]) + tg[a][..., None] * f[a] + pa[a]
    return idx.reshape(b, nk * p).int()


def _dense_from_tokens(x, keep, grid_tok, f, c):
    """(B*nk*P, C) token-major rows -> dense (B, C, *level_grid) with zeros at masked tokens, plus the flat index."""
    b, nk = keep.shape
    idx = expand_token_index(keep, grid_tok, f).long()  # (B, nk*P)
    level = [g * ff for g, ff in zip(grid_tok, f)]
    dense = torch.zeros(b, math.prod(level), c)
    dense.scatter_(1, idx[..., None].expand(-1, -1, c), x.float().reshape(b, -1, c))
    return dense.transpose(1, 2).reshape(b, c, *level), idx, level


def dwconv_tokens(x, out, w, bias, mask, slot, keep, grid_tok, f, transpose=False):
    import torch.nn.functional as F

    c = x.shape[-1]
    dense, idx, level = _dense_from_tokens(x, keep, grid_tok, f, c)
    wf = w.float()
    if transpose:
        wf = wf.flip(dims=list(range(2, wf.dim())))
    conv = F.conv2d if len(f) == 2 else F.conv3d
    y = conv(dense, wf, bias.float() if bias is not None else None, padding=2, groups=c)
    y = y.reshape(y.shape[0], c, -1).transpose(1, 2)
    out.copy_(torch.gather(y, 1, idx[..., None].expand(-1, -1, c)).reshape(out.shape).to(BF16))


def dwconv_tokens_wgrad(x, dy, dw, db, mask, slot, keep, grid_tok, f):
    import torch.nn.functional as F

    c = x.shape[-1]
    dense, idx, level = _dense_from_tokens(x, keep, grid_tok, f, c)
    ddense, _, _ = _dense_from_tokens(dy, keep, grid_tok, f, c)
    wz = torch.zeros_like(dw).requires_grad_()
    conv = F.conv2d if len(f) == 2 else F.conv3d
    with torch.enable_grad():
        y = conv(dense, wz, None, padding=2, groups=c)
        (y * ddense).sum().backward()
    dw.add_(wz.grad)
    if db is not None:
        db.add_(dy.float().sum(0))


def sumsq(x, out):
    out[0] += (x.double() ** 2).sum().float()


def adamw_flat(p, g, m, v, p16, hyper, beta1, beta2, eps, weight_decay, gnorm_sq, max_norm, grad_scale):
    lr, bc1, bc2 = (float(hyper[i]) for i in range(3))
    gs = grad_scale
    if gnorm_sq is not None:
        norm = float(gnorm_sq.reshape(-1)[0]) ** 0.5 * grad_scale
        if not math.isfinite(norm):
            return
        if max_norm > 0:
            gs *= min(1.0, max_norm / (norm + 1e-6))
    gg = g * gs
    m.mul_(beta1).add_(gg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.mul_(1 - lr * weight_decay).addcdiv_(m, denom, value=-lr / bc1)
    if p16 is not None:
        p16.copy_(p.to(BF16))


def _seg_restated(logits, labels):
    import torch.nn.functional as F

    lg = logits.float()
    lab = labels.reshape(lg.shape[0], *lg.shape[2:]).long()
    n_cls = lg.shape[1]
    ce = F.cross_entropy(lg, lab, ignore_index=-1)
    onehot = F.one_hot(lab.clamp(min=0), n_cls).movedim(-1, 1).to(lg.dtype)
    prob = lg.softmax(dim=1)
    dims = tuple(range(2, lg.dim()))
    inter = (prob[:, 1:] * onehot[:, 1:]).sum(dims)
    denom = prob[:, 1:].sum(dims) + onehot[:, 1:].sum(dims)
    dice = (1.0 - (2.0 * inter + 1e-5) / (denom + 1e-5)).mean()
    return ce, dice


def seg_loss_fwd(logits, labels):
    ce, dice = _seg_restated(logits.detach(), labels)
    b, c = logits.shape[:2]
    return torch.stack([ce + dice, ce, dice]).float(), torch.zeros(b * c * 2 + 1)


def seg_loss_bwd(logits, labels, coef, grad_out):
    with torch.enable_grad():  # (called from inside an autograd backward, where grad mode is off)
        lg = logits.detach().float().requires_grad_(True)
        ce, dice = _seg_restated(lg, labels)
        (g,) = torch.autograd.grad(ce + dice, lg)
    return (g * grad_out.reshape(())).to(logits.dtype)


def device_info():
    return 148, 10, 0


ALL = [
    "conv_gemm",
    "gemm", "colsum", "attention_fwd", "attention_bwd", "layernorm_fwd", "layernorm_bwd", "cast_bf16", "mask_to_index",
    "gather_rows", "scatter_rows", "embed_rows", "colsum_seg", "scale_cast", "mae_loss_finalize", "patchify",
    "gather_patches", "scatter_patches", "masked_mse_fwd", "device_info", "seg_loss_fwd", "seg_loss_bwd", "rope_apply", "sumsq", "adamw_flat", "expand_token_index", "dwconv_tokens", "dwconv_tokens_wgrad",
]
