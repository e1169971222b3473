"""Host-side orchestration of the sm_100a kernels: linear / LayerNorm / attention / transformer-block
forward and hand-written backward chains over 2-D row-major activations.

Precision contract (mirrors what ``torch.autocast("cuda", bf16)`` does to the reference modules,
SURVEY.md section 3.2): the residual stream is fp32; LayerNorm reads fp32 and its output is consumed
as bf16; every Linear reads bf16 activations and bf16 weights, accumulates in fp32 (TMEM) and
emits bf16, except where its result is added to the fp32 residual stream, which is fused into
the GEMM epilogue; attention runs in bf16 with fp32 softmax statistics.  Weight gradients are
accumulated in fp32 directly into the flat gradient arena.

Nothing here falls back to torch math: every tensor op is a call into ``cinema_b200._C``.
torch is used for allocation (caching allocator), views and stream ownership only.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from cinema_b200 import _C
from cinema_b200.arena import ParamArena

BF16 = torch.bfloat16
F32 = torch.float32


# ------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------
@dataclass
class LinW:
    """One GEMM operand set: bf16 weight (N, K), fp32 bias (N) and their fp32 gradient targets."""

    w16: torch.Tensor
    bias: torch.Tensor | None
    gw: torch.Tensor | None
    gb: torch.Tensor | None

    @property
    def n(self) -> int:
        return self.w16.shape[0]

    @property
    def k(self) -> int:
        return self.w16.shape[1]


def linw(arena: ParamArena, weight: nn.Parameter, bias: nn.Parameter | None, train: bool) -> LinW:
    n = weight.shape[0]
    w16 = arena.w16(weight).view(n, -1)
    gw = arena.grad_view(weight).view(n, -1) if train and weight.requires_grad else None
    gb = arena.grad_view(bias) if train and bias is not None and bias.requires_grad else None
    return LinW(w16, bias.data if bias is not None else None, gw, gb)


def linw_fused(arena: ParamArena, weights: list[nn.Parameter], biases: list[nn.Parameter | None], train: bool) -> LinW | None:
    """Several Linear layers with the same input read as ONE (sum N, K) matrix, if the arena laid them out
    back to back; ``None`` otherwise (callers then issue one GEMM per layer)."""
    if any(b is None for b in biases) or not arena.adjacent(weights) or not arena.adjacent(biases):  # type: ignore[arg-type]
        return None
    k = weights[0][0].numel()
    n = sum(w.shape[0] for w in weights)
    tr_w = [w.requires_grad for w in weights]
    tr_b = [b.requires_grad for b in biases]  # type: ignore[union-attr]
    if len(set(tr_w)) != 1 or len(set(tr_b)) != 1:
        return None
    return LinW(
        arena.fused16(weights, (n, k)),
        arena.fused32(biases, (n,)),  # type: ignore[arg-type]
        arena.fused_grad(weights, (n, k)) if train and tr_w[0] else None,
        arena.fused_grad(biases, (n,)) if train and tr_b[0] else None,  # type: ignore[arg-type]
    )


# ------------------------------------------------------------------------------------------
# linear
# ------------------------------------------------------------------------------------------
def linear_fwd(x16: torch.Tensor, w: LinW, *, out_dtype: torch.dtype = BF16, residual: torch.Tensor | None = None,
               out: torch.Tensor | None = None, row_scale: torch.Tensor | None = None, rows_per_group: int = 0) -> torch.Tensor:
    """y = s * (x W^T + b) (+ residual).  x16 (M, K) bf16 -> (M, N) bf16 / fp32; ``row_scale`` s: one fp32 factor per
    ``rows_per_group`` rows (stochastic depth), default 1."""
    m = x16.shape[0]
    if out is None:
        out = torch.empty((m, w.n), dtype=out_dtype, device=x16.device)
    _C.gemm(x16, w.w16, out, bias=w.bias, residual=residual, row_scale=row_scale, rows_per_group=rows_per_group)
    return out


def linear_gelu_fwd(x16: torch.Tensor, w: LinW) -> tuple[torch.Tensor, torch.Tensor]:
    """(pre, act) = (x W^T + b, GELU_erf(pre)), both bf16; pre is kept for the backward."""
    m = x16.shape[0]
    pre = torch.empty((m, w.n), dtype=BF16, device=x16.device)
    act = torch.empty((m, w.n), dtype=BF16, device=x16.device)
    _C.gemm(x16, w.w16, pre, out2=act, bias=w.bias, epilogue=_C.EPI_GELU)
    return pre, act


def linear_bwd(dy16: torch.Tensor, x16: torch.Tensor | None, w: LinW, *, need_dx: bool = True,
               gelu_aux: torch.Tensor | None = None, dx_dtype: torch.dtype = BF16,
               dx_out: torch.Tensor | None = None, bias_done: bool = False,
               dx_colsum: torch.Tensor | None = None) -> torch.Tensor | None:
    """dW += dy^T x, db += colsum(dy), and dx = dy W (optionally * GELU'(gelu_aux), the input's pre-activation).
    ``bias_done``: the kernel that produced ``dy16`` already accumulated its column sums into ``w.gb``
    (``ln_bwd(..., dxsum=w.gb)`` or an upstream ``linear_bwd(..., dx_colsum=w.gb)``).
    ``dx_colsum``: fp32 (K,) buffer that receives the column sums of the bf16 dx written here -- the bias gradient of
    the Linear layer that produced this layer's input (fused into the dgrad GEMM epilogue)."""
    if w.gw is not None:
        _C.gemm(dy16, x16, w.gw, a_mn=True, b_mn=True, accumulate=True)
    if w.gb is not None and not bias_done:
        _C.colsum(dy16, w.gb)
    if not need_dx:
        return None
    if dx_out is None:
        dx_out = torch.empty((dy16.shape[0], w.k), dtype=dx_dtype, device=dy16.device)
    if dx_colsum is not None and dx_out.dtype != BF16:
        raise ValueError("dx_colsum needs a bf16 dx")
    if gelu_aux is not None:
        _C.gemm(dy16, w.w16, dx_out, b_mn=True, aux=gelu_aux, epilogue=_C.EPI_GELU_BWD, colsum=dx_colsum)
    else:
        _C.gemm(dy16, w.w16, dx_out, b_mn=True, colsum=dx_colsum)
    return dx_out


# ------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------
@dataclass
class NormW:
    gamma: torch.Tensor
    beta: torch.Tensor
    gg: torch.Tensor | None
    gb: torch.Tensor | None
    eps: float


def normw(arena: ParamArena, ln: nn.LayerNorm, train: bool) -> NormW:
    if ln.weight is None or ln.bias is None:
        raise NotImplementedError("LayerNorm without affine parameters is not supported by the B200 path")
    return NormW(ln.weight.data, ln.bias.data,
                 arena.grad_view(ln.weight) if train and ln.weight.requires_grad else None,
                 arena.grad_view(ln.bias) if train and ln.bias.requires_grad else None, float(ln.eps))


def ln_fwd(x32: torch.Tensor, w: NormW, *, want16: bool = True, want32: bool = False, stats: bool = True,
           y16: torch.Tensor | None = None, act: bool = False):
    m, d = x32.shape
    dev = x32.device
    if want16 and y16 is None:
        y16 = torch.empty((m, d), dtype=BF16, device=dev)
    y32 = torch.empty((m, d), dtype=F32, device=dev) if want32 else None
    mean = torch.empty(m, dtype=F32, device=dev) if stats else None
    rstd = torch.empty(m, dtype=F32, device=dev) if stats else None
    _C.layernorm_fwd(x32, w.gamma, w.beta, w.eps, y16=y16, y32=y32, mean=mean, rstd=rstd, act=act)
    return y16, y32, mean, rstd


def fusable_bias(w: "LinW | None", d: int) -> torch.Tensor | None:
    """Bias-gradient target of ``w`` if the LayerNorm backward can accumulate it (``dxsum``: TMA-staged path,
    short-row kernel for D in {16, 32, 64, 128}, TMA-staged path for 256 <= D <= 1024)."""
    return w.gb if (w is not None and w.gb is not None and (d in (16, 32, 64, 128) or (d % 8 == 0 and 256 <= d <= 1024))) else None


def ln_bwd(dy: torch.Tensor, x32: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, w: NormW, *,
           dres: torch.Tensor | None = None, want16: bool = True, dx32: torch.Tensor | None = None,
           beta_act: torch.Tensor | None = None, dxsum: torch.Tensor | None = None):
    """dx = LN'(dy) + dres -> (dx32, dx16).  ``dx32`` may alias ``dres`` (row-local read-then-write).
    ``dxsum`` (D,) fp32 += column sums of dx16: the bias gradient of the Linear whose output gradient dx16 is."""
    m, d = x32.shape
    if dx32 is None:
        dx32 = torch.empty((m, d), dtype=F32, device=x32.device)
    dx16 = torch.empty((m, d), dtype=BF16, device=x32.device) if want16 else None
    _C.layernorm_bwd(dy, x32, mean, rstd, w.gamma, dres=dres, dx32=dx32, dx16=dx16, dgamma=w.gg, dbeta=w.gb,
                     beta_act=beta_act, dxsum=dxsum)
    return dx32, dx16


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def check_head_dim(d: int) -> None:
    if d not in (32, 64):
        raise NotImplementedError(
            f"cinema_b200 attention kernels are built for head_dim 32 and 64 (ViT-B/L encoder, MAE decoder); got {d}")


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float):
    """q (B, Nq, H, d), k/v (B, Nk, H, d) strided bf16 views -> o (B, Nq, H, d) bf16 contiguous, lse (B, H, Nq)."""
    b, nq, h, d = q.shape
    check_head_dim(d)
    if k.shape[1] == 0:  # no keys (every token of every view masked away): softmax over nothing, the output is 0
        return (torch.zeros((b, nq, h, d), dtype=BF16, device=q.device),
                torch.full((b, h, nq), float("-inf"), dtype=F32, device=q.device))
    o = torch.empty((b, nq, h, d), dtype=BF16, device=q.device)
    lse = torch.empty((b, h, nq), dtype=F32, device=q.device)
    _C.attention_fwd(q, k, v, o, lse, scale)
    return o, lse


def attn_bwd(q, k, v, o, do, lse, dq, dk, dv, scale: float, dq_colsum=None, dk_colsum=None, dv_colsum=None) -> None:
    """``d?_colsum``: bias-gradient buffers of the q / k / v projections (fp32 (H * d,)), accumulated by the kernel."""
    b, nq, h, d = q.shape
    if k.shape[1] == 0:  # no keys: the output did not depend on q (and there is no k / v to differentiate)
        dq.zero_()
        return
    delta, dq_acc = _C.attention_bwd_workspace(b, h, nq, d, q.device)
    _C.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dq_acc, scale, dq_colsum, dk_colsum, dv_colsum)


# ------------------------------------------------------------------------------------------
# transformer block (cinema/vit.py:525-609): x += Attn(LN1(x)[, k]);  x += Mlp(LN2(x))
# ------------------------------------------------------------------------------------------
@dataclass
class BlockW:
    norm1: NormW
    norm2: NormW
    qkv: LinW | None  # fused [q | k | v] projection (self-attention)
    q: LinW
    kv: LinW
    proj: LinW
    fc1: LinW
    fc2: LinW
    n_heads: int
    scale: float
    drop_prob: float = 0.0  # stochastic depth of both residual branches (active in training mode only)
    drop_scale_by_keep: bool = True


def draw_drop_scales(b: int, drop_prob: float, scale_by_keep: bool, device: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-sample factors of the two DropPath layers of a block, drawn in the reference's order (attention branch, then
    MLP branch): Bernoulli(keep) / keep, timm ``drop_path`` semantics (cinema/vit.py:562,577)."""
    keep = 1.0 - drop_prob
    out = []
    for _ in range(2):
        t = torch.empty(b, dtype=F32, device=device).bernoulli_(keep)
        if keep > 0.0 and scale_by_keep:
            t.div_(keep)
        out.append(t)
    return out[0], out[1]


def out_bias_of(ws: "list[BlockW]", j: int, d: int) -> torch.Tensor | None:
    """``fc2.bias`` gradient buffer of block ``j`` if the kernel producing that block's output gradient may accumulate it
    (:func:`fusable_bias`); not with stochastic depth, where the branch sees a per-sample scaled gradient."""
    if j < 0 or ws[j].drop_prob > 0.0:
        return None
    return fusable_bias(ws[j].fc2, d)


def blockw(arena: ParamArena, blk: nn.Module, train: bool) -> BlockW:
    at = blk.attn
    if not isinstance(blk.norm1, nn.LayerNorm) or not isinstance(blk.norm2, nn.LayerNorm):
        raise NotImplementedError("the B200 block path supports nn.LayerNorm only")
    if not isinstance(at.q_norm, nn.Identity) or not isinstance(blk.ls1, nn.Identity):
        raise NotImplementedError("qk_norm / LayerScale are not part of the MAE hot path (cinema/vit.py:561-577)")
    dp = blk.drop_path1
    drop_prob = float(getattr(dp, "drop_prob", 0.0)) if blk.training and not isinstance(dp, nn.Identity) else 0.0
    return BlockW(
        normw(arena, blk.norm1, train), normw(arena, blk.norm2, train),
        linw_fused(arena, [at.q.weight, at.kv.weight], [at.q.bias, at.kv.bias], train),
        linw(arena, at.q.weight, at.q.bias, train), linw(arena, at.kv.weight, at.kv.bias, train),
        linw(arena, at.proj.weight, at.proj.bias, train),
        linw(arena, blk.mlp.fc1.weight, blk.mlp.fc1.bias, train), linw(arena, blk.mlp.fc2.weight, blk.mlp.fc2.bias, train),
        at.n_heads, float(at.scale), drop_prob, bool(getattr(dp, "scale_by_keep", True)),
    )


def _self_qkv_fwd(h16: torch.Tensor, w: BlockW) -> torch.Tensor:
    """[q | k | v] = h W_qkv^T: one GEMM when the arena fused the weights, else two into one buffer."""
    m, d = h16.shape
    qkv = torch.empty((m, 3 * d), dtype=BF16, device=h16.device)
    if w.qkv is not None:
        _C.gemm(h16, w.qkv.w16, qkv, bias=w.qkv.bias)
    else:
        _C.gemm(h16, w.q.w16, qkv[:, :d], bias=w.q.bias)
        _C.gemm(h16, w.kv.w16, qkv[:, d:], bias=w.kv.bias)
    return qkv


def block_fwd(x: torch.Tensor, w: BlockW, b: int, kv: tuple[torch.Tensor, torch.Tensor] | None, save: bool):
    """x: (B*N, D) fp32.  ``kv``: pre-projected (k, v) views (B, Nk, H, d) for cross-attention, else None.
    Returns (x_out fp32, saved-for-backward tuple or None)."""
    m, d = x.shape
    n = m // b
    hd = d // w.n_heads
    drop = draw_drop_scales(b, w.drop_prob, w.drop_scale_by_keep, x.device) if w.drop_prob > 0.0 else None
    s1, s2 = drop if drop is not None else (None, None)
    h1, _, mean1, rstd1 = ln_fwd(x, w.norm1, stats=save)
    if kv is None:
        qkv = _self_qkv_fwd(h1, w)
        q5 = qkv.view(b, n, 3, w.n_heads, hd)
        q, k, v = q5[:, :, 0], q5[:, :, 1], q5[:, :, 2]
        qsave = qkv
    else:
        q2 = linear_fwd(h1, w.q)
        q = q2.view(b, n, w.n_heads, hd)
        k, v = kv
        qsave = q2
    o, lse = attn_fwd(q, k, v, w.scale)
    o2 = o.view(m, d)
    x1 = linear_fwd(o2, w.proj, out_dtype=F32, residual=x, row_scale=s1, rows_per_group=n)
    h2, _, mean2, rstd2 = ln_fwd(x1, w.norm2, stats=save)
    pre, act = linear_gelu_fwd(h2, w.fc1)
    x2 = linear_fwd(act, w.fc2, out_dtype=F32, residual=x1, row_scale=s2, rows_per_group=n)
    if not save:
        return x2, None
    return x2, (x, mean1, rstd1, h1, qsave, o2, lse, x1, mean2, rstd2, h2, pre, act, drop)


def block_bwd(dx32: torch.Tensor, dx16: torch.Tensor, w: BlockW, b: int, saved,
              kv: tuple[torch.Tensor, torch.Tensor] | None, dkv: tuple[torch.Tensor, torch.Tensor] | None, *,
              fc2_bias_done: bool = False, out_bias: torch.Tensor | None = None):
    """Backward of :func:`block_fwd`.  (dx32, dx16) is the gradient of the block output in fp32 and bf16.
    For cross-attention ``dkv`` are the (dk, dv) views this block's key / value gradients are written to.
    ``fc2_bias_done``: the producer of ``dx16`` already accumulated ``fc2.bias``'s gradient; ``out_bias``: bias-gradient
    buffer of the Linear that consumes the returned dx16 as its output gradient (fused into the last LayerNorm backward).
    With stochastic depth (``saved[-1]`` = the per-sample factors of the forward) each branch sees s_b * dx: the bf16
    branch gradient is re-derived from dx32 by one scale-and-cast pass and the bias-gradient fusions across the residual
    are off (callers use :func:`out_bias_of`).
    Returns the (fp32, bf16) gradient of the block input; dx32 is updated in place."""
    x, mean1, rstd1, h1, qsave, o2, lse, x1, mean2, rstd2, h2, pre, act, drop = saved
    m, d = x.shape
    n = m // b
    hd = d // w.n_heads
    if drop is not None:
        if fc2_bias_done:
            raise ValueError("fc2 bias gradient cannot be pre-accumulated for a block with stochastic depth")
        dx16 = torch.empty((m, d), dtype=BF16, device=x.device)
        _C.scale_cast(dx32, dx16, drop[1], group=n * d)
    # ---- MLP path
    dpre = linear_bwd(dx16, act, w.fc2, gelu_aux=pre, bias_done=fc2_bias_done, dx_colsum=w.fc1.gb)
    dh2 = linear_bwd(dpre, h2, w.fc1, bias_done=w.fc1.gb is not None)
    proj_gb = fusable_bias(w.proj, d) if drop is None else None
    dx32, dx16 = ln_bwd(dh2, x1, mean2, rstd2, w.norm2, dres=dx32, dx32=dx32, dxsum=proj_gb)
    if drop is not None:
        _C.scale_cast(dx32, dx16, drop[0], group=n * d)
    # ---- attention path
    do2 = linear_bwd(dx16, o2, w.proj, bias_done=proj_gb is not None)
    do = do2.view(b, n, w.n_heads, hd)
    o = o2.view(b, n, w.n_heads, hd)
    if kv is None:
        q5 = qsave.view(b, n, 3, w.n_heads, hd)
        dqkv = torch.empty_like(qsave)
        d5 = dqkv.view(b, n, 3, w.n_heads, hd)
        # fused [q | k | v] bias gradient: the q part is summed inside the attention backward (its dQ conversion pass, free),
        # the k | v part by one column-sum pass over dqkv[:, d:].  The kernel can also sum dK / dV in its epilogue
        # (dk_colsum / dv_colsum), but at 685 x 685 x 12 heads that puts 590 k single-lane atomics on 48 cache lines:
        # 193 us against 152 + 12 us for the separate pass (tools/ab_attn_colsum.py, profiles/r02_attention.md)
        gb = w.qkv.gb if w.qkv is not None else None
        attn_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], o, do, lse, d5[:, :, 0], d5[:, :, 1], d5[:, :, 2], w.scale,
                 dq_colsum=gb[:d] if gb is not None else None)
        if gb is not None:
            _C.colsum(dqkv[:, d:], gb[d:])
        if w.qkv is not None:
            dh1 = linear_bwd(dqkv, h1, w.qkv, bias_done=gb is not None)
        else:
            # separate q / kv weights: two dgrads summed through the fp32 accumulate path
            tmp = torch.zeros((m, d), dtype=F32, device=x.device)
            for part, lw in ((dqkv[:, :d], w.q), (dqkv[:, d:], w.kv)):
                if lw.gw is not None:
                    _C.gemm(part, h1, lw.gw, a_mn=True, b_mn=True, accumulate=True)
                if lw.gb is not None:
                    _C.colsum(part, lw.gb)
                _C.gemm(part, lw.w16, tmp, b_mn=True, accumulate=True)
            dh1 = tmp
    else:
        q = qsave.view(b, n, w.n_heads, hd)
        dq2 = torch.empty_like(qsave)
        attn_bwd(q, kv[0], kv[1], o, do, lse, dq2.view(b, n, w.n_heads, hd), dkv[0], dkv[1], w.scale, dq_colsum=w.q.gb)
        dh1 = linear_bwd(dq2, h1, w.q, bias_done=w.q.gb is not None)
    dx32, dx16 = ln_bwd(dh1, x, mean1, rstd1, w.norm1, dres=dx32, dx32=dx32, dxsum=out_bias)
    return dx32, dx16


# ------------------------------------------------------------------------------------------
# per-view concurrency
# ------------------------------------------------------------------------------------------
_SIDE_STREAMS: dict[tuple, list] = {}
VIEW_STREAMS_ENABLED = True  # bench.py turns this off for its per-kernel instrumented step (serial launches)


class ViewStreams:
    """Fork / join of per-view work onto side streams.

    The views of a frame-set (SAX + 3 LAX) are independent until their tokens are concatenated, and the LAX kernels
    are far too small to fill 148 SMs (a 12 x 12 token grid per sample), so view i > 0 runs on side stream i - 1
    while view 0 stays on the caller's stream.  Every fork starts with ``side.wait_stream(main)`` and ``join()``
    makes the caller's stream wait for the side streams: memory handed between streams is always ordered, and the
    pattern is capturable in a CUDA graph (the side streams join the capture through the fork event)."""

    def __init__(self, device: torch.device, enabled: bool = True) -> None:
        self.enabled = enabled and VIEW_STREAMS_ENABLED and device.type == "cuda"
        self.used: list = []
        if self.enabled:
            self.main = torch.cuda.current_stream(device)
            key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(3)]
            self.pool = _SIDE_STREAMS[key]

    def view(self, i: int):
        """Context manager for the work of view ``i``."""
        import contextlib

        if not self.enabled or i == 0:
            return contextlib.nullcontext()
        st = self.pool[(i - 1) % len(self.pool)]
        st.wait_stream(self.main)
        if st not in self.used:
            self.used.append(st)
        return torch.cuda.stream(st)

    def join(self) -> None:
        if self.enabled:
            for st in self.used:
                self.main.wait_stream(st)
            self.used = []


# ------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------
_ARANGE_CACHE: dict[tuple, torch.Tensor] = {}


def arange_idx(b: int, start: int, k: int, device: torch.device) -> torch.Tensor:
    """(B, k) int32 index rows [start, start + k): the 'slice' form of the row gather / scatter kernels."""
    key = (b, start, k, str(device))
    t = _ARANGE_CACHE.get(key)
    if t is None:
        t = torch.arange(start, start + k, dtype=torch.int32, device=device).repeat(b, 1).contiguous()
        if len(_ARANGE_CACHE) > 256:
            _ARANGE_CACHE.clear()
        _ARANGE_CACHE[key] = t
    return t
