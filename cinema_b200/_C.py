"""ctypes binding of libcinema_b200.so -- the only door from Python to the CUDA kernels.

Every wrapper takes torch CUDA tensors, checks dtype / alignment, passes raw device pointers,
sizes and the current CUDA stream across the C-ABI (include/cinema_b200.h) and raises
``RuntimeError`` with the library's message on a non-zero status.  There is NO fallback: if the
shared library is missing or a tensor lives on the CPU the call fails loudly.
"""

from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libcinema_b200.so"

DT_BF16, DT_F32 = 0, 1
EPI_NONE, EPI_GELU, EPI_GELU_BWD = 0, 1, 2

_lib = None
launches = 0  # number of kernel-launching C-ABI calls made by this process (bench.py reports it)


class KernelLibraryMissing(RuntimeError):
    pass


_SIGNATURES = {
    "cb_version": (c_int, []),
    "cb_last_error": (c_char_p, []),
    "cb_device_info": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cb_set_pdl": (c_int, [c_int]),
    "cb_attention_trace": (c_int, [c_void_p]),
    "cb_gemm_trace": (c_int, [c_void_p]),
    "cb_scale_intensity": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p]),
    "cb_zoom_intensity": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                  c_void_p]),
    "cb_seg_loss_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_seg_loss_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_gemm_bf16": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_void_p,
                             c_longlong, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_longlong, c_void_p,
                             c_longlong, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "cb_conv_gemm_bf16": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p,
                                  c_longlong, c_int, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "cb_colsum_bf16": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "cb_attention_fwd": (c_int, [c_void_p, c_longlong, c_longlong, c_longlong] * 4
                         + [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "cb_attention_bwd": (c_int, [c_void_p, c_longlong, c_longlong, c_longlong] * 5 + [c_void_p]
                         + [c_void_p, c_longlong, c_longlong, c_longlong] * 3
                         + [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                            c_void_p]),
    "cb_layernorm_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_longlong,
                                 c_void_p, c_longlong, c_void_p, c_void_p, c_int, c_void_p]),
    "cb_layernorm_bwd": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_longlong, c_int, c_int, c_void_p, c_longlong, c_void_p, c_longlong,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_expand_token_index": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_dwconv_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "cb_dwconv_tokens_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cb_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "cb_mask_to_index": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_gather_rows": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_void_p, c_longlong, c_longlong,
                               c_longlong, c_void_p]),
    "cb_scatter_rows": (c_int, [c_void_p, c_longlong, c_longlong, c_void_p, c_int, c_int, c_void_p, c_longlong,
                                c_longlong, c_void_p]),
    "cb_embed_rows_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_longlong, c_longlong, c_void_p]),
    "cb_colsum_seg_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cb_scale_cast_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_float, c_longlong, c_void_p]),
    "cb_sumsq_f32": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p]),
    "cb_adamw_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_float, c_float,
                              c_float, c_float, c_void_p, c_float, c_float, c_void_p]),
    "cb_mae_loss_finalize": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_patchify": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cb_gather_patches": (c_int, [c_void_p, c_int, c_longlong, c_longlong, c_void_p, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "cb_scatter_patches": (c_int, [c_void_p, c_int, c_void_p, c_int, c_longlong, c_longlong, c_void_p, c_int, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cb_rope_apply": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p]),
    "cb_masked_mse_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise KernelLibraryMissing(
                f"{LIB_PATH} not found. cinema_b200 has no CPU / PyTorch fallback: build the CUDA library first "
                f"(python -m cinema_b200.build, or __graft_entry__.build())."
            )
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(rc: int, what: str) -> None:
    global launches
    if rc != 0:
        raise RuntimeError(f"cinema_b200 {what} failed (status {rc}): {lib().cb_last_error().decode()}")
    launches += 1


def _stream() -> int:
    """Stream of the CURRENT device.  The library keeps per-process state for the current device (SM count, opt-in
    shared-memory attributes), so kernels must be launched with the tensors' device current: `_ptr` checks that."""
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("cinema_b200 kernels need CUDA tensors (no CPU fallback); got a CPU tensor")
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"cinema_b200 launches on the current device (cuda:{torch.cuda.current_device()}) but got a tensor on "
                           f"{t.device}: call torch.cuda.set_device({t.device.index}) first (one process per GPU)")
    return t.data_ptr()


def _ints(vals) -> ctypes.Array:
    return (c_int * len(vals))(*[int(v) for v in vals])


def _lls(vals) -> ctypes.Array:
    return (c_longlong * len(vals))(*[int(v) for v in vals])


def _row_major_2d(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name} must be a 2-D tensor with unit inner stride, got {tuple(t.shape)} {t.stride()}")
    return t.stride(0)


def set_pdl(enabled: bool) -> bool:
    """Programmatic dependent launch for all kernels (default on); returns the previous setting."""
    return bool(lib().cb_set_pdl(int(enabled)))


def device_info() -> tuple[int, int, int]:
    a, b, c = c_int(), c_int(), c_int()
    _check(lib().cb_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "device_info")
    return a.value, b.value, c.value


# --------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None, *, a_mn: bool = False, b_mn: bool = False,
         accumulate: bool = False, out2: torch.Tensor | None = None, bias: torch.Tensor | None = None,
         residual: torch.Tensor | None = None, aux: torch.Tensor | None = None, epilogue: int = EPI_NONE,
         alpha: float = 1.0, split_k: int = 0, block_n: int = 0, colsum: torch.Tensor | None = None,
         row_scale: torch.Tensor | None = None, rows_per_group: int = 0) -> None:
    """C[M,N] = alpha * A . B^T with the fused epilogue of cb_gemm_bf16 (see include/cinema_b200.h).
    ``colsum`` (fp32 (N,), optional, bf16 outputs): += column sums of the stored bf16 values.
    ``row_scale`` (fp32, one factor per ``rows_per_group`` rows): scales alpha * acc + bias before the residual."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    lda, ldb = _row_major_2d(a, "A"), _row_major_2d(b, "B")
    m, k = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    n, kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if k != kb:
        raise RuntimeError(f"gemm: contraction mismatch {k} vs {kb}")
    ref = out if out is not None else out2
    if tuple(ref.shape) != (m, n):
        raise RuntimeError(f"gemm: output shape {tuple(ref.shape)} != ({m}, {n})")
    out_dt = DT_BF16
    ldo = 0
    if out is not None:
        ldo = _row_major_2d(out, "out")
        out_dt = DT_F32 if out.dtype == torch.float32 else DT_BF16
        assert out.dtype in (torch.float32, torch.bfloat16)
    ldo2 = _row_major_2d(out2, "out2") if out2 is not None else 0
    ldr = _row_major_2d(residual, "residual") if residual is not None else 0
    ldaux = _row_major_2d(aux, "aux") if aux is not None else 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n and bias.is_contiguous()
    if residual is not None:
        assert residual.dtype == torch.float32 and tuple(residual.shape) == (m, n)
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and tuple(aux.shape) == (m, n)
    if out2 is not None:
        assert out2.dtype == torch.bfloat16
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.numel() == n and colsum.is_contiguous()
    if row_scale is not None:
        assert row_scale.dtype == torch.float32 and row_scale.is_contiguous() and rows_per_group > 0
        assert row_scale.numel() * rows_per_group >= m
    _check(lib().cb_gemm_bf16(_ptr(a), lda, int(a_mn), _ptr(b), ldb, int(b_mn), m, n, k, _ptr(out), ldo, out_dt,
                              int(accumulate), _ptr(out2), ldo2, _ptr(bias), _ptr(residual), ldr, _ptr(aux), ldaux,
                              epilogue, float(alpha), split_k, block_n, _ptr(colsum), _ptr(row_scale), int(rows_per_group),
                              _stream()), "gemm")


def conv_gemm(x_rows: torch.Tensor, w_taps: torch.Tensor, out: torch.Tensor, row_off, *, bias: torch.Tensor | None = None,
              residual: torch.Tensor | None = None, block_n: int = 0) -> None:
    """EXPERIMENTAL (see include/cinema_b200.h): out[r] = sum_t x_rows[r + row_off[t]] @ w_taps[:, t*C:(t+1)*C]^T (+ bias)
    (+ residual).  x_rows (R, C_in) bf16 zero-haloed row space, w_taps (C_out, taps * C_in) bf16, out (R, C_out) bf16 / fp32."""
    assert x_rows.dtype == torch.bfloat16 and w_taps.dtype == torch.bfloat16
    ldx, ldw, ldo = _row_major_2d(x_rows, "X"), _row_major_2d(w_taps, "W"), _row_major_2d(out, "out")
    rows, c_in = x_rows.shape
    c_out = w_taps.shape[0]
    n_taps = len(row_off)
    assert w_taps.shape[1] == n_taps * c_in and tuple(out.shape) == (rows, c_out)
    out_dt = DT_BF16 if out.dtype == torch.bfloat16 else DT_F32
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.float32 and tuple(residual.shape) == (rows, c_out)
        ldr = _row_major_2d(residual, "residual")
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == c_out
    _check(lib().cb_conv_gemm_bf16(_ptr(x_rows), ldx, rows, c_in, _ptr(w_taps), ldw, c_out, n_taps, _ints(row_off), _ptr(out), ldo,
                                   out_dt, _ptr(bias), _ptr(residual), ldr, block_n, _stream()),
           f"conv_gemm (x {tuple(x_rows.shape)}, w {tuple(w_taps.shape)}, out {tuple(out.shape)})")


def colsum(x: torch.Tensor, out: torch.Tensor) -> None:
    """out[N] (fp32) += column sums of bf16 x[M,N]."""
    assert x.dtype == torch.bfloat16 and out.dtype == torch.float32 and out.numel() == x.shape[1]
    _check(lib().cb_colsum_bf16(_ptr(x), _row_major_2d(x, "x"), x.shape[0], x.shape[1], _ptr(out), _stream()), "colsum")


def _bnh(t: torch.Tensor, name: str) -> tuple[int, int, int, int]:
    """(ptr, stride_b, stride_n, stride_h) of a (B, N, H, d) bf16 view with unit stride on d."""
    if t.dim() != 4 or t.stride(3) != 1 or t.dtype != torch.bfloat16:
        raise RuntimeError(f"{name} must be a bf16 (B, N, H, d) view with unit inner stride")
    return _ptr(t), t.stride(0), t.stride(1), t.stride(2)


def attention_fwd(q, k, v, o, lse, scale: float) -> None:
    b, nq, h, d = q.shape
    nk = k.shape[1]
    assert lse.dtype == torch.float32 and lse.numel() == b * h * nq
    _check(lib().cb_attention_fwd(*_bnh(q, "q"), *_bnh(k, "k"), *_bnh(v, "v"), *_bnh(o, "o"), _ptr(lse), b, h, nq, nk,
                                  d, float(scale), _stream()), "attention_fwd")


def attention_bwd_workspace(b: int, h: int, nq: int, d: int, device) -> tuple[torch.Tensor, torch.Tensor]:
    """(delta, dq_acc) scratch of cb_attention_bwd: delta holds two (B*H, NqP) fp32 vectors (NqP = Nq rounded up to
    128), dq_acc the fp32 dQ accumulator (B, H, Nq, d)."""
    nqp = (nq + 127) // 128 * 128
    return (torch.empty((2, b * h, nqp), dtype=torch.float32, device=device),
            torch.empty((b, h, nq, d), dtype=torch.float32, device=device))


def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dq_acc, scale: float, dq_colsum=None, dk_colsum=None,
                  dv_colsum=None) -> None:
    """``d?_colsum`` (fp32 (H * d,), optional): += column sums of the stored bf16 dQ / dK / dV (projection bias gradients)."""
    b, nq, h, d = q.shape
    nk = k.shape[1]
    for cs in (dq_colsum, dk_colsum, dv_colsum):
        assert cs is None or (cs.dtype == torch.float32 and cs.numel() == h * d and cs.is_contiguous())
    assert delta.dtype == torch.float32 and delta.numel() >= 2 * b * h * ((nq + 127) // 128 * 128), "delta workspace too small"
    assert dq_acc.dtype == torch.float32 and dq_acc.numel() >= b * h * nq * d
    _check(lib().cb_attention_bwd(*_bnh(q, "q"), *_bnh(k, "k"), *_bnh(v, "v"), *_bnh(o, "o"), *_bnh(do, "do"),
                                  _ptr(lse), *_bnh(dq, "dq"), *_bnh(dk, "dk"), *_bnh(dv, "dv"), _ptr(delta),
                                  _ptr(dq_acc), b, h, nq, nk, d, float(scale), _ptr(dq_colsum), _ptr(dk_colsum),
                                  _ptr(dv_colsum), _stream()), "attention_bwd")


def layernorm_fwd(x, gamma, beta, eps, y16=None, y32=None, mean=None, rstd=None, act: bool = False) -> None:
    m, d = x.shape
    assert x.dtype == torch.float32 and gamma.dtype == torch.float32 and beta.dtype == torch.float32
    _check(lib().cb_layernorm_fwd(_ptr(x), _row_major_2d(x, "x"), _ptr(gamma), _ptr(beta), m, d, float(eps),
                                  _ptr(y16), _row_major_2d(y16, "y16") if y16 is not None else 0,
                                  _ptr(y32), _row_major_2d(y32, "y32") if y32 is not None else 0,
                                  _ptr(mean), _ptr(rstd), int(act), _stream()), "layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, dres=None, dx32=None, dx16=None, dgamma=None, dbeta=None,
                  beta_act=None, dxsum=None) -> None:
    """dxsum (fp32 (D), optional) += column sums of the bf16-rounded dx: the bias gradient of the Linear fed by dx16."""
    m, d = x.shape
    dt = DT_BF16 if dy.dtype == torch.bfloat16 else DT_F32
    assert dy.dtype in (torch.bfloat16, torch.float32) and x.dtype == torch.float32
    _check(lib().cb_layernorm_bwd(_ptr(dy), _row_major_2d(dy, "dy"), dt, _ptr(x), _row_major_2d(x, "x"), _ptr(mean),
                                  _ptr(rstd), _ptr(gamma), _ptr(dres),
                                  _row_major_2d(dres, "dres") if dres is not None else 0, m, d,
                                  _ptr(dx32), _row_major_2d(dx32, "dx32") if dx32 is not None else 0,
                                  _ptr(dx16), _row_major_2d(dx16, "dx16") if dx16 is not None else 0,
                                  _ptr(dgamma), _ptr(dbeta), _ptr(dxsum), _ptr(beta_act), _stream()), "layernorm_bwd")


def cast_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    _check(lib().cb_cast_f32_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "cast")


def mask_to_index(mask: torch.Tensor, n_keep: int):
    """(B,n) bool mask -> keep_idx (B,n_keep), drop_idx (B,n-n_keep), slot (B,n); all int32, ascending."""
    assert mask.dtype == torch.bool and mask.dim() == 2 and mask.is_contiguous()
    b, n = mask.shape
    keep = torch.empty((b, n_keep), dtype=torch.int32, device=mask.device)
    drop = torch.empty((b, n - n_keep), dtype=torch.int32, device=mask.device)
    slot = torch.empty((b, n), dtype=torch.int32, device=mask.device)
    _check(lib().cb_mask_to_index(_ptr(mask), b, n, n_keep, _ptr(keep), _ptr(drop), _ptr(slot), _stream()),
           "mask_to_index")
    return keep, drop, slot


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: torch.Tensor, out_off: int = 0) -> None:
    """out[b, out_off+i] = src[b, idx[b,i]] (src (B,n,D) or (n,D)/(1,n,D) broadcast); bit-exact."""
    b, k = idx.shape
    assert idx.dtype == torch.int32 and idx.is_contiguous() and src.is_contiguous() and out.is_contiguous()
    row_bytes = src.shape[-1] * src.element_size()
    assert out.dtype == src.dtype and out.shape[-1] == src.shape[-1] and out.dim() == 3
    if src.dim() == 2 or (src.shape[0] == 1 and b > 1):
        bstride = 0  # one table broadcast over the batch
    else:
        assert src.dim() == 3 and src.shape[0] == b
        bstride = src.shape[1]
    _check(lib().cb_gather_rows(_ptr(src), bstride, _ptr(idx), b, k, _ptr(out), out.shape[1], out_off, row_bytes,
                                _stream()), "gather_rows")


def scatter_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor, src_off: int = 0) -> None:
    """dst[b, idx[b,i]] = src[b, src_off+i]; rows of dst not listed stay untouched."""
    b, k = idx.shape
    assert idx.dtype == torch.int32 and idx.is_contiguous() and src.is_contiguous() and dst.is_contiguous()
    assert src.dim() == 3 and dst.dim() == 3 and src.dtype == dst.dtype
    row_bytes = src.shape[-1] * src.element_size()
    _check(lib().cb_scatter_rows(_ptr(src), src.shape[1], src_off, _ptr(idx), b, k, _ptr(dst), dst.shape[1], row_bytes,
                                 _stream()), "scatter_rows")


def embed_rows(a, a_off: int, row, table, idx, b: int, k: int, out=None, out16=None, out_off: int = 0) -> None:
    """out[b, out_off+i] = a[b, a_off+i] (opt) + row (opt) + table[idx[b,i]] (opt); fp32 math, stored to
    ``out`` (fp32 (B, n, D)) and / or ``out16`` (bf16, same shape)."""
    ref = out if out is not None else out16
    d = ref.shape[-1]
    assert ref.dim() == 3 and ref.is_contiguous() and ref.shape[0] == b
    if out is not None:
        assert out.dtype == torch.float32
    if out16 is not None:
        assert out16.dtype == torch.bfloat16 and out16.is_contiguous()
        assert out is None or out16.shape == out.shape
    if table is not None:
        assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[-1] == d
        assert idx.dtype == torch.int32 and idx.is_contiguous() and tuple(idx.shape) == (b, k)
    if a is not None:
        assert a.dtype == torch.float32 and a.dim() == 3 and a.is_contiguous() and a.shape[-1] == d
    if row is not None:
        assert row.dtype == torch.float32 and row.numel() == d and row.is_contiguous()
    _check(lib().cb_embed_rows_f32(_ptr(a), a.shape[1] if a is not None else 0, a_off, _ptr(row), _ptr(table),
                                   _ptr(idx) if table is not None else None, b, k, d, _ptr(out), _ptr(out16),
                                   ref.shape[1], out_off, _stream()), "embed_rows")


def colsum_seg(x: torch.Tensor, off: int, k: int, out: torch.Tensor) -> None:
    """out[D] (fp32) += sum_{b, i<k} x[b, off+i, :] for fp32 x (B, n, D)."""
    assert x.dtype == torch.float32 and x.dim() == 3 and x.is_contiguous() and out.dtype == torch.float32
    assert out.numel() == x.shape[-1] and out.is_contiguous() and off + k <= x.shape[1]
    _check(lib().cb_colsum_seg_f32(_ptr(x), x.shape[1], off, x.shape[0], k, x.shape[-1], _ptr(out), _stream()),
           "colsum_seg")


def scale_cast(src: torch.Tensor, dst: torch.Tensor, scale_dev: torch.Tensor | None = None, scale: float = 1.0,
               group: int = 0) -> None:
    """dst (bf16) = src (fp32) * scale * scale_dev[0], or * scale_dev[i // group] per element i when ``group`` > 0."""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    if scale_dev is not None:
        assert scale_dev.dtype == torch.float32 and scale_dev.is_contiguous()
    if group:
        assert scale_dev is not None and scale_dev.numel() * group == src.numel()
    _check(lib().cb_scale_cast_bf16(_ptr(src), _ptr(dst), src.numel(), _ptr(scale_dev), float(scale), int(group),
                                    _stream()), "scale_cast")


_RAW_DT = {torch.uint8: 2, torch.int16: 3, torch.uint16: 4, torch.float32: 1}


def scale_intensity(raw: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (B, ...) fp32 = (raw - lo[b]) / (hi[b] - lo[b]) per sample (0 for a constant sample): MONAI ScaleIntensity(0, 1)
    on the device in one pass.  raw: uint8 / int16 / uint16 / float32, contiguous; lo / hi: fp32 (B,)."""
    assert raw.dtype in _RAW_DT and raw.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
    assert out.shape == raw.shape and lo.dtype == torch.float32 and hi.dtype == torch.float32
    b = raw.shape[0]
    assert lo.numel() == b and hi.numel() == b and lo.is_contiguous() and hi.is_contiguous()
    _check(lib().cb_scale_intensity(_ptr(raw), _RAW_DT[raw.dtype], _ptr(lo), _ptr(hi), _ptr(out), b, raw.numel() // b,
                                    _stream()), "scale_intensity")
    return out


def zoom_intensity(raw: torch.Tensor, extent: torch.Tensor, zoom: torch.Tensor, out: torch.Tensor,
                   keys_ws: torch.Tensor | None = None) -> torch.Tensor:
    """RandZoom(keep_size) -> ScaleIntensity -> SpatialPad("end") of a raw-dtype batch on the device (cb_zoom_intensity).
    raw (B, 1, *size): frames stored in the model's input size, each occupying the corner [0, extent[b]) of it; extent
    (B, 3) int32 (unused axes 1); zoom (B,) fp32, 1 = identity; 3 spatial axes -> trilinear, 2 -> bicubic.  Returns
    ``out`` (B, 1, *size) fp32 in [0, 1], zero outside the frame."""
    assert raw.dtype in _RAW_DT and raw.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
    assert out.shape == raw.shape and raw.dim() in (4, 5) and raw.shape[1] == 1
    b = raw.shape[0]
    size = tuple(raw.shape[2:])
    assert extent.dtype == torch.int32 and extent.is_contiguous() and tuple(extent.shape) == (b, 3)
    assert zoom.dtype == torch.float32 and zoom.is_contiguous() and zoom.numel() == b
    if keys_ws is None:
        keys_ws = torch.empty(2 * b, dtype=torch.int32, device=raw.device)
    assert keys_ws.numel() >= 2 * b and keys_ws.element_size() == 4 and keys_ws.is_contiguous()
    sz = (c_int * len(size))(*size)
    _check(lib().cb_zoom_intensity(_ptr(raw), _RAW_DT[raw.dtype], _ptr(extent), _ptr(zoom), _ptr(out), _ptr(keys_ws), b,
                                   len(size), sz, int(len(size) == 2), _stream()), "zoom_intensity")
    return out


_LABEL_DT = {torch.int64: 0, torch.int32: 1, torch.int16: 2, torch.uint8: 3}


def _seg_args(logits: torch.Tensor, labels: torch.Tensor):
    assert logits.dtype in (torch.float32, torch.bfloat16) and logits.is_contiguous() and logits.dim() >= 3
    b, c = logits.shape[0], logits.shape[1]
    s = logits[0, 0].numel()
    assert labels.dtype in _LABEL_DT and labels.is_contiguous() and labels.numel() == b * s, "labels: (B, 1, *spatial) integers"
    return b, c, s, (DT_F32 if logits.dtype == torch.float32 else DT_BF16), _LABEL_DT[labels.dtype]


def seg_loss_fwd(logits: torch.Tensor, labels: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Cross-entropy (ignore_index -1) + foreground soft-Dice loss of channel-first logits (B, C, *spatial) against integer
    labels (B, 1, *spatial): -> (out (3,) fp32 = [loss, cross-entropy, mean Dice loss], coef for :func:`seg_loss_bwd`)."""
    b, c, s, ldt, ydt = _seg_args(logits, labels)
    acc = torch.empty(b * c * 3 + 2, dtype=torch.float32, device=logits.device)
    out = torch.empty(3, dtype=torch.float32, device=logits.device)
    coef = torch.empty(b * c * 2 + 1, dtype=torch.float32, device=logits.device)
    _check(lib().cb_seg_loss_fwd(_ptr(logits), ldt, _ptr(labels), ydt, b, c, s, _ptr(acc), _ptr(out), _ptr(coef), _stream()),
           "seg_loss_fwd")
    return out, coef


def seg_loss_bwd(logits: torch.Tensor, labels: torch.Tensor, coef: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    """d loss / d logits * grad_out (a one-element fp32 device tensor), in the logits' dtype and layout."""
    b, c, s, ldt, ydt = _seg_args(logits, labels)
    assert grad_out.dtype == torch.float32 and grad_out.numel() == 1 and coef.numel() == b * c * 2 + 1
    dl = torch.empty_like(logits)
    _check(lib().cb_seg_loss_bwd(_ptr(logits), ldt, _ptr(labels), ydt, b, c, s, _ptr(coef), _ptr(grad_out), _ptr(dl), _stream()),
           "seg_loss_bwd")
    return dl


def mae_loss_finalize(acc: torch.Tensor, sq_count, patch_count, out: torch.Tensor, scales: torch.Tensor) -> None:
    """acc (V, 8) -> out[0] = loss, out[1+5v:6+5v] = per-view metrics, scales[v] = d loss / d diff factor."""
    v = acc.shape[0]
    assert acc.dtype == torch.float32 and acc.is_contiguous() and acc.shape[1] == 8
    assert out.dtype == torch.float32 and out.numel() >= 1 + 5 * v and scales.dtype == torch.float32
    sq = (c_float * v)(*[float(x) for x in sq_count])
    pc = (c_float * v)(*[float(x) for x in patch_count])
    _check(lib().cb_mae_loss_finalize(_ptr(acc), v, sq, pc, _ptr(out), _ptr(scales), _stream()), "mae_loss_finalize")


def patchify(src: torch.Tensor, dst: torch.Tensor, b: int, c: int, spatial, patch, inverse: bool) -> None:
    assert src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype
    _check(lib().cb_patchify(_ptr(src), _ptr(dst), b, c, len(spatial), _ints(spatial), _ints(patch),
                             src.element_size(), int(inverse), _stream()), "patchify")


def _src_strides(x: torch.Tensor):
    return x.stride(0), x.stride(1), [x.stride(i) for i in range(2, x.dim())]


def gather_patches(src: torch.Tensor, grid, patch, idx: torch.Tensor | None, chan_last: bool, out: torch.Tensor) -> None:
    """out rows (bf16) = patches of tokens idx[b,i] of a strided (B,C,*spatial) fp32/bf16 source."""
    assert src.dtype in (torch.float32, torch.bfloat16) and out.dtype == torch.bfloat16 and out.is_contiguous()
    sb, sc, ss = _src_strides(src)
    k = idx.shape[1] if idx is not None else 0
    if idx is not None:
        assert idx.dtype == torch.int32 and idx.is_contiguous()
    _check(lib().cb_gather_patches(_ptr(src), DT_F32 if src.dtype == torch.float32 else DT_BF16, sb, sc, _lls(ss),
                                   src.shape[0], src.shape[1], len(grid), _ints(grid), _ints(patch), _ptr(idx), k,
                                   int(chan_last), _ptr(out), _stream()), "gather_patches")


def scatter_patches(rows: torch.Tensor, dst: torch.Tensor, grid, patch, idx: torch.Tensor | None, chan_last: bool,
                    accumulate: bool = False) -> None:
    """inverse of gather_patches into a strided (B,C,*spatial) gradient buffer (overwrite or accumulate)."""
    assert rows.is_contiguous() and rows.dtype in (torch.float32, torch.bfloat16)
    assert dst.dtype in (torch.float32, torch.bfloat16)
    sb, sc, ss = _src_strides(dst)
    k = idx.shape[1] if idx is not None else 0
    dt = lambda t: DT_F32 if t.dtype == torch.float32 else DT_BF16  # noqa: E731
    _check(lib().cb_scatter_patches(_ptr(rows), dt(rows), _ptr(dst), dt(dst), sb, sc, _lls(ss), dst.shape[0],
                                    dst.shape[1], len(grid), _ints(grid), _ints(patch), _ptr(idx), k, int(chan_last),
                                    int(accumulate), _stream()), "scatter_patches")


def rope_apply(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    """x (B, N, H, d) fp32 / bf16 contiguous, cos / sin (>= N, ro/2) fp32 -> rotated copy."""
    assert x.dim() == 4 and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    assert cos.dtype == torch.float32 and sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
    b, n, h, d = x.shape
    assert cos.shape[0] >= n and cos.shape == sin.shape
    y = torch.empty_like(x)
    _check(lib().cb_rope_apply(_ptr(x), _ptr(y), DT_F32 if x.dtype == torch.float32 else DT_BF16, _ptr(cos), _ptr(sin),
                               b, n, h, d, 2 * cos.shape[1], int(transpose), _stream()), "rope_apply")
    return y


def masked_mse_fwd(image, patch, mask, slot, pred, norm_target: bool, eps: float, acc, diff) -> None:
    assert image.dtype == torch.float32 and image.is_contiguous() and pred.dtype == torch.float32
    assert pred.is_contiguous() and mask.dtype == torch.bool and mask.is_contiguous() and slot.dtype == torch.int32
    assert acc.dtype == torch.float32 and acc.numel() >= 8
    b, c, *spatial = image.shape
    _check(lib().cb_masked_mse_fwd(_ptr(image), b, c, len(spatial), _ints(spatial), _ints(patch), _ptr(mask),
                                   _ptr(slot), _ptr(pred), pred.shape[1], int(norm_target), float(eps), _ptr(acc),
                                   _ptr(diff), _stream()), "masked_mse_fwd")


def sumsq(x: torch.Tensor, out: torch.Tensor) -> None:
    """out[0] += sum(x^2) over a flat fp32 buffer."""
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.float32
    _check(lib().cb_sumsq_f32(_ptr(x), x.numel(), _ptr(out), _stream()), "sumsq")


def adamw_flat(p, g, m, v, p16, hyper, beta1: float, beta2: float, eps: float, weight_decay: float, gnorm_sq,
               max_norm: float, grad_scale: float) -> None:
    """AdamW + clipping + bf16 shadow refresh on a flat fp32 segment; hyper = device [lr, 1-b1^t, 1-b2^t]."""
    n = p.numel()
    for t in (p, g, m, v):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    assert p16 is None or (p16.dtype == torch.bfloat16 and p16.numel() == n)
    assert hyper.dtype == torch.float32 and hyper.numel() >= 3
    _check(lib().cb_adamw_flat(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p16), n, _ptr(hyper), float(beta1), float(beta2),
                               float(eps), float(weight_decay), _ptr(gnorm_sq), float(max_norm), float(grad_scale),
                               _stream()), "adamw_flat")


def expand_token_index(keep: torch.Tensor, grid_tok, f) -> torch.Tensor:
    """(B, nk) visible token ids -> (B, nk * prod(f)) position ids in the level grid grid_tok * f."""
    assert keep.dtype == torch.int32 and keep.is_contiguous()
    b, nk = keep.shape
    p = 1
    for x in f:
        p *= int(x)
    out = torch.empty((b, nk * p), dtype=torch.int32, device=keep.device)
    _check(lib().cb_expand_token_index(_ptr(keep), b, nk, len(f), _ints(grid_tok), _ints(f), _ptr(out), _stream()),
           "expand_token_index")
    return out


def dwconv_tokens(x, out, w, bias, mask, slot, keep, grid_tok, f, transpose: bool = False) -> None:
    """Depth-wise 5^n conv over token-major channel-last rows x (B*nk*P, C) bf16 -> out (same shape)."""
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert x.is_contiguous() and out.is_contiguous() and w.is_contiguous() and x.shape == out.shape
    assert mask.dtype == torch.bool and mask.is_contiguous() and slot.dtype == torch.int32 and keep.dtype == torch.int32
    b, nk = keep.shape
    c = x.shape[-1]
    assert w.shape[0] == c and (bias is None or (bias.dtype == torch.float32 and bias.numel() == c))
    _check(lib().cb_dwconv_tokens(_ptr(x), _ptr(out), _ptr(w), _ptr(bias), _ptr(mask), _ptr(slot), _ptr(keep), b, nk, c,
                                  len(f), _ints(grid_tok), _ints(f), int(transpose), _stream()), "dwconv_tokens")


def dwconv_tokens_wgrad(x, dy, dw, db, mask, slot, keep, grid_tok, f) -> None:
    """dw (C, 1, 5, 5[, 5]) fp32 += correlation(dy, x); db (C) fp32 += colsum(dy)."""
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.shape == dy.shape
    assert x.is_contiguous() and dy.is_contiguous() and dw.dtype == torch.float32 and dw.is_contiguous()
    b, nk = keep.shape
    c = x.shape[-1]
    _check(lib().cb_dwconv_tokens_wgrad(_ptr(x), _ptr(dy), _ptr(dw), _ptr(db), _ptr(mask), _ptr(slot), _ptr(keep), b, nk, c,
                                        len(f), _ints(grid_tok), _ints(f), _stream()), "dwconv_tokens_wgrad")
