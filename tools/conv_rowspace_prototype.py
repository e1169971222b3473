"""Design check (CPU, torch) for the next kernel of DESIGN.md section 8b: a 3^n "same" convolution as ONE K-concatenated
GEMM over a zero-haloed, channel-last ROW SPACE.  Nothing here is product code; tests/test_conv_rowspace_design.py runs it
against ``F.conv2d`` / ``F.conv3d`` + autograd so that the index arithmetic the TMA producer will need is pinned before any
CUDA is written.

Row space.  A (B, C, *S) feature map is stored as X[R, C] with R = B * prod(S_i + 2): every spatial axis carries a one-voxel
halo of zeros, channels are contiguous.  In that space a filter tap d = (d_1..d_n), d_i in {-1, 0, 1}, is the CONSTANT row
offset off(d) = sum_i d_i * stride_i (stride_n = 1, stride_i = stride_{i+1} * (S_{i+1} + 2)), so

    forward   Y[r, :]  = sum_t X[r + off_t, :] @ W_t^T                  one GEMM, K = 3^n * C_in; the A tile of k-block kb is
                                                                        the 2-D box at row m0 + off[kb // (C_in / BK)]
    dgrad     dX[r, :] = sum_t dY[r - off_t, :] @ W_t                   the same kernel: offsets negated, weights transposed
    wgrad     dW_t     = dY^T @ X[. + off_t, :]                          3^n GEMMs (M = C_out, N = C_in, K = R) on the existing
                                                                        MN-major path, B operand shifted by off_t rows

Rows outside [0, R) read as zero (TMA out-of-bounds fill).  Halo rows of Y / dX are garbage by construction and must be
re-zeroed before the next convolution reads them (the following LayerNorm + GELU kernel gets a row-validity predicate);
halo rows of dY must be zero before dgrad / wgrad."""

from __future__ import annotations

import itertools
import math

import torch


def row_strides(spatial: tuple[int, ...]) -> list[int]:
    strides = [1] * len(spatial)
    for i in range(len(spatial) - 2, -1, -1):
        strides[i] = strides[i + 1] * (spatial[i + 1] + 2)
    return strides


def tap_offsets(spatial: tuple[int, ...]) -> list[int]:
    """Row offset of every tap, in the order of a (C_out, C_in, 3, .., 3) weight's trailing axes (row-major)."""
    st = row_strides(spatial)
    return [sum((k - 1) * s for k, s in zip(tap, st)) for tap in itertools.product(range(3), repeat=len(spatial))]


def to_rows(x: torch.Tensor) -> torch.Tensor:
    """(B, C, *S) -> zero-haloed channel-last rows (B * prod(S + 2), C)."""
    n = x.dim() - 2
    xp = torch.nn.functional.pad(x, (1, 1) * n)
    return xp.movedim(1, -1).reshape(-1, x.shape[1]).contiguous()


def from_rows(rows: torch.Tensor, batch: int, spatial: tuple[int, ...]) -> torch.Tensor:
    """rows (B * prod(S + 2), C) -> interior (B, C, *S)."""
    full = rows.reshape(batch, *[s + 2 for s in spatial], rows.shape[1]).movedim(-1, 1)
    return full[(slice(None), slice(None), *[slice(1, s + 1) for s in spatial])]


def interior_mask(batch: int, spatial: tuple[int, ...]) -> torch.Tensor:
    """(R,) bool: True on interior rows -- the predicate the halo re-zeroing needs, computable from the row index alone."""
    m = torch.zeros((batch, *[s + 2 for s in spatial]), dtype=torch.bool)
    m[(slice(None), *[slice(1, s + 1) for s in spatial])] = True
    return m.reshape(-1)


def shifted(x: torch.Tensor, off: int) -> torch.Tensor:
    """rows r -> x[r + off] with zero fill outside [0, R): what a 2-D TMA box at row (m0 + off) delivers."""
    out = torch.zeros_like(x)
    r = x.shape[0]
    lo, hi = max(0, -off), min(r, r - off)
    if hi > lo:
        out[lo:hi] = x[lo + off:hi + off]
    return out


def permute_weight(w: torch.Tensor) -> torch.Tensor:
    """(C_out, C_in, 3, .., 3) -> [C_out, taps * C_in], tap-major K: the B operand of the forward GEMM."""
    co, ci = w.shape[:2]
    return w.reshape(co, ci, -1).permute(0, 2, 1).reshape(co, -1).contiguous()


def conv_fwd(x_rows: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, spatial: tuple[int, ...]) -> torch.Tensor:
    offs = tap_offsets(spatial)
    a = torch.cat([shifted(x_rows, o) for o in offs], dim=1)  # [R, taps * C_in]: the K-concatenated A operand (never materialised on the GPU)
    y = a @ permute_weight(w).t()
    return y if bias is None else y + bias


def conv_dgrad(dy_rows: torch.Tensor, w: torch.Tensor, spatial: tuple[int, ...]) -> torch.Tensor:
    offs = tap_offsets(spatial)
    a = torch.cat([shifted(dy_rows, -o) for o in offs], dim=1)  # [R, taps * C_out]
    co, ci = w.shape[:2]
    wd = w.reshape(co, ci, -1).permute(1, 2, 0).reshape(ci, -1)  # [C_in, taps * C_out]
    return a @ wd.t()


def conv_wgrad(dy_rows: torch.Tensor, x_rows: torch.Tensor, w_shape: tuple[int, ...], spatial: tuple[int, ...]) -> torch.Tensor:
    offs = tap_offsets(spatial)
    co, ci = w_shape[:2]
    dw = torch.stack([dy_rows.t() @ shifted(x_rows, o) for o in offs], dim=2)  # [C_out, C_in, taps]
    return dw.reshape(co, ci, *w_shape[2:])


def halo_overhead(spatial: tuple[int, ...]) -> float:
    return math.prod(s + 2 for s in spatial) / math.prod(spatial)
