"""3^n "same" convolution on the tcgen05 GEMM pipeline.  The kernel entry point
(``cb_conv_gemm_bf16``) is validated on the B200 against ``F.conv2d / conv3d`` (``tests/test_conv_gemm_gpu.py``), the
arithmetic is pinned on the CPU (``tools/conv_rowspace_prototype.py``) and this host logic is exercised through the emulated
kernels (``tests/test_conv_gemm_host.py``).  ``ConvResBlock`` (the residual unit of the UNETR decoder) uses it when its
``native`` switch is on (``ConvUNetR.set_native_convs``, opt-in).

Feature maps live in a zero-haloed channel-last ROW SPACE: a (B, C, *S) map is the matrix X[B * prod(S + 2), C] (bf16).
``RowSpace`` owns the geometry (tap offsets, interior mask, guard rows); ``Conv3x3Fn`` is the autograd function:

    forward   one ``cb_conv_gemm_bf16``                      K = 3^n C_in, bias (+ skip) fused in the epilogue
    dgrad     the same entry point                           offsets negated, weights [C_in, 3^n C_out]
    wgrad     3^n ``cb_gemm_bf16`` (MN-major operands)       dW_t += dY^T X[. + off_t]; X sits between guard rows of zeros so the
                                                             shifted B operand is a plain pointer offset; fp32 accumulation
    bias      ``cb_colsum_bf16`` of dY

Outputs carry garbage on halo rows; ``RowSpace.zero_halo_`` restores the invariant before the next convolution reads them
(to be fused into the LayerNorm + GELU kernel that always follows in ``ConvResBlock``)."""

from __future__ import annotations

import itertools
import math

import torch

from cinema_b200 import _C

BF16 = torch.bfloat16
F32 = torch.float32


class RowSpace:
    """Geometry of the zero-haloed row space of (batch, *spatial) feature maps."""

    def __init__(self, batch: int, spatial: tuple[int, ...]) -> None:
        self.batch, self.spatial = batch, tuple(spatial)
        self.padded = tuple(s + 2 for s in spatial)
        self.rows = batch * math.prod(self.padded)
        strides = [1] * len(spatial)
        for i in range(len(spatial) - 2, -1, -1):
            strides[i] = strides[i + 1] * self.padded[i + 1]
        self.offsets = [sum((k - 1) * s for k, s in zip(tap, strides)) for tap in itertools.product(range(3), repeat=len(spatial))]
        self.guard = max(self.offsets)  # rows of zeros kept before and after X for the shifted wgrad operand
        self._mask: dict[str, torch.Tensor] = {}

    def interior(self, device) -> torch.Tensor:
        """(rows, 1) bf16 {0, 1}: 1 on interior rows."""
        key = str(device)
        if key not in self._mask:
            m = torch.zeros((self.batch, *self.padded), dtype=BF16, device=device)
            m[(slice(None), *[slice(1, s + 1) for s in self.spatial])] = 1
            self._mask[key] = m.reshape(-1, 1)
        return self._mask[key]

    def to_rows(self, x: torch.Tensor, channels: int | None = None) -> torch.Tensor:
        """(B, C, *S) any float dtype -> guarded row-space storage; returns the (rows, C) bf16 view between the guards.
        ``channels`` > C pads the channel axis with zeros (the GEMM needs C_in in multiples of 64)."""
        c = x.shape[1]
        cp = c if channels is None else channels
        store = torch.zeros((self.rows + 2 * self.guard, cp), dtype=BF16, device=x.device)
        body = store[self.guard:self.guard + self.rows].view(self.batch, *self.padded, cp)
        body[(slice(None), *[slice(1, s + 1) for s in self.spatial], slice(0, c))] = x.movedim(1, -1).to(BF16)
        return store[self.guard:self.guard + self.rows]

    def from_rows(self, rows: torch.Tensor) -> torch.Tensor:
        full = rows.reshape(self.batch, *self.padded, rows.shape[1]).movedim(-1, 1)
        return full[(slice(None), slice(None), *[slice(1, s + 1) for s in self.spatial])]

    def new_rows(self, channels: int, device, dtype=BF16) -> torch.Tensor:
        store = torch.zeros((self.rows + 2 * self.guard, channels), dtype=dtype, device=device)
        return store[self.guard:self.guard + self.rows]

    def zero_halo_(self, rows: torch.Tensor) -> torch.Tensor:
        return rows.mul_(self.interior(rows.device).to(rows.dtype))

    def shifted_view(self, rows: torch.Tensor, off: int) -> torch.Tensor:
        """rows r -> rows[r + off] as a VIEW into the guarded storage (no copy): valid for tensors made by this class."""
        base = rows._base if rows._base is not None else rows
        start = rows.storage_offset() // rows.shape[1] + off
        if start < 0 or start + self.rows > base.shape[0]:
            raise ValueError("tensor was not allocated with guard rows (use RowSpace.to_rows / new_rows)")
        return base[start:start + self.rows]


def tap_major(weight: torch.Tensor) -> torch.Tensor:
    """(C_out, C_in, 3, .., 3) -> (C_out, taps * C_in) bf16, the forward B operand."""
    co, ci = weight.shape[:2]
    return weight.reshape(co, ci, -1).permute(0, 2, 1).reshape(co, -1).to(BF16).contiguous()


def tap_major_transposed(weight: torch.Tensor) -> torch.Tensor:
    """(C_out, C_in, 3, .., 3) -> (C_in, taps * C_out) bf16, the dgrad B operand (used with negated offsets)."""
    co, ci = weight.shape[:2]
    return weight.reshape(co, ci, -1).permute(1, 2, 0).reshape(ci, -1).to(BF16).contiguous()


class Conv3x3Fn(torch.autograd.Function):
    """y_rows = conv3^n(x_rows) + bias (+ skip32) over a RowSpace; x_rows must have a zero halo and guard rows."""

    @staticmethod
    def forward(ctx, x_rows, weight, bias, space: RowSpace, skip32=None):
        y = space.new_rows(weight.shape[0], x_rows.device)
        _C.conv_gemm(x_rows, tap_major(weight), y, space.offsets, bias=bias.detach().float() if bias is not None else None,
                     residual=skip32)
        space.zero_halo_(y)
        ctx.space = space
        ctx.save_for_backward(x_rows, weight)
        ctx.has_bias, ctx.has_skip = bias is not None, skip32 is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        space: RowSpace = ctx.space
        x_rows, weight = ctx.saved_tensors
        co, ci = weight.shape[:2]
        dy16 = space.new_rows(co, dy.device)
        dy16.copy_(dy)
        space.zero_halo_(dy16)  # the halo of the upstream gradient must not leak into dX / dW
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = space.new_rows(ci, dy.device)
            _C.conv_gemm(dy16, tap_major_transposed(weight), dx, [-o for o in space.offsets])
            space.zero_halo_(dx)
        if ctx.needs_input_grad[1]:
            taps = len(space.offsets)
            dwt = torch.zeros((co, taps * ci), dtype=F32, device=dy.device)  # tap-major, accumulated in fp32
            for t, off in enumerate(space.offsets):
                _C.gemm(dy16, space.shifted_view(x_rows, off), dwt[:, t * ci:(t + 1) * ci], a_mn=True, b_mn=True, accumulate=True)
            dw = dwt.view(co, taps, ci).permute(0, 2, 1).reshape(weight.shape).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(co, dtype=F32, device=dy.device)
            _C.colsum(dy16, db)
        dskip = dy.float() if ctx.has_skip and ctx.needs_input_grad[4] else None
        return dx, dw, db, None, dskip


def conv3x3(x_rows: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, space: RowSpace,
            skip32: torch.Tensor | None = None) -> torch.Tensor:
    if weight.shape[2:] != (3,) * len(space.spatial):
        raise ValueError(f"kernel {tuple(weight.shape[2:])} is not 3^{len(space.spatial)}")
    if x_rows.shape != (space.rows, weight.shape[1]):
        raise ValueError(f"x_rows {tuple(x_rows.shape)} does not match the row space ({space.rows}, {weight.shape[1]})")
    return Conv3x3Fn.apply(x_rows, weight, bias, space, skip32)


def conv3x3_dense(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, space: RowSpace) -> torch.Tensor:
    """(B, C_in, *S) feature map -> (B, C_out, *S) bf16 (a channel-last strided VIEW of the haloed output rows): the same
    contract as ``F.conv{2,3}d(x, weight, bias, padding="same")`` under bf16 autocast.  C_in that is not a multiple of 64 is
    zero-padded to one (activations and weights; the wasted k-blocks are the price of the 128B-swizzled operand tiles), and
    so is C_out (zero filters whose outputs are sliced away: the output gradient then has the width the dgrad GEMM needs)."""
    import torch.nn.functional as F

    c_out, c_in = weight.shape[:2]
    cp, cop = (c_in + 63) // 64 * 64, (c_out + 63) // 64 * 64
    if cp != c_in or cop != c_out:  # (C_out too: it is the contraction width of the dgrad GEMM)
        weight = F.pad(weight, (0, 0) * (weight.dim() - 2) + (0, cp - c_in, 0, cop - c_out))
        if bias is not None and cop != c_out:
            bias = F.pad(bias, (0, cop - c_out))
    y = space.from_rows(conv3x3(space.to_rows(x, cp), weight, bias, space))
    return y[:, :c_out] if cop != c_out else y
