"""Convolutional stem layers with the reference's module API (cinema/conv.py).

Class names, constructor signatures and parameter names follow the reference so that the state
dict is key-compatible (``conv_blocks.{i}.patch_embed.{conv,norm}``, ``conv.{j}.{norm1,norm2,conv1,conv2,
dw_conv,mlp.fc1,mlp.fc2}``).  ``set_grad_ckpt`` is kept for API compatibility and is a no-op: on
a 180 GB B200 the activations are saved, not recomputed.

``Linear`` runs on the tcgen05 GEMM.  The ConvMAE stem blocks (``ConvNormActBlock``,
``MaskedConvBlock``) keep a cuDNN forward for standalone use; the MAE hot path (``CineMA.forward``)
drives the stem through ``cinema_b200.stem`` instead.
"""

from __future__ import annotations

import torch
from torch import nn

from cinema_b200 import _C

KernelSizeType = tuple[int, ...] | int
F32, BF16 = torch.float32, torch.bfloat16


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM; bf16 operands, fp32 accumulation, fp32 weight gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        k = weight.shape[1]
        x2 = x.detach().reshape(-1, k)
        x16 = torch.empty(x2.shape, dtype=BF16, device=x.device)
        if x2.dtype == BF16:
            x16 = x2.contiguous()
        else:
            _C.cast_bf16(x2.to(F32).contiguous(), x16)
        w16 = torch.empty(weight.shape, dtype=BF16, device=x.device)
        _C.cast_bf16(weight.detach().contiguous(), w16)
        out = torch.empty((x16.shape[0], weight.shape[0]), dtype=F32, device=x.device)
        _C.gemm(x16, w16, out, bias=bias.detach() if bias is not None else None)
        ctx.save_for_backward(x16, w16)
        ctx.meta = (x.shape, x.dtype, bias is not None)
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        xshape, xdtype, has_bias = ctx.meta
        dy2 = dy.detach().reshape(-1, w16.shape[0]).to(F32).contiguous()
        dy16 = torch.empty(dy2.shape, dtype=BF16, device=dy.device)
        _C.cast_bf16(dy2, dy16)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x16.shape, dtype=F32, device=dy.device)
            _C.gemm(dy16, w16, dx, b_mn=True)
            dx = dx.view(xshape).to(xdtype)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(w16.shape, dtype=F32, device=dy.device)
            _C.gemm(dy16, x16, dw, a_mn=True, b_mn=True, accumulate=True)
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(w16.shape[0], dtype=F32, device=dy.device)
            _C.colsum(dy16, db)
        return dx, dw, db


class Linear(nn.Linear):
    """nn.Linear executed by the sm_100a GEMM (cinema/conv.py:21-36)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.grad_ckpt = False

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _LinearFn.apply(x, self.weight, self.bias)


class Conv2d(nn.Conv2d):
    """cinema/conv.py:39-54."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.grad_ckpt = False

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable


class Conv3d(nn.Conv3d):
    """cinema/conv.py:57-72."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.grad_ckpt = False

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable


class ConvMlp(nn.Module):
    """1x1-conv MLP: fc2(GELU(fc1 x)) (cinema/conv.py:111-166)."""

    def __init__(self, n_dims, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0) -> None:
        if n_dims not in {2, 3}:
            raise ValueError(f"Invalid n_dims, must be 2 or 3, got {n_dims}.")
        super().__init__()
        self.grad_ckpt = False
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        bias = tuple(bias) if isinstance(bias, (tuple, list)) else (bias, bias)
        drop = tuple(drop) if isinstance(drop, (tuple, list)) else (drop, drop)
        conv_cls = Conv2d if n_dims == 2 else Conv3d
        self.fc1 = conv_cls(in_features, hidden_features, kernel_size=1, bias=bias[0])
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop[0])
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = conv_cls(hidden_features, out_features, kernel_size=1, bias=bias[1])
        self.drop2 = nn.Dropout(drop[1])

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        self.fc1.set_grad_ckpt(enable)
        self.fc2.set_grad_ckpt(enable)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class ConvLayerNorm(nn.LayerNorm):
    """LayerNorm over the channel axis of a channel-first tensor (cinema/conv.py:169-187).

    The reference permutes to channel-last, normalises and copies back to NC(D)HW -- two transposes per call.  With
    ``keep_channels_last`` the result keeps the channel-last strides it is produced with (same logical shape and values):
    the next LayerNorm's permute is then a free view and cuDNN picks its NDHWC kernels for the convolutions in between."""

    keep_channels_last = False

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = super().forward(x.movedim(1, -1))
        x = x.movedim(-1, 1)
        return x if self.keep_channels_last else x.contiguous()


def get_conv_norm(n_dims: int, in_chans: int, norm: str, eps: float = 1e-6, n_groups: int = 32) -> nn.Module:
    """cinema/conv.py:190-210."""
    if norm == "instance":
        return nn.InstanceNorm2d(in_chans, eps=eps) if n_dims == 2 else nn.InstanceNorm3d(in_chans, eps=eps)
    if norm == "layer":
        return ConvLayerNorm(in_chans, eps=eps)
    if norm == "group":
        return nn.GroupNorm(num_groups=min(n_groups, in_chans), num_channels=in_chans, eps=eps, affine=True)
    raise ValueError(f"Invalid norm type, got {norm}, must be 'instance' or 'layer' or 'group'.")


class ConvNormActBlock(nn.Module):
    """conv -> norm -> act (cinema/conv.py:213-273)."""

    def __init__(self, n_dims, in_chans, out_chans, norm, kernel_size: KernelSizeType = 3, stride: KernelSizeType = 1,
                 padding: str = "same", act_layer=nn.GELU) -> None:
        if n_dims not in {2, 3}:
            raise ValueError(f"Invalid n_dims, must be 2 or 3, got {n_dims}.")
        if not isinstance(kernel_size, int) and len(kernel_size) != n_dims:
            raise ValueError(f"Invalid kernel_size {kernel_size}, must be an integer or a tuple of {n_dims} integers.")
        if not isinstance(stride, int) and len(stride) != n_dims:
            raise ValueError(f"Invalid stride {stride}, must be an integer or a tuple of {n_dims} integers.")
        super().__init__()
        self.grad_ckpt = False
        conv_cls = Conv2d if n_dims == 2 else Conv3d
        self.conv = conv_cls(in_chans, out_chans, kernel_size=kernel_size, stride=stride, padding=padding)
        self.norm = get_conv_norm(n_dims=n_dims, in_chans=out_chans, norm=norm)
        self.act = act_layer()

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        self.conv.set_grad_ckpt(enable)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.act(self.norm(self.conv(x)))


class MaskedConvBlock(nn.Module):
    """ConvMAE block: x += conv2(dw5(mask * conv1(norm1 x)));  x += mlp(norm2 x)  (cinema/conv.py:349-415)."""

    def __init__(self, n_dims, in_chans, mlp_ratio=4, dropout=0.0, drop_path=0.0, act_layer=nn.GELU, norm="layer") -> None:
        if n_dims not in {2, 3}:
            raise ValueError(f"Invalid n_dims, must be 2 or 3, got {n_dims}.")
        super().__init__()
        self.grad_ckpt = False
        self.norm1 = get_conv_norm(n_dims=n_dims, in_chans=in_chans, norm=norm)
        self.norm2 = get_conv_norm(n_dims=n_dims, in_chans=in_chans, norm=norm)
        conv_cls = Conv2d if n_dims == 2 else Conv3d
        self.conv1 = conv_cls(in_chans, in_chans, kernel_size=1, padding="same")
        self.conv2 = conv_cls(in_chans, in_chans, kernel_size=1, padding="same")
        self.dw_conv = conv_cls(in_chans, in_chans, kernel_size=5, padding="same", groups=in_chans)
        if drop_path > 0.0:  # never enabled by the reference configs; the dense (cuDNN) form below supports it
            from cinema_b200.vit import DropPath

            self.drop_path = DropPath(drop_path)
        else:
            self.drop_path = nn.Identity()
        self.mlp = ConvMlp(n_dims=n_dims, in_features=in_chans, hidden_features=in_chans * mlp_ratio,
                           act_layer=act_layer, drop=dropout)

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        for m in (self.conv1, self.conv2, self.dw_conv, self.mlp):
            m.set_grad_ckpt(enable)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        """x (B, C, *spatial); mask (B, *spatial), 1 = visible."""
        h = self.conv1(self.norm1(x))
        if mask is not None:
            h = mask.unsqueeze(1).to(h.dtype) * h
        x = x + self.drop_path(self.conv2(self.dw_conv(h)))
        return x + self.drop_path(self.mlp(self.norm2(x)))


class ConvTranspose2d(nn.ConvTranspose2d):
    """nn.ConvTranspose2d with the reference's grad-ckpt switch (cinema/conv.py:75-89); kept as an attribute only."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.grad_ckpt = False

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable


class ConvTranspose3d(nn.ConvTranspose3d):
    """nn.ConvTranspose3d with the reference's grad-ckpt switch (cinema/conv.py:92-106); kept as an attribute only."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.grad_ckpt = False

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable


# narrowest layer the opt-in GEMM convolution takes (CB_NATIVE_CONV_MIN_C for measurements; 32-channel layers work -- padded to
# 64 on both sides -- but measured slower than cuDNN at full resolution, tools/bench_finetune.py)
_NATIVE_MIN_C = int(__import__("os").environ.get("CB_NATIVE_CONV_MIN_C", "64"))


class ConvResBlock(nn.Module):
    """x -> conv2(drop(act(norm2(conv1(act(norm1 x)))))) + shortcut(x)  (cinema/conv.py:276-346): the residual unit of
    the segmentation decoder.  Dense k x k (x k) convolutions at up to full image resolution: cuDNN under bf16 autocast."""

    def __init__(self, n_dims, in_chans, out_chans, norm, dropout: float = 0.0, kernel_size: KernelSizeType = 3,
                 act_layer=nn.GELU) -> None:
        if n_dims not in {2, 3}:
            raise ValueError(f"Invalid n_dims, must be 2 or 3, got {n_dims}.")
        if not isinstance(kernel_size, int) and len(kernel_size) != n_dims:
            raise ValueError(f"Invalid kernel_size {kernel_size}, must be an integer or a tuple of {n_dims} integers.")
        super().__init__()
        self.grad_ckpt = False
        conv_cls = Conv2d if n_dims == 2 else Conv3d
        self.norm1 = get_conv_norm(n_dims=n_dims, in_chans=in_chans, norm=norm)
        self.norm2 = get_conv_norm(n_dims=n_dims, in_chans=out_chans, norm=norm)
        self.conv1 = conv_cls(in_chans, out_chans, kernel_size=kernel_size, padding="same")
        self.conv2 = conv_cls(out_chans, out_chans, kernel_size=kernel_size, padding="same")
        self.dropout = nn.Dropout(dropout)
        self.act = act_layer()
        self.shortcut = conv_cls(in_chans, out_chans, kernel_size=1) if in_chans != out_chans else nn.Identity()

    @torch.jit.ignore
    def set_grad_ckpt(self, enable: bool = True) -> None:
        self.grad_ckpt = enable
        self.conv1.set_grad_ckpt(enable)
        self.conv2.set_grad_ckpt(enable)
        if hasattr(self.shortcut, "set_grad_ckpt"):
            self.shortcut.set_grad_ckpt(enable)

    native = False  # opt-in (ConvUNetR.set_native_convs): the two k^n convolutions on the tcgen05 GEMM instead of cuDNN
    _spaces: dict | None = None

    def _native_ok(self, x: torch.Tensor) -> bool:
        if not (self.native and x.is_cuda and x.dim() in (4, 5)):
            return False
        for conv in (self.conv1, self.conv2):
            if (tuple(conv.kernel_size) != (3,) * (x.dim() - 2) or tuple(conv.stride) != (1,) * (x.dim() - 2)
                    or tuple(conv.dilation) != (1,) * (x.dim() - 2) or conv.groups != 1 or conv.padding != "same"
                    or conv.in_channels < _NATIVE_MIN_C or conv.out_channels < _NATIVE_MIN_C or conv.out_channels % 8 != 0):
                return False  # (narrow layers would be zero-padded to 64 channels on both sides: left to cuDNN)
        return True

    def _space(self, x: torch.Tensor):
        from cinema_b200.conv_gemm import RowSpace

        key = (x.shape[0], tuple(x.shape[2:]))
        if self._spaces is None:
            self._spaces = {}
        if key not in self._spaces:
            if len(self._spaces) > 8:
                self._spaces.clear()
            self._spaces[key] = RowSpace(*key)
        return self._spaces[key]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._native_ok(x):
            from cinema_b200.conv_gemm import conv3x3_dense

            space = self._space(x)
            h = conv3x3_dense(self.act(self.norm1(x)), self.conv1.weight, self.conv1.bias, space)
            h = conv3x3_dense(self.dropout(self.act(self.norm2(h))), self.conv2.weight, self.conv2.bias, space)
            return h + self.shortcut(x)
        h = self.conv1(self.act(self.norm1(x)))
        h = self.conv2(self.dropout(self.act(self.norm2(h))))
        return h + self.shortcut(x)
