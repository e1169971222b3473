"""Host logic of the EXPERIMENTAL GEMM convolution (cinema_b200/conv_gemm.py) through the emulated kernels: forward, input /
weight / bias gradients and the fused skip against torch's convolutions (bf16 operands, fp32 accumulation: 2e-2 relative)."""

import pytest
import torch
import torch.nn.functional as F  # noqa: N812

from cinema_b200 import conv_gemm as CG


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize(("spatial", "cin", "cout", "skip"), [((6, 5, 4), 64, 64, False), ((4, 6), 64, 128, False), ((5, 4, 3), 128, 64, True)])
def test_conv3x3_rows_forward_backward(spatial, cin, cout, skip, emulated_kernels):
    torch.manual_seed(0)
    b = 2
    space = CG.RowSpace(b, spatial)
    assert len(space.offsets) == 3 ** len(spatial) and space.offsets[len(space.offsets) // 2] == 0
    conv = F.conv2d if len(spatial) == 2 else F.conv3d
    x = torch.randn(b, cin, *spatial)
    w = (torch.randn(cout, cin, *([3] * len(spatial))) * 0.05).requires_grad_()
    bias = torch.randn(cout, requires_grad=True)
    xr = x.to(torch.bfloat16).float().requires_grad_()
    w_ref = w.detach().to(torch.bfloat16).float().requires_grad_()  # the reference differentiates its own leaves
    b_ref = bias.detach().clone().requires_grad_()
    ref = conv(xr, w_ref, b_ref, padding=1)
    x_rows = space.to_rows(x).requires_grad_()
    skip32 = None
    if skip:
        skip32 = space.new_rows(cout, "cpu", torch.float32)
        skip32.copy_(torch.randn(space.rows, cout))
        space.zero_halo_(skip32)
        ref = ref + space.from_rows(skip32)
    y_rows = CG.conv3x3(x_rows, w, bias, space, skip32)
    assert y_rows.dtype == torch.bfloat16 and y_rows.shape == (space.rows, cout)
    halo = space.interior("cpu")[:, 0] == 0
    assert bool((y_rows[halo] == 0).all())  # the invariant the next layer relies on
    assert rel(space.from_rows(y_rows), ref) < 2e-2
    g = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(g)
    g_rows = space.to_rows(g)
    y_rows.backward(g_rows + 7.0 * halo[:, None])  # garbage on the halo of the upstream gradient must not matter
    assert rel(space.from_rows(x_rows.grad), xr.grad) < 2e-2
    assert bool((x_rows.grad[halo] == 0).all())
    assert rel(w.grad, w_ref.grad) < 2e-2
    assert rel(bias.grad, b_ref.grad) < 2e-2


def test_rowspace_views_and_errors(emulated_kernels):
    space = CG.RowSpace(1, (4, 4, 4))
    rows = space.to_rows(torch.randn(1, 64, 4, 4, 4))
    assert rows.shape == (6 * 6 * 6, 64) and space.guard == 6 * 6 + 6 + 1
    v = space.shifted_view(rows, -space.guard)
    assert v.shape == rows.shape and bool((v[: space.guard] == 0).all())
    assert torch.equal(space.shifted_view(rows, 5)[:-5], rows[5:])
    with pytest.raises(ValueError):
        space.shifted_view(torch.zeros(space.rows, 64, dtype=torch.bfloat16), 1)  # no guard rows
    with pytest.raises(ValueError):
        CG.conv3x3(rows, torch.zeros(64, 64, 5, 5, 5), None, space)
    with pytest.raises(ValueError):
        CG.conv3x3(rows[:10], torch.zeros(64, 64, 3, 3, 3), None, space)
