"""Pins the index arithmetic of the planned GEMM convolution (tools/conv_rowspace_prototype.py) against
torch's convolutions and autograd: forward, input gradient, weight gradient, halo handling, 2-D and 3-D."""

import pytest
import torch
import torch.nn.functional as F  # noqa: N812

from tools import conv_rowspace_prototype as P


@pytest.mark.parametrize(("spatial", "cin", "cout"), [((6, 5), 3, 4), ((4, 6, 3), 2, 5), ((12, 12, 16), 8, 8), ((5, 4, 1), 3, 2)])
def test_rowspace_convolution_equals_torch(spatial, cin, cout):
    torch.manual_seed(0)
    b = 2
    conv = F.conv2d if len(spatial) == 2 else F.conv3d
    x = torch.randn(b, cin, *spatial, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, *([3] * len(spatial)), dtype=torch.float64, requires_grad=True)
    bias = torch.randn(cout, dtype=torch.float64)
    y_ref = conv(x, w, bias, padding=1)
    g = torch.randn_like(y_ref)
    y_ref.backward(g)
    rows = P.to_rows(x.detach())
    assert rows.shape == (b * int(torch.tensor([s + 2 for s in spatial]).prod()), cin)
    y_rows = P.conv_fwd(rows, w.detach(), bias, spatial)
    torch.testing.assert_close(P.from_rows(y_rows, b, spatial), y_ref.detach(), rtol=1e-10, atol=1e-10)
    # the upstream gradient lives in the same row space with a ZERO halo
    dy_rows = P.to_rows(g)
    mask = P.interior_mask(b, spatial)
    assert bool((dy_rows[~mask] == 0).all()) and int(mask.sum()) == b * int(torch.tensor(spatial).prod())
    dx_rows = P.conv_dgrad(dy_rows, w.detach(), spatial)
    torch.testing.assert_close(P.from_rows(dx_rows, b, spatial), x.grad, rtol=1e-10, atol=1e-10)
    dw = P.conv_wgrad(dy_rows, rows, tuple(w.shape), spatial)
    torch.testing.assert_close(dw, w.grad, rtol=1e-10, atol=1e-10)
    torch.testing.assert_close(dy_rows.sum(0), (g.movedim(1, -1).reshape(-1, cout)).sum(0))  # bias gradient = column sums


def test_halo_rows_are_garbage_and_must_be_rezeroed():
    """Forward output on halo rows is NOT zero (taps reach into the interior): chaining two convolutions without re-zeroing
    breaks the second one, with the re-zeroing it equals torch."""
    torch.manual_seed(1)
    spatial, b = (5, 4, 3), 1
    x = torch.randn(b, 3, *spatial, dtype=torch.float64)
    w1 = torch.randn(3, 3, 3, 3, 3, dtype=torch.float64)
    w2 = torch.randn(2, 3, 3, 3, 3, dtype=torch.float64)
    ref = F.conv3d(F.conv3d(x, w1, padding=1), w2, padding=1)
    mid = P.conv_fwd(P.to_rows(x), w1, None, spatial)
    mask = P.interior_mask(b, spatial)
    assert float(mid[~mask].abs().max()) > 0
    wrong = P.from_rows(P.conv_fwd(mid, w2, None, spatial), b, spatial)
    assert not torch.allclose(wrong, ref)
    mid = mid * mask[:, None]
    torch.testing.assert_close(P.from_rows(P.conv_fwd(mid, w2, None, spatial), b, spatial), ref, rtol=1e-10, atol=1e-10)


def test_tap_offsets_and_halo_cost():
    assert P.tap_offsets((4, 5)) == [-(5 + 2) - 1, -(5 + 2), -(5 + 2) + 1, -1, 0, 1, (5 + 2) - 1, 5 + 2, (5 + 2) + 1]
    offs = P.tap_offsets((192, 192, 16))
    assert len(offs) == 27 and offs[13] == 0 and offs[0] == -((192 + 2) * 18 + 18 + 1) and offs == sorted(offs)
    assert abs(P.halo_overhead((192, 192, 16)) - 1.148) < 1e-3 and abs(P.halo_overhead((12, 12, 16)) - 1.531) < 1e-3
