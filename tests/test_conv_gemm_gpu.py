"""GPU checks of the EXPERIMENTAL GEMM convolution (cb_conv_gemm_bf16, cinema_b200/conv_gemm.py).  The kernel variant was
written at the end of round 1 after the GPU budget was spent, so these tests are opt-in until it has been validated:

    CB_EXPERIMENTAL_CONV=1 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q
"""

import os

import pytest
import torch
import torch.nn.functional as F  # noqa: N812

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CB_EXPERIMENTAL_CONV") != "1", reason="experimental kernel: opt in with CB_EXPERIMENTAL_CONV=1")]

DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize(("spatial", "cin", "cout"), [((12, 12, 16), 512, 512), ((24, 24, 16), 256, 256), ((48, 48, 16), 64, 64),
                                                      ((96, 96), 64, 128), ((5, 7, 3), 64, 64), ((48, 48, 16), 128, 64)])
def test_conv_gemm_kernel_against_torch(spatial, cin, cout):
    from cinema_b200 import conv_gemm as CG

    torch.manual_seed(0)
    b = 2
    space = CG.RowSpace(b, spatial)
    conv = F.conv2d if len(spatial) == 2 else F.conv3d
    x = torch.randn(b, cin, *spatial, device=DEV)
    w = (torch.randn(cout, cin, *([3] * len(spatial)), device=DEV) * (cin * 27) ** -0.5).requires_grad_()
    bias = torch.randn(cout, device=DEV, requires_grad=True)
    xr = x.to(torch.bfloat16).float().requires_grad_()
    w_ref = w.detach().to(torch.bfloat16).float().requires_grad_()
    b_ref = bias.detach().clone().requires_grad_()
    ref = conv(xr, w_ref, b_ref, padding=1)
    x_rows = space.to_rows(x).requires_grad_()
    y_rows = CG.conv3x3(x_rows, w, bias, space)
    torch.cuda.synchronize()
    halo = space.interior(DEV)[:, 0] == 0
    assert bool((y_rows[halo] == 0).all())
    assert rel(space.from_rows(y_rows), ref) < 1e-2
    g = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(g)
    y_rows.backward(space.to_rows(g))
    torch.cuda.synchronize()
    assert rel(space.from_rows(x_rows.grad), xr.grad) < 1e-2
    assert rel(w.grad, w_ref.grad) < 1e-2
    assert rel(bias.grad, b_ref.grad) < 1e-2
