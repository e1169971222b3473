"""GPU checks of the GEMM convolution (cb_conv_gemm_bf16, cinema_b200/conv_gemm.py): the kernel entry point with its
autograd function against F.conv2d / conv3d, and the opt-in ``native`` switch of ConvResBlock against its cuDNN path.  (Opt-in
behind CB_EXPERIMENTAL_CONV=1 until the kernel was validated on the B200 in round 2; part of the default GPU run since.)"""

import pytest
import torch
import torch.nn.functional as F  # noqa: N812

pytestmark = [pytest.mark.gpu]

DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).norm() / (b.norm() + 1e-30)).detach())


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


@pytest.mark.parametrize(("nd", "cin", "cout", "shape"), [(3, 64, 64, (2, 64, 24, 20, 6)), (3, 128, 64, (1, 128, 12, 12, 8)),
                                                            (2, 96, 40, (2, 96, 30, 28)), (3, 32, 32, (1, 32, 16, 16, 4))])
def test_conv_res_block_native_switch_against_cudnn(nd, cin, cout, shape, monkeypatch):
    """``ConvResBlock.native`` (ConvUNetR.set_native_convs): both 3^n convolutions as K-concatenated GEMMs over the haloed row
    space -- channel counts that are not multiples of 64 are zero-padded on both sides -- against the same block on cuDNN under
    bf16 autocast: output, input gradient, weight and bias gradients within bf16 rounding of each other."""
    from cinema_b200 import conv as C

    monkeypatch.setattr(C, "_NATIVE_MIN_C", 32)
    torch.manual_seed(0)
    blk = C.ConvResBlock(n_dims=nd, in_chans=cin, out_chans=cout, norm="layer").to(DEV)
    x = torch.randn(shape, device=DEV, requires_grad=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref = blk(x).float()
    ref.square().mean().backward()
    want = [x.grad.clone(), blk.conv1.weight.grad.clone(), blk.conv2.weight.grad.clone(), blk.conv2.bias.grad.clone()]
    x.grad = None
    blk.zero_grad()
    blk.native = True
    assert blk._native_ok(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = blk(x).float()
    out.square().mean().backward()
    got = [x.grad, blk.conv1.weight.grad, blk.conv2.weight.grad, blk.conv2.bias.grad]
    assert rel(out, ref) < 1e-2
    for g, w in zip(got, want):
        assert rel(g, w) < 2.5e-2
