"""GPU parity of the tcgen05 attention kernels against an fp32 PyTorch reference of the same op
(softmax(q k^T / sqrt(d)) v, cinema/vit.py:505-517) on bf16-representable inputs.

Tolerance: the kernel rounds P to bf16 before P.V (like every flash implementation) and the output
to bf16, so outputs agree to ~2^-8 relative per element; we assert a norm-wise relative error of
4e-3 for outputs / gradients and 1e-4 absolute for the log-sum-exp."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cinema_b200 import _C

DEV = "cuda"


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_qkv(b, h, nq, nk, d, seed, fused):
    g = torch.Generator(device=DEV).manual_seed(seed)
    if fused and nq == nk:  # (B, N, 3, H, d) projection output consumed in place
        qkv = (torch.randn(b, nq, 3, h, d, device=DEV, generator=g)).to(torch.bfloat16)
        return qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    q = torch.randn(b, nq, h, d, device=DEV, generator=g).to(torch.bfloat16)
    kv = torch.randn(b, nk, 2, h, d, device=DEV, generator=g).to(torch.bfloat16)
    return q, kv[:, :, 0], kv[:, :, 1]


def reference(q, k, v, scale):
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))  # (B, H, N, d)
    s = qf @ kf.transpose(-1, -2) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ vf
    return o.permute(0, 2, 1, 3), lse


CASES = [
    # b, h, nq, nk, d, fused
    (2, 3, 128, 128, 64, False), (2, 2, 256, 256, 64, True), (1, 2, 577, 577, 64, True), (2, 12, 769, 769, 64, True),
    (1, 2, 130, 70, 64, False), (1, 1, 1, 5, 64, False), (2, 4, 300, 768, 32, False), (1, 16, 2305, 768, 32, False),
    (1, 2, 64, 1000, 32, False), (1, 1, 257, 129, 32, False),
]


@pytest.mark.parametrize("b,h,nq,nk,d,fused", CASES)
def test_attention_forward(b, h, nq, nk, d, fused):
    q, k, v = make_qkv(b, h, nq, nk, d, seed=nq + nk, fused=fused)
    scale = d ** -0.5
    o = torch.zeros(b, nq, h, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(b, h, nq, device=DEV)
    _C.attention_fwd(q, k, v, o, lse, scale)
    o_ref, lse_ref = reference(q, k, v, scale)
    assert rel_err(o, o_ref) < 4e-3
    torch.testing.assert_close(lse, lse_ref, rtol=0, atol=1e-4)


def test_attention_forward_large_logits_lazy_rescale():
    # rows whose running max keeps growing across key tiles exercise the O-rescale path
    b, h, n, d = 1, 2, 512, 64
    g = torch.Generator(device=DEV).manual_seed(0)
    q = (torch.randn(b, n, h, d, device=DEV, generator=g) * 3).to(torch.bfloat16)
    k = (torch.randn(b, n, h, d, device=DEV, generator=g) * 3).to(torch.bfloat16)
    k = k * torch.linspace(0.2, 2.0, n, device=DEV).view(1, n, 1, 1).to(torch.bfloat16)
    v = torch.randn(b, n, h, d, device=DEV, generator=g).to(torch.bfloat16)
    o = torch.empty(b, n, h, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(b, h, n, device=DEV)
    _C.attention_fwd(q, k, v, o, lse, d ** -0.5)
    o_ref, lse_ref = reference(q, k, v, d ** -0.5)
    assert rel_err(o, o_ref) < 5e-3
    torch.testing.assert_close(lse, lse_ref, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("b,h,nq,nk,d,fused", CASES)
def test_attention_backward(b, h, nq, nk, d, fused):
    q, k, v = make_qkv(b, h, nq, nk, d, seed=7 + nq + nk, fused=fused)
    scale = d ** -0.5
    o = torch.empty(b, nq, h, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(b, h, nq, device=DEV)
    _C.attention_fwd(q, k, v, o, lse, scale)
    g = torch.Generator(device=DEV).manual_seed(1)
    do = torch.randn(b, nq, h, d, device=DEV, generator=g).to(torch.bfloat16)
    dq = torch.empty(b, nq, h, d, device=DEV, dtype=torch.bfloat16)
    dkv = torch.empty(b, nk, 2, h, d, device=DEV, dtype=torch.bfloat16)
    delta, dq_acc = _C.attention_bwd_workspace(b, h, nq, d, DEV)
    _C.attention_bwd(q, k, v, o, do, lse, dq, dkv[:, :, 0], dkv[:, :, 1], delta, dq_acc, scale)

    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = reference(qr, kr, vr, scale)
    o_ref.backward(do.float())
    assert rel_err(dq, qr.grad) < 6e-3
    assert rel_err(dkv[:, :, 0], kr.grad) < 6e-3
    assert rel_err(dkv[:, :, 1], vr.grad) < 6e-3


@pytest.mark.parametrize("b,h,nq,nk,d", [(2, 12, 685, 685, 64), (2, 16, 517, 300, 32), (1, 2, 130, 70, 64), (3, 4, 64, 129, 32)])
def test_attention_backward_fused_bias_column_sums(b, h, nq, nk, d):
    """dq / dk / dv column sums (the q / k / v projection bias gradients, cinema/vit.py:472-473) accumulated by the
    backward itself: equal to a separate column sum over the stored bf16 gradients, added on top of what the buffers held."""
    q, k, v = make_qkv(b, h, nq, nk, d, seed=3 + nq, fused=False)
    scale = d ** -0.5
    o = torch.empty(b, nq, h, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(b, h, nq, device=DEV)
    _C.attention_fwd(q, k, v, o, lse, scale)
    g = torch.Generator(device=DEV).manual_seed(2)
    do = torch.randn(b, nq, h, d, device=DEV, generator=g).to(torch.bfloat16)
    dq, dk, dv = torch.empty_like(q), torch.empty(b, nk, h, d, device=DEV, dtype=torch.bfloat16), torch.empty(b, nk, h, d, device=DEV, dtype=torch.bfloat16)
    delta, dq_acc = _C.attention_bwd_workspace(b, h, nq, d, DEV)
    base = [torch.randn(h * d, device=DEV, generator=g) for _ in range(3)]
    cs = [t.clone() for t in base]
    _C.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dq_acc, scale, cs[0], cs[1], cs[2])
    dq2, dk2, dv2 = torch.empty_like(dq), torch.empty_like(dk), torch.empty_like(dv)
    _C.attention_bwd(q, k, v, o, do, lse, dq2, dk2, dv2, delta, dq_acc, scale)
    assert torch.equal(dk, dk2) and torch.equal(dv, dv2)  # the side outputs do not change the main ones
    assert rel_err(dq, dq2.float()) < 1e-3                # (dQ: fp32 atomics reorder the sums between runs)
    for got, b0, t in zip(cs, base, (dq, dk, dv)):
        want = b0.double() + t.double().sum((0, 1)).reshape(-1)
        torch.testing.assert_close(got.double(), want, rtol=1e-4, atol=1e-3 * float(t.float().abs().max()) * (b * t.shape[1]) ** 0.5 + 1e-4)
