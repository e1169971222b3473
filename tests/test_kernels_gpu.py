"""GPU unit tests: every C-ABI kernel against a plain PyTorch reference of the same op.

Bit-exact for copies / index work; for floating point the tolerance is written per test:
inputs are bf16-representable so the only differences are fp32 accumulation order and the
final bf16 rounding of the output."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from cinema_b200 import _C

DEV = "cuda"


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf16_randn(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(torch.bfloat16)


# ----------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [
    (128, 256, 64), (128, 128, 128), (256, 512, 768), (1154, 768, 768), (1154, 3072, 768), (1154, 768, 3072),
    (300, 256, 512), (77, 64, 16), (50, 16, 16), (129, 24, 40), (2305, 512, 2048), (513, 1536, 768),
    # ViT-L (cinema/vit.py:807-822: dim 1024, mlp 4096, qkv 3072) at two-sample token counts (2 x 769 rows)
    (1538, 1024, 1024), (1538, 4096, 1024), (1538, 1024, 4096), (1538, 3072, 1024), (1538, 512, 1024),
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_gemm_forward_kmajor(m, n, k, block_n):
    a, b = bf16_randn(m, k, seed=1), bf16_randn(n, k, scale=0.05, seed=2)
    out = torch.empty(m, n, device=DEV, dtype=torch.float32)
    _C.gemm(a, b, out, block_n=block_n)
    ref = a.float() @ b.float().t()
    assert rel_err(out, ref) < 2e-5  # fp32 accumulate, different summation order only
    out16 = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    _C.gemm(a, b, out16, block_n=block_n)
    assert rel_err(out16, ref) < 3e-3  # one bf16 rounding of the output (2^-9 relative per element)


@pytest.mark.parametrize("m,n,k", [(1154, 768, 3072), (1154, 3072, 768), (300, 512, 256), (77, 16, 64), (2305, 512, 512),
                                   (1538, 1024, 4096), (1538, 4096, 1024), (1538, 1024, 3072)])
def test_gemm_dgrad_b_mn_major(m, n, k):
    # dX[m, n] = dY[m, k] @ W[k, n]  with W stored (k rows, n contiguous) == B MN-major
    dy, w = bf16_randn(m, k, seed=3), bf16_randn(k, n, scale=0.05, seed=4)
    out = torch.empty(m, n, device=DEV, dtype=torch.float32)
    _C.gemm(dy, w, out, b_mn=True)
    assert rel_err(out, dy.float() @ w.float()) < 2e-5


@pytest.mark.parametrize("t,n,k", [(1154, 768, 768), (1154, 3072, 768), (1154, 768, 3072), (9232, 768, 768),
                                   (300, 256, 512), (77, 16, 64), (130, 64, 16), (4610, 512, 2048),
                                   (1538, 1024, 4096), (1538, 4096, 1024), (12304, 3072, 1024)])
@pytest.mark.parametrize("split_k", [0, 1, 3])
def test_gemm_wgrad_both_mn_major_accumulate(t, n, k, split_k):
    # dW[n, k] += dY[t, n]^T @ X[t, k]
    dy, x = bf16_randn(t, n, scale=0.1, seed=5), bf16_randn(t, k, seed=6)
    base = torch.randn(n, k, device=DEV)
    out = base.clone()
    _C.gemm(dy, x, out, a_mn=True, b_mn=True, accumulate=True, split_k=split_k)
    ref = base.double() + dy.double().t() @ x.double()
    assert rel_err(out, ref) < 2e-5


def test_gemm_a_mn_b_k():
    # C[m, n] = A^T B^T with A stored (k, m), B stored (n, k)
    a, b = bf16_randn(192, 304, seed=7), bf16_randn(136, 192, seed=8)
    out = torch.empty(304, 136, device=DEV)
    _C.gemm(a, b, out, a_mn=True)
    assert rel_err(out, a.float().t() @ b.float().t()) < 2e-5


@pytest.mark.parametrize("m,n,k", [(1154, 3072, 768), (130, 64, 16), (2305, 2048, 512), (1538, 4096, 1024)])
def test_gemm_bias_gelu_epilogue(m, n, k):
    a, b = bf16_randn(m, k, seed=9), bf16_randn(n, k, scale=0.05, seed=10)
    bias = torch.randn(n, device=DEV) * 0.1
    pre = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    act = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    _C.gemm(a, b, pre, out2=act, bias=bias, epilogue=_C.EPI_GELU)
    ref_pre = (a.float() @ b.float().t() + bias)
    assert rel_err(pre, ref_pre) < 3e-3
    # activation is GELU(erf) of the *stored* bf16 pre-activation (what autocast feeds nn.GELU)
    ref_act = torch.nn.functional.gelu(pre.float())
    assert float((act.float() - ref_act).abs().max()) <= 2 ** -8 * float(ref_act.abs().max()) + 1e-6
    assert rel_err(act, ref_act) < 3e-3


def test_gemm_gelu_bwd_epilogue_and_shadow():
    m, n, k = 1154, 768, 3072  # dh = (g @ W2) * gelu'(pre): here M x "N=3072"... use (m, 3072, 768)
    m, n, k = 1154, 3072, 768
    g, w = bf16_randn(m, k, seed=11), bf16_randn(k, n, scale=0.05, seed=12)
    pre = bf16_randn(m, n, seed=13)
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    _C.gemm(g, w, out, b_mn=True, aux=pre, epilogue=_C.EPI_GELU_BWD)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(g.float() @ w.float())
    assert rel_err(out, x.grad) < 3e-3


def test_gemm_residual_fp32_with_bf16_shadow():
    m, n, k = 1154, 768, 3072
    a, b = bf16_randn(m, k, seed=14), bf16_randn(n, k, scale=0.02, seed=15)
    bias, res = torch.randn(n, device=DEV), torch.randn(m, n, device=DEV)
    out = torch.empty(m, n, device=DEV)
    shadow = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    _C.gemm(a, b, out, out2=shadow, bias=bias, residual=res)
    ref = a.float() @ b.float().t() + bias + res
    assert rel_err(out, ref) < 2e-5
    assert torch.equal(shadow, out.to(torch.bfloat16))
    # in-place residual (out aliases residual), the way the blocks use it
    x = res.clone()
    _C.gemm(a, b, x, bias=bias, residual=x)
    assert rel_err(x, ref) < 2e-5


def test_gemm_strided_views():
    # consume / produce column slices of wider buffers (fused qkv layout)
    m = 700
    buf = bf16_randn(m, 3 * 256, seed=16)
    w = bf16_randn(512, 256, scale=0.05, seed=17)
    big = torch.zeros(m, 1024, device=DEV, dtype=torch.bfloat16)
    _C.gemm(buf[:, 256:512], w, big[:, 512:])
    assert rel_err(big[:, 512:], buf[:, 256:512].float() @ w.float().t()) < 3e-3
    assert float(big[:, :512].abs().max()) == 0.0


def test_gemm_bad_args_raise():
    a, b = bf16_randn(64, 64), bf16_randn(60, 64)
    with pytest.raises(RuntimeError):
        _C.gemm(a, b, torch.empty(64, 60, device=DEV))  # N not a multiple of 8
    with pytest.raises(RuntimeError):
        _C.gemm(a.cpu(), b.cpu(), torch.empty(64, 60))


def test_colsum():
    x = bf16_randn(4099, 776, seed=18)
    out = torch.ones(776, device=DEV)
    _C.colsum(x, out)
    assert rel_err(out, 1 + x.double().sum(0)) < 1e-5


# ----------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("m,d", [(1154, 768), (4610, 512), (33, 16), (100, 64), (257, 128), (64, 1024), (16, 1280), (9, 2048)])
def test_layernorm_fwd_bwd(m, d):
    g = torch.Generator(device=DEV).manual_seed(d)
    x = torch.randn(m, d, device=DEV, generator=g) * 2 + 0.5
    gamma = torch.randn(d, device=DEV, generator=g)
    beta = torch.randn(d, device=DEV, generator=g)
    y16 = torch.empty(m, d, device=DEV, dtype=torch.bfloat16)
    y32 = torch.empty(m, d, device=DEV)
    mean, rstd = torch.empty(m, device=DEV), torch.empty(m, device=DEV)
    _C.layernorm_fwd(x, gamma, beta, 1e-5, y16, y32, mean, rstd)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5)
    torch.testing.assert_close(y32, ref, rtol=1e-5, atol=2e-5)
    assert torch.equal(y16, y32.to(torch.bfloat16))
    torch.testing.assert_close(mean, x.mean(1), rtol=1e-5, atol=1e-5)

    dy = torch.randn(m, d, device=DEV, generator=g)
    dres = torch.randn(m, d, device=DEV, generator=g)
    ref.backward(dy)
    dx32 = torch.empty(m, d, device=DEV)
    dx16 = torch.empty(m, d, device=DEV, dtype=torch.bfloat16)
    dgamma, dbeta = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    _C.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx32, dx16, dgamma, dbeta)
    torch.testing.assert_close(dx32, xr.grad + dres, rtol=1e-4, atol=1e-4)
    assert torch.equal(dx16, dx32.to(torch.bfloat16))
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dbeta, br.grad, rtol=1e-4, atol=1e-3)
    # bf16 upstream gradient, no residual, no parameter grads
    dy16 = dy.to(torch.bfloat16)
    _C.layernorm_bwd(dy16, x, mean, rstd, gamma, None, dx32, None, None, None)
    xr.grad = None
    torch.nn.functional.layer_norm(xr, (d,), gamma, beta, 1e-5).backward(dy16.float())
    torch.testing.assert_close(dx32, xr.grad, rtol=1e-4, atol=1e-4)
    if d in (16, 32, 64, 128) or (d % 8 == 0 and 256 <= d <= 1024):
        # in-place residual (dx32 aliases dres) with the fused column sums of the bf16 output (bias gradient of the
        # consuming Linear): must equal a separate colsum over dx16
        acc = dres.clone()
        dxsum = torch.zeros(d, device=DEV)
        _C.layernorm_bwd(dy16, x, mean, rstd, gamma, dres=acc, dx32=acc, dx16=dx16, dxsum=dxsum)
        torch.testing.assert_close(acc, xr.grad + dres, rtol=1e-4, atol=1e-4)
        assert torch.equal(dx16, acc.to(torch.bfloat16))
        torch.testing.assert_close(dxsum, dx16.float().sum(0), rtol=1e-4, atol=2e-3)


# ----------------------------------------------------------------------------- data movement
def test_cast_matches_torch_rounding():
    x = torch.randn(1_000_003, device=DEV)
    y = torch.empty(1_000_003, device=DEV, dtype=torch.bfloat16)
    _C.cast_bf16(x, y)
    assert torch.equal(y, x.to(torch.bfloat16))


@pytest.mark.parametrize("b,n,ratio", [(16, 2304, 0.75), (3, 256, 0.75), (2, 37, 0.6), (4, 16, 0.0), (5, 100, 1.0)])
def test_mask_to_index_bit_exact(b, n, ratio):
    n_keep = int(n * (1 - ratio))
    g = torch.Generator(device=DEV).manual_seed(n)
    rank = torch.argsort(torch.argsort(torch.rand(b, n, device=DEV, generator=g), dim=1), dim=1)
    mask = rank >= n_keep
    keep, drop, slot = _C.mask_to_index(mask, n_keep)
    ar = torch.arange(n, device=DEV).expand(b, n)
    assert torch.equal(keep.long(), ar[~mask].reshape(b, n_keep))
    assert torch.equal(drop.long(), ar[mask].reshape(b, n - n_keep))
    if n_keep:
        assert torch.equal(torch.gather(slot, 1, keep.long()), torch.arange(n_keep, device=DEV, dtype=torch.int32).expand(b, -1))
    if n - n_keep:
        assert torch.equal(torch.gather(slot, 1, drop.long()),
                           torch.arange(n - n_keep, device=DEV, dtype=torch.int32).expand(b, -1))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gather_scatter_rows_bit_exact(dtype):
    b, n, d, n_keep = 5, 2304, 768, 576
    x = torch.randn(b, n, d, device=DEV).to(dtype)
    mask = torch.argsort(torch.argsort(torch.rand(b, n, device=DEV), dim=1), dim=1) >= n_keep
    keep, drop, _ = _C.mask_to_index(mask, n_keep)
    out = torch.zeros(b, 1 + n_keep, d, device=DEV, dtype=dtype)
    _C.gather_rows(x, keep, out, out_off=1)
    assert torch.equal(out[:, 1:], x[~mask].reshape(b, n_keep, d))  # cinema/mae/mae.py:550
    assert float(out[:, 0].abs().max()) == 0
    # broadcast table (pos-embed): cinema/mae/mae.py:97-99
    table = torch.randn(1, n, d, device=DEV).to(dtype)
    out2 = torch.empty(b, n - n_keep, d, device=DEV, dtype=dtype)
    _C.gather_rows(table, drop, out2)
    assert torch.equal(out2, table.expand(b, -1, -1)[mask].reshape(b, n - n_keep, d))
    # scatter is the exact inverse on the selected rows
    back = torch.zeros_like(x)
    _C.scatter_rows(out, keep, back, src_off=1)
    assert torch.equal(back[~mask], x[~mask])
    assert float(back[mask].abs().max()) == 0


def test_embed_rows():
    b, n, d, k = 3, 144, 512, 36
    table = torch.randn(n, d, device=DEV)
    a = torch.randn(b, 1 + k, d, device=DEV)
    row = torch.randn(d, device=DEV)
    idx = torch.stack([torch.randperm(n, device=DEV)[:k].sort().values for _ in range(b)]).int()
    out = torch.zeros(b, 5 + k, d, device=DEV)
    _C.embed_rows(a, 1, None, table, idx, b, k, out=out, out_off=5)
    torch.testing.assert_close(out[:, 5:], a[:, 1:] + table[idx.long()], rtol=0, atol=0)
    out16 = torch.zeros(b, 5 + k, d, device=DEV, dtype=torch.bfloat16)
    _C.embed_rows(None, 0, row, table, idx, b, k, out=out, out16=out16, out_off=5)
    torch.testing.assert_close(out[:, 5:], row + table[idx.long()], rtol=0, atol=0)
    assert torch.equal(out16[:, 5:], (row + table[idx.long()]).to(torch.bfloat16))
    # table-less form: plain strided row copy (cls rows)
    _C.embed_rows(a, 0, None, None, None, b, 1, out=out, out_off=2)
    assert torch.equal(out[:, 2], a[:, 0])


def test_colsum_seg_and_scale_cast():
    b, n, d = 5, 77, 512
    x = torch.randn(b, n, d, device=DEV)
    out = torch.ones(d, device=DEV)
    _C.colsum_seg(x, 3, 40, out)
    torch.testing.assert_close(out, 1 + x[:, 3:43].double().sum(dim=(0, 1)).float(), rtol=1e-5, atol=1e-4)
    out.zero_()
    _C.colsum_seg(x, 0, 1, out)
    torch.testing.assert_close(out, x[:, 0].sum(0), rtol=1e-5, atol=1e-5)
    src = torch.randn(1003, device=DEV)
    dst = torch.empty(1003, device=DEV, dtype=torch.bfloat16)
    sc = torch.tensor([0.25], device=DEV)
    _C.scale_cast(src, dst, sc, 2.0)
    assert torch.equal(dst, (src * 0.5).to(torch.bfloat16))


def test_mae_loss_finalize():
    acc = torch.tensor([[8.0, 6.0, 3.0, 1.5, 2.5, 0, 0, 0], [float("nan"), 1, 1, 0, 0, 0, 0, 0],
                        [2.0, 4.0, 2.0, -1.0, -2.0, 0, 0, 0]], device=DEV)
    out = torch.zeros(16, device=DEV)
    scales = torch.zeros(3, device=DEV)
    _C.mae_loss_finalize(acc, [4.0, 2.0, 8.0], [2.0, 1.0, 4.0], out, scales)
    o = out.cpu()
    assert o[1] == 2.0 and o[2] == 3.0 and o[3] == 1.5 and o[4] == 1.5 and o[5] == 2.5
    assert torch.isnan(o[6]) and o[11] == 0.25 and o[12] == 1.0
    assert o[0] == (2.0 + 0.25) / 2  # the non-finite view is dropped
    assert scales.cpu().tolist() == [2.0 / (4.0 * 2), 0.0, 2.0 / (8.0 * 2)]


def _torch_patchify(image, patch):
    n = len(patch)
    b, c, *sp = image.shape
    grid = [s // p for s, p in zip(sp, patch)]
    split = []
    for g_, p_ in zip(grid, patch):
        split += [g_, p_]
    x = image.reshape(b, c, *split)
    x = x.permute(0, *[2 + 2 * i for i in range(n)], *[3 + 2 * i for i in range(n)], 1).contiguous()
    return x.reshape(b, math.prod(grid), math.prod(patch) * c)


def _torch_unpatchify(tokens, patch, grid, c):
    """inverse of _torch_patchify: (B, prod(grid), prod(patch) * C) channel fastest -> (B, C, *spatial)."""
    n = len(patch)
    b = tokens.shape[0]
    x = tokens.reshape(b, *grid, *patch, c)
    perm = [0, 2 * n + 1]
    for i in range(n):
        perm += [1 + i, 1 + n + i]
    x = x.permute(*perm).contiguous()
    return x.reshape(b, c, *[g_ * p_ for g_, p_ in zip(grid, patch)])


@pytest.mark.parametrize("shape,patch", [((2, 1, 192, 192, 16), (16, 16, 1)), ((3, 1, 256, 256), (16, 16)),
                                         ((2, 3, 8, 12), (2, 4)), ((1, 2, 4, 4, 2, 6), (2, 2, 1, 3)),
                                         ((2, 128, 24, 24, 16), (2, 2, 1))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patchify_unpatchify_bit_exact(shape, patch, dtype):
    img = torch.randn(*shape, device=DEV).to(dtype)
    ref = _torch_patchify(img, patch)
    tok = torch.empty_like(ref)
    _C.patchify(img, tok, shape[0], shape[1], shape[2:], patch, inverse=False)
    assert torch.equal(tok, ref)
    back = torch.empty_like(img)
    _C.patchify(tok, back, shape[0], shape[1], shape[2:], patch, inverse=True)
    assert torch.equal(back, img)


@pytest.mark.parametrize("chan_last_mem", [False, True])
@pytest.mark.parametrize("chan_last_order", [True, False])
def test_gather_scatter_patches(chan_last_mem, chan_last_order):
    b, c, sp, patch = 2, 32, (24, 24, 4), (2, 2, 1)
    grid = tuple(s // p for s, p in zip(sp, patch))
    n = math.prod(grid)
    x = torch.randn(b, c, *sp, device=DEV).to(torch.bfloat16)
    if chan_last_mem:
        x = x.to(memory_format=torch.channels_last_3d)
    n_keep = n // 4
    mask = torch.argsort(torch.argsort(torch.rand(b, n, device=DEV), dim=1), dim=1) >= n_keep
    keep, _, _ = _C.mask_to_index(mask, n_keep)
    e = math.prod(patch) * c
    out = torch.empty(b * n_keep, e, device=DEV, dtype=torch.bfloat16)
    _C.gather_patches(x, grid, patch, keep, chan_last_order, out)
    tok = _torch_patchify(x.contiguous(), patch)  # (b, n, p*q*r*c) channel fastest
    if not chan_last_order:
        tok = tok.reshape(b, n, math.prod(patch), c).transpose(2, 3).reshape(b, n, e)
    ref = tok[~mask].reshape(b * n_keep, e)
    assert torch.equal(out, ref)
    # all tokens when idx is None
    out_all = torch.empty(b * n, e, device=DEV, dtype=torch.bfloat16)
    _C.gather_patches(x, grid, patch, None, chan_last_order, out_all)
    assert torch.equal(out_all, tok.reshape(b * n, e))
    # scatter back: visible patches restored, the rest untouched (zero)
    dst = torch.zeros_like(x)
    _C.scatter_patches(out, dst, grid, patch, keep, chan_last_order)
    vis = (~mask).reshape(b, 1, *grid)
    for a, p in enumerate(patch):
        vis = vis.repeat_interleave(p, dim=2 + a)
    assert torch.equal(dst, x * vis)


@pytest.mark.parametrize("t,c,f,grid,patch,chan_last", [
    (576, 128, (2, 2), (1, 1), (2, 2), True), (576, 128, (2, 2), (1, 1), (2, 2), False),
    (576, 64, (4, 4), (1, 1), (4, 4), False), (576, 64, (4, 4), (2, 2), (2, 2), False), (576, 64, (4, 4), (2, 2), (2, 2), True),
    (1000, 128, (2, 2, 1), (1, 1, 1), (2, 2, 1), True), (1000, 64, (4, 4, 1), (2, 2, 1), (2, 2, 1), False),
    (37, 64, (4, 4, 1), (1, 1, 1), (4, 4, 1), False), (5, 32, (2, 2), (1, 1), (2, 2), False),
])
def test_token_major_patch_rows_bulk_copy_path(t, c, f, grid, patch, chan_last):
    """The stem's token-major sources ((T, C, *f) fp32 views of [T][positions][C] memory, channel stride 1) take the
    bulk-copy staged kernels (cp.async.bulk into shared memory, permutation in shared memory, bulk reduce-add back):
    bit-exact against a plain torch restatement of gather_patches / scatter_patches, overwrite and accumulate, bf16 and
    fp32 rows.  The last case (C = 32: block of 128 elements) falls back to the generic kernel and must agree too."""
    g = torch.Generator(device=DEV).manual_seed(t + c)
    mem = torch.randn(t, math.prod(f), c, device=DEV, generator=g)                 # [T][positions][C]
    src = mem.view(t, *f, c).permute(0, len(f) + 1, *range(1, len(f) + 1))         # (T, C, *f), channel stride 1
    assert src.stride(1) == 1
    r, e = math.prod(grid), math.prod(patch) * c
    out = torch.empty(t * r, e, device=DEV, dtype=torch.bfloat16)
    _C.gather_patches(src, grid, patch, None, chan_last, out)
    tok = _torch_patchify(src.contiguous(), patch)                                  # (T, R, prod(patch) * C) channel fastest
    if not chan_last:
        tok = tok.reshape(t, r, math.prod(patch), c).transpose(2, 3)
    ref = tok.reshape(t * r, e)
    assert torch.equal(out, ref.to(torch.bfloat16))
    for rows in (out, ref.contiguous()):                                            # bf16 rows and fp32 rows
        for accumulate in (False, True):
            base = torch.randn(t, math.prod(f), c, device=DEV, generator=g)
            dmem = base.clone()
            dst = dmem.view(t, *f, c).permute(0, len(f) + 1, *range(1, len(f) + 1))
            _C.scatter_patches(rows, dst, grid, patch, None, chan_last, accumulate=accumulate)
            want = rows.float().reshape(t, r, -1)
            if not chan_last:
                want = want.reshape(t, r, c, math.prod(patch)).transpose(2, 3).reshape(t, r, -1)
            img = _torch_unpatchify(want, patch, grid, c)                           # (T, C, *f)
            want_mem = img.permute(0, *range(2, len(f) + 2), 1).reshape(t, math.prod(f), c)
            assert torch.equal(dmem, base + want_mem if accumulate else want_mem)


# ----------------------------------------------------------------------------- loss
@pytest.mark.parametrize("shape,patch", [((2, 1, 64, 64, 4), (16, 16, 1)), ((3, 1, 64, 64), (16, 16)), ((2, 2, 8, 8), (2, 4))])
@pytest.mark.parametrize("norm_target", [False, True])
def test_masked_mse(shape, patch, norm_target):
    b, c, *sp = shape
    img = torch.rand(*shape, device=DEV)
    tgt = _torch_patchify(img, patch)
    n, e = tgt.shape[1], tgt.shape[2]
    n_keep = n // 4
    mask = torch.argsort(torch.argsort(torch.rand(b, n, device=DEV), dim=1), dim=1) >= n_keep
    _, drop, slot = _C.mask_to_index(mask, n_keep)
    pred = torch.randn(b, n - n_keep, e, device=DEV)
    acc = torch.zeros(8, device=DEV)
    acc[3:5] = -float("inf")
    diff = torch.empty_like(pred)
    _C.masked_mse_fwd(img, patch, mask, slot, pred, norm_target, 1e-6, acc, diff)
    mean = tgt.mean(-1, keepdim=True)
    std = tgt.var(-1, keepdim=True) ** 0.5
    t = (tgt - mean) / (std + 1e-6) if norm_target else tgt
    t = t[mask].reshape(pred.shape)
    torch.testing.assert_close(diff, pred - t, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(acc[0] / pred.numel(), ((pred - t) ** 2).mean(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(acc[1] / (b * n), mean.mean(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(acc[2] / (b * n), std.mean(), rtol=1e-5, atol=1e-7)
    if norm_target:
        torch.testing.assert_close(acc[3], t.max(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(acc[4], pred.max(), rtol=0, atol=0)


# ----------------------------------------------------------------------------- native stem kernels
def _token_setup(b, grid_tok, keep_frac, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    n = math.prod(grid_tok)
    nk = max(1, int(n * keep_frac))
    mask = torch.ones(b, n, dtype=torch.bool, device=DEV)
    for i in range(b):
        mask[i, torch.randperm(n, device=DEV, generator=g)[:nk]] = False
    keep, drop, slot = _C.mask_to_index(mask, nk)
    return mask, keep, slot, nk


def _dense_from_tokens(x, idx, level, c):
    b = idx.shape[0]
    dense = torch.zeros(b, math.prod(level), c, device=DEV)
    dense.scatter_(1, idx.long()[..., None].expand(-1, -1, c), x.float().reshape(b, -1, c))
    return dense.transpose(1, 2).reshape(b, c, *level)


@pytest.mark.parametrize("grid_tok,f,c,keep_frac", [((12, 12, 16), (4, 4, 1), 64, 0.25), ((12, 12, 16), (2, 2, 1), 128, 0.25),
                                                      ((12, 12), (4, 4), 64, 0.25), ((6, 5), (2, 2), 16, 0.5),
                                                      ((3, 4, 2), (4, 4, 1), 8, 1.0), ((12, 12), (2, 2), 128, 1.0)])
def test_dwconv_tokens_fwd_bwd_against_dense_conv(grid_tok, f, c, keep_frac):
    import torch.nn.functional as F

    b = 2
    nd = len(f)
    mask, keep, slot, nk = _token_setup(b, grid_tok, keep_frac, 0)
    p = math.prod(f)
    level = [gt * ff for gt, ff in zip(grid_tok, f)]
    idx = _C.expand_token_index(keep, grid_tok, f)
    # index expansion against plain arithmetic
    t = keep.long()
    coords = []
    for a in range(nd - 1, -1, -1):
        coords.insert(0, t % grid_tok[a])
        t = t // grid_tok[a]
    pid = torch.arange(p, device=DEV)
    pa = []
    for a in range(nd - 1, -1, -1):
        pa.insert(0, pid % f[a])
        pid = pid // f[a]
    ref_idx = torch.zeros(b, nk, p, dtype=torch.long, device=DEV)
    for a in range(nd):
        ref_idx = ref_idx * level[a] + coords[a][..., None] * f[a] + pa[a]
    assert torch.equal(idx.long(), ref_idx.reshape(b, -1))

    x = bf16_randn(b * nk * p, c, seed=1)
    w = bf16_randn(c, 1, *([5] * nd), scale=0.2, seed=2)
    bias = torch.randn(c, device=DEV)
    out = torch.empty_like(x)
    _C.dwconv_tokens(x, out, w, bias, mask, slot, keep, grid_tok, f)
    conv = F.conv2d if nd == 2 else F.conv3d
    dense = _dense_from_tokens(x, idx, level, c)
    y = conv(dense, w.float(), bias, padding=2, groups=c)
    y = torch.gather(y.reshape(b, c, -1).transpose(1, 2), 1, idx.long()[..., None].expand(-1, -1, c)).reshape(-1, c)
    assert rel_err(out, y) < 4e-3  # bf16 output rounding
    # transpose = gradient w.r.t. the input of the same conv restricted to visible rows
    dy = bf16_randn(b * nk * p, c, seed=3)
    dxk = torch.empty_like(dy)
    _C.dwconv_tokens(dy, dxk, w, None, mask, slot, keep, grid_tok, f, transpose=True)
    dense_in = dense.clone().requires_grad_()
    wf = w.float().requires_grad_()
    yy = conv(dense_in, wf, None, padding=2, groups=c)
    ddense = _dense_from_tokens(dy, idx, level, c)
    (yy * ddense).sum().backward()
    dx_ref = torch.gather(dense_in.grad.reshape(b, c, -1).transpose(1, 2), 1,
                          idx.long()[..., None].expand(-1, -1, c)).reshape(-1, c)
    assert rel_err(dxk, dx_ref) < 4e-3
    dw = torch.zeros(c, 1, *([5] * nd), device=DEV)
    db = torch.zeros(c, device=DEV)
    _C.dwconv_tokens_wgrad(x, dy, dw, db, mask, slot, keep, grid_tok, f)
    assert rel_err(dw, wf.grad) < 1e-4
    assert rel_err(db, dy.float().sum(0)) < 1e-4


@pytest.mark.parametrize("m,d", [(300, 64), (77, 128), (33, 16)])
def test_layernorm_gelu_fwd_bwd(m, d):
    x = torch.randn(m, d, device=DEV) * 2 + 0.3
    gamma, beta = torch.rand(d, device=DEV) + 0.5, torch.randn(d, device=DEV) * 0.2
    y32 = torch.empty(m, d, device=DEV)
    mean, rstd = torch.empty(m, device=DEV), torch.empty(m, device=DEV)
    _C.layernorm_fwd(x, gamma, beta, 1e-6, y32=y32, mean=mean, rstd=rstd, act=True)
    xr = x.clone().requires_grad_()
    gr, br = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-6))
    torch.testing.assert_close(y32, ref, rtol=1e-5, atol=2e-6)
    dy = torch.randn(m, d, device=DEV)
    (ref * dy).sum().backward()
    dx = torch.empty(m, d, device=DEV)
    dg, dbt = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    _C.layernorm_bwd(dy, x, mean, rstd, gamma, dx32=dx, dgamma=dg, dbeta=dbt, beta_act=beta)
    torch.testing.assert_close(dx, xr.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dg, gr.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbt, br.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("m,n,k", [(1000, 512, 256), (333, 72, 64), (4200, 3072, 768)])
def test_gemm_epilogue_column_sums(m, n, k):
    """colsum: the bias gradient of the consumer layer, fused into the dgrad epilogue (plain and GELU' flavours); it must
    equal a separate column sum over the stored bf16 output."""
    dy = bf16_randn(m, k, seed=1)
    w = bf16_randn(k, n, scale=0.05, seed=2)  # [K, N]: read MN-major like a dgrad
    aux = bf16_randn(m, n, seed=3)
    for use_aux in (False, True):
        out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        cs = torch.zeros(n, device=DEV)
        if use_aux:
            _C.gemm(dy, w, out, b_mn=True, aux=aux, epilogue=_C.EPI_GELU_BWD, colsum=cs)
        else:
            _C.gemm(dy, w, out, b_mn=True, colsum=cs)
        ref = out.float().sum(0)
        torch.testing.assert_close(cs, ref, rtol=2e-4, atol=2e-3 * float(out.float().abs().max()) + 1e-3)
        out2 = torch.empty_like(out)
        if use_aux:
            _C.gemm(dy, w, out2, b_mn=True, aux=aux, epilogue=_C.EPI_GELU_BWD)
        else:
            _C.gemm(dy, w, out2, b_mn=True)
        assert torch.equal(out, out2)  # the side output does not change the main one


@pytest.mark.parametrize("m,n,k,rows", [(6 * 197, 768, 768, 197), (4 * 2305, 512, 256, 2305), (3 * 50, 64, 128, 50)])
def test_gemm_row_scale_epilogue(m, n, k, rows):
    """Stochastic depth fused into the branch's last Linear: out = residual + s[row // rows] * (A W^T + b), fp32 and bf16
    outputs; s = 1 reproduces the plain epilogue bit for bit."""
    a = bf16_randn(m, k, seed=4)
    w = bf16_randn(n, k, scale=0.05, seed=5)
    bias = torch.randn(n, device=DEV)
    res = torch.randn(m, n, device=DEV)
    groups = m // rows
    s = torch.tensor([0.0, 1.0 / 0.6] * groups, device=DEV)[:groups].contiguous()
    ref = (a.float() @ w.float().t() + bias) * s.repeat_interleave(rows)[:, None] + res
    out = torch.empty(m, n, device=DEV)
    _C.gemm(a, w, out, bias=bias, residual=res, row_scale=s, rows_per_group=rows)
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)
    dropped = (s == 0).repeat_interleave(rows)
    assert torch.equal(out[dropped], res[dropped])  # a dropped sample passes the residual through untouched
    out16 = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    _C.gemm(a, w, out16, bias=bias, residual=res, row_scale=s, rows_per_group=rows)
    torch.testing.assert_close(out16.float(), ref, rtol=1e-2, atol=2e-2)
    ones = torch.ones(groups, device=DEV)
    o1, o2 = torch.empty(m, n, device=DEV), torch.empty(m, n, device=DEV)
    _C.gemm(a, w, o1, bias=bias, residual=res, row_scale=ones, rows_per_group=rows)
    _C.gemm(a, w, o2, bias=bias, residual=res)
    assert torch.equal(o1, o2)


def test_scale_cast_grouped():
    """bf16(src * s[i // group]): the per-sample branch gradient of stochastic depth; scalar mode unchanged."""
    b, n, d = 5, 37, 64
    src = torch.randn(b * n, d, device=DEV)
    s = torch.tensor([0.0, 2.0, 1.25, 0.0, 1.0], device=DEV)
    dst = torch.empty(b * n, d, device=DEV, dtype=torch.bfloat16)
    _C.scale_cast(src, dst, s, group=n * d)
    ref = (src.view(b, -1) * s[:, None]).view(b * n, d).to(torch.bfloat16)
    assert torch.equal(dst, ref)
    one = torch.tensor([0.5], device=DEV)
    _C.scale_cast(src, dst, one, scale=3.0)
    assert torch.equal(dst, (src * 1.5).to(torch.bfloat16))
    with pytest.raises((RuntimeError, AssertionError)):
        _C.scale_cast(src, dst, s[:4].contiguous(), group=6)  # group must be a multiple of 4 dividing n


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int16, torch.uint16, torch.float32])
def test_scale_intensity_kernel(dtype):
    """cb_scale_intensity == MONAI ScaleIntensity(0, 1) per sample ((x - min) / (max - min), constant sample -> 0),
    bit-exact against the same fp32 arithmetic in torch (cinema/mae/pretrain.py:184)."""
    from cinema_b200.data import scale_intensity

    g = torch.Generator().manual_seed(0)
    b, shape = 5, (1, 24, 20, 6)
    if dtype == torch.float32:
        raw = torch.rand(b, *shape, generator=g) * 3000 - 1000
    elif dtype == torch.uint16:
        raw = torch.randint(0, 60000, (b, *shape), generator=g, dtype=torch.int32).to(torch.uint16)
    else:
        lim = 255 if dtype == torch.uint8 else 3000
        raw = torch.randint(0 if dtype == torch.uint8 else -500, lim, (b, *shape), generator=g, dtype=torch.int32).to(dtype)
    raw[2] = 7  # constant sample
    flat = raw.reshape(b, -1).to(torch.float32)
    lo, hi = flat.min(1).values.contiguous(), flat.max(1).values.contiguous()
    out = torch.empty(raw.shape, device=DEV, dtype=torch.float32)
    _C.scale_intensity(raw.to(DEV), lo.to(DEV), hi.to(DEV), out)
    ref = scale_intensity(raw.to(torch.float32) if dtype == torch.uint16 else raw, lo, hi)
    assert torch.equal(out.cpu(), ref)
    assert float(out[2].abs().max()) == 0.0 and float(out.max()) == 1.0 and float(out.min()) == 0.0


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int16, torch.float32])
@pytest.mark.parametrize("nd", [3, 2])
def test_zoom_intensity_kernel(dtype, nd):
    """cb_zoom_intensity (RandZoom keep_size -> ScaleIntensity -> SpatialPad "end", cinema/mae/pretrain.py:163-199) against
    MONAI's Zoom algorithm restated over torch's own interpolate on the same device (data.zoom_intensity): trilinear for 3-D
    frames, bicubic for 2-D; zoom factors below / above 1, the identity (bit-exact), frames smaller than the padded sample,
    a constant frame.  fp32 tolerance: the two resamplers differ in FMA contraction only; ScaleIntensity divides by the range."""
    from cinema_b200.data import scale_intensity, zoom_intensity

    g = torch.Generator().manual_seed(3)
    size = (40, 36, 10) if nd == 3 else (48, 44)
    ext = [[40, 36, 10], [33, 36, 7], [40, 29, 10], [17, 20, 5], [40, 36, 10], [30, 30, 8]] if nd == 3 else \
          [[48, 44, 1], [41, 44, 1], [48, 37, 1], [20, 23, 1], [48, 44, 1], [30, 30, 1]]
    b = len(ext)
    raw = torch.zeros(b, 1, *size)
    for j, e in enumerate(ext):
        box = tuple(slice(0, a) for a in e[:nd])
        raw[j, 0][box] = torch.rand(tuple(e[:nd]), generator=g) * (250 if dtype == torch.uint8 else 2000) + 3
    raw[5] = 0
    raw[5, 0][tuple(slice(0, a) for a in ext[5][:nd])] = 9  # a constant frame
    raw = raw.floor().to(dtype) if dtype != torch.float32 else raw
    zoom = torch.tensor([1.0, 0.9, 1.1, 0.937, 1.063, 1.0], dtype=torch.float32)
    extent = torch.tensor(ext, dtype=torch.int32)
    out = torch.full(raw.shape, -7.0, device=DEV, dtype=torch.float32)
    _C.zoom_intensity(raw.to(DEV), extent.to(DEV), zoom.to(DEV), out)
    ref = zoom_intensity(raw.to(DEV), extent, zoom)  # torch's CUDA interpolate
    assert float((out - ref).abs().max()) < 2e-5
    # identity: exactly ScaleIntensity of the frame
    f0 = raw[0:1].to(torch.float32)
    lo, hi = f0.min().reshape(1), f0.max().reshape(1)
    assert torch.equal(out[0:1].cpu(), scale_intensity(f0, lo, hi))
    for j, e in enumerate(ext):
        box = tuple(slice(0, a) for a in e[:nd])
        inside = out[j, 0][box]
        assert float(out[j].sum()) == pytest.approx(float(inside.sum()), rel=1e-6)  # zeros outside the frame
        if j != 5:
            assert float(inside.min()) == 0.0 and float(inside.max()) == pytest.approx(1.0, abs=1e-6)
    assert float(out[5].abs().max()) == 0.0  # a constant frame (not zoomed) maps to 0


@pytest.mark.parametrize("shape", [(2, 4, 48, 40, 6), (3, 4, 64, 56), (1, 2, 33, 17), (2, 8, 20, 20, 4)])
@pytest.mark.parametrize("logit_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("label_dtype", [torch.int64, torch.int16, torch.uint8])
def test_seg_loss_kernels(shape, logit_dtype, label_dtype):
    """cb_seg_loss_fwd / _bwd (cross-entropy with ignore_index -1 + foreground soft Dice, cinema/segmentation/train.py:77-103)
    against the torch restatement on the same device: loss, both metrics and d loss / d logits (scaled by an upstream
    gradient); unlabelled voxels with signed label types, every class count up to 8."""
    from cinema_b200.segmentation.loss import segmentation_loss, segmentation_loss_restated

    g = torch.Generator().manual_seed(11)
    logits = (torch.randn(shape, generator=g) * 2.5).to(logit_dtype).to(DEV)
    labels = torch.randint(0, shape[1], (shape[0], 1, *shape[2:]), generator=g)
    if label_dtype != torch.uint8:
        labels[torch.rand(labels.shape, generator=g) < 0.15] = -1
    labels = labels.to(label_dtype).to(DEV)
    x = logits.clone().requires_grad_(True)
    loss, metrics = segmentation_loss(x, labels)
    (loss * 1.7).backward()
    ref_x = logits.float().clone().requires_grad_(True)
    ref_loss, ref_ce, ref_dice = segmentation_loss_restated(ref_x, labels)
    (ref_loss * 1.7).backward()
    assert abs(float(loss) - float(ref_loss)) < 2e-5 * abs(float(ref_loss))
    assert abs(float(metrics["cross_entropy"]) - float(ref_ce)) < 2e-5 * abs(float(ref_ce))
    assert abs(float(metrics["mean_dice_loss"]) - float(ref_dice)) < 2e-5
    tol = 1e-5 if logit_dtype == torch.float32 else 6e-3  # the gradient is stored in the logits' dtype
    assert rel_err(x.grad.float(), ref_x.grad) < tol
    assert x.grad.dtype == logit_dtype
