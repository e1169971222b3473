"""Workload for compute-sanitizer (tools/gpu_sanitize.sh): every kernel family of the library once, on reduced shapes
that still cover the ragged tiles (token counts that are not multiples of 128, both head dims, every GEMM tile width /
CTA-pair mode / epilogue flavour), plus one small 4-view MAE step (forward + backward + AdamW) through the modules."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cinema_b200 import CineMA, _C  # noqa: E402
from cinema_b200.train import MAETrainer  # noqa: E402

DEV, BF = "cuda", torch.bfloat16


def rn(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


def gemms():
    for m, n, k in ((300, 256, 128), (517, 512, 192), (130, 64, 64)):
        a, b = rn(m, k, dtype=BF), rn(n, k, dtype=BF, scale=0.05)
        for bn in (0, 64, 128, 256):
            _C.gemm(a, b, torch.empty(m, n, device=DEV, dtype=BF), block_n=bn)
        bias, res = rn(n), rn(m, n)
        _C.gemm(a, b, torch.empty(m, n, device=DEV, dtype=BF), bias=bias, block_n=256)  # wide-store epilogue when paired
        _C.gemm(a, b, res, out2=torch.empty(m, n, device=DEV, dtype=BF), bias=bias, residual=res)
        pre, act = torch.empty(m, n, device=DEV, dtype=BF), torch.empty(m, n, device=DEV, dtype=BF)
        _C.gemm(a, b, pre, out2=act, bias=bias, epilogue=_C.EPI_GELU)
        dy, w = rn(m, n, dtype=BF), rn(n, k, dtype=BF, scale=0.05)
        cs = torch.zeros(k, device=DEV)
        _C.gemm(dy, w, torch.empty(m, k, device=DEV, dtype=BF), b_mn=True, colsum=cs)
        _C.gemm(rn(m, k, dtype=BF), rn(k, n, dtype=BF), torch.empty(m, n, device=DEV, dtype=BF), b_mn=True, aux=pre,
                epilogue=_C.EPI_GELU_BWD)
        for sk in (0, 3):
            _C.gemm(dy, a, torch.zeros(n, k, device=DEV), a_mn=True, b_mn=True, accumulate=True, split_k=sk)


def attention():
    for b, nq, nk, h, d in ((2, 301, 301, 3, 64), (2, 389, 140, 4, 32), (1, 64, 17, 2, 32), (1, 129, 257, 2, 64)):
        q, k, v = (rn(b, n, h, d, dtype=BF) for n in (nq, nk, nk))
        o, lse = torch.empty(b, nq, h, d, device=DEV, dtype=BF), torch.empty(b, h, nq, device=DEV)
        _C.attention_fwd(q, k, v, o, lse, d ** -0.5)
        do = rn(b, nq, h, d, dtype=BF)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta, dqa = _C.attention_bwd_workspace(b, h, nq, d, DEV)
        _C.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dqa, d ** -0.5)


def layernorm():
    for m, d in ((301, 768), (389, 512), (1000, 64), (77, 128), (50, 1024)):
        x, g, b = rn(m, d), rn(d), rn(d)
        y16 = torch.empty(m, d, device=DEV, dtype=BF)
        mean, rstd = torch.empty(m, device=DEV), torch.empty(m, device=DEV)
        _C.layernorm_fwd(x, g, b, 1e-5, y16=y16, mean=mean, rstd=rstd)
        dres = rn(m, d)
        _C.layernorm_bwd(rn(m, d, dtype=BF), x, mean, rstd, g, dres=dres, dx32=dres, dx16=torch.empty(m, d, device=DEV, dtype=BF),
                         dgamma=torch.zeros(d, device=DEV), dbeta=torch.zeros(d, device=DEV))


def input_pipeline():
    for shape, ext in (((3, 1, 24, 20, 6), [[24, 20, 6], [17, 20, 5], [24, 13, 6]]), ((3, 1, 40, 36), [[40, 36, 1], [33, 36, 1], [40, 29, 1]])):
        raw = torch.randint(0, 900, shape, device=DEV, dtype=torch.int16)
        extent = torch.tensor(ext, dtype=torch.int32, device=DEV)
        zoom = torch.tensor([1.0, 0.93, 1.08], device=DEV)
        _C.zoom_intensity(raw, extent, zoom, torch.empty(shape, device=DEV))
        flat = raw.reshape(3, -1).float()
        _C.scale_intensity(raw, flat.min(1).values.contiguous(), flat.max(1).values.contiguous(), torch.empty(shape, device=DEV))
    for dt in (torch.float32, BF):  # segmentation loss, forward + backward
        logits = rn(2, 4, 24, 20, 6, dtype=dt)
        labels = torch.randint(-1, 4, (2, 1, 24, 20, 6), device=DEV)
        out, coef = _C.seg_loss_fwd(logits, labels)
        _C.seg_loss_bwd(logits, labels, coef, torch.ones(1, device=DEV))


def model_step():
    g = torch.load(ROOT / "tests" / "golden" / "mae_small_4view.pt")
    model = CineMA(**g["kw"]).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.train()
    images = {k: v.to(DEV) for k, v in g["images"].items()}
    loss, _, _, _ = model(images, g["ratio"])
    loss.backward()
    tr = MAETrainer(model, lr=1e-3, use_cuda_graph=False)
    for _ in range(2):
        tr.step(images)
    torch.cuda.synchronize()
    print("loss", float(loss))


if __name__ == "__main__":
    which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"gemm", "attn", "ln", "model"}
    if "gemm" in which:
        gemms()
    if "attn" in which:
        attention()
    if "ln" in which:
        layernorm()
        input_pipeline()
    if "model" in which:
        model_step()
    torch.cuda.synchronize()
    print("sanitize target done:", sorted(which), _C.launches, "launches")
