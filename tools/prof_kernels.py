"""One launch of each hot kernel at the bench shapes (ViT-B 4-view, B=16), for `ncu --set full` captures and for
warm CUDA-event timings.   python tools/prof_kernels.py [--time] [--only ln,attn,gemm,dw,misc]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def rn(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


def mk(f, *a, **k):
    return lambda: f(*a, **k)


def build(only):
    K = []  # (name, fn, bytes, flops)
    B = 16
    if "ln" in only:
        for tag, m, d in (("enc", 10960, 768), ("dec", 32848, 512), ("stem0", 147456, 64)):
            x, g, b = rn(m, d), rn(d), rn(d)
            y16 = torch.empty(m, d, device=DEV, dtype=BF)
            mean, rstd = torch.empty(m, device=DEV), torch.empty(m, device=DEV)
            K.append((f"ln_fwd {tag} {m}x{d}", lambda x=x, g=g, b=b, y16=y16, mean=mean, rstd=rstd: _C.layernorm_fwd(x, g, b, 1e-5, y16=y16, mean=mean, rstd=rstd), m * d * 6, 0))
            dy, dres = rn(m, d, dtype=BF), rn(m, d)
            dx16 = torch.empty(m, d, device=DEV, dtype=BF)
            dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
            _C.layernorm_fwd(x, g, b, 1e-5, y16=y16, mean=mean, rstd=rstd)
            K.append((f"ln_bwd {tag} {m}x{d}", lambda dy=dy, x=x, mean=mean, rstd=rstd, g=g, dres=dres, dx16=dx16, dg=dg, db=db: _C.layernorm_bwd(dy, x, mean, rstd, g, dres=dres, dx32=dres, dx16=dx16, dgamma=dg, dbeta=db), m * d * 16, 0))
    if "attn" in only:
        for tag, nq, nk, h, d in (("enc", 685, 685, 12, 64), ("dec", 2053, 684, 16, 32)):
            q, k, v = (rn(B, n, h, d, dtype=BF) for n in (nq, nk, nk))
            o = torch.empty(B, nq, h, d, device=DEV, dtype=BF)
            lse = torch.empty(B, h, nq, device=DEV)
            sc = d ** -0.5
            fl = 4.0 * B * h * nq * nk * d
            K.append((f"attn_fwd {tag} d{d}", lambda q=q, k=k, v=v, o=o, lse=lse, sc=sc: _C.attention_fwd(q, k, v, o, lse, sc), 0, fl))
            do = rn(B, nq, h, d, dtype=BF)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            delta, dqa = _C.attention_bwd_workspace(B, h, nq, d, DEV)
            _C.attention_fwd(q, k, v, o, lse, sc)
            K.append((f"attn_bwd {tag} d{d}", lambda q=q, k=k, v=v, o=o, do=do, lse=lse, dq=dq, dk=dk, dv=dv, delta=delta, dqa=dqa, sc=sc: _C.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dqa, sc), 0, 2 * fl))
    if "gemm" in only:
        m, n, k = 10960, 3072, 768
        x, w, bias = rn(m, k, dtype=BF), rn(n, k, dtype=BF, scale=0.02), rn(n)
        y, y2 = torch.empty(m, n, device=DEV, dtype=BF), torch.empty(m, n, device=DEV, dtype=BF)
        K.append(("gemm fc1+gelu enc", mk(_C.gemm, x, w, y, out2=y2, bias=bias, epilogue=_C.EPI_GELU), 0, 2.0 * m * n * k))
        dy, dx, aux = rn(m, k, dtype=BF), torch.empty(m, n, device=DEV, dtype=BF), rn(m, n, dtype=BF)
        w2 = rn(k, n, dtype=BF, scale=0.02)
        K.append(("gemm fc2 dgrad gelu' enc", mk(_C.gemm, dy, w2, dx, b_mn=True, aux=aux, epilogue=_C.EPI_GELU_BWD), 0, 2.0 * m * n * k))
        h, res, b2 = rn(m, n, dtype=BF), rn(m, k), rn(k)
        K.append(("gemm fc2+res enc", mk(_C.gemm, h, w2, res, bias=b2, residual=res), 0, 2.0 * m * n * k))
        dw = torch.zeros(n, k, device=DEV)
        dyy = rn(m, n, dtype=BF)
        K.append(("gemm wgrad fc1 enc", mk(_C.gemm, dyy, x, dw, a_mn=True, b_mn=True, accumulate=True), 0, 2.0 * m * n * k))
        md, nd_, kd = 32848, 2048, 512
        xd, dyd, dwd = rn(md, kd, dtype=BF), rn(md, nd_, dtype=BF), torch.zeros(nd_, kd, device=DEV)
        K.append(("gemm wgrad fc1 dec", mk(_C.gemm, dyd, xd, dwd, a_mn=True, b_mn=True, accumulate=True), 0, 2.0 * md * nd_ * kd))
        # decoder (D = 512): the short-K shapes
        xq, wq, bq = rn(md, kd, dtype=BF), rn(nd_, kd, dtype=BF, scale=0.02), rn(nd_)
        yq, yq2 = torch.empty(md, nd_, device=DEV, dtype=BF), torch.empty(md, nd_, device=DEV, dtype=BF)
        K.append(("gemm fc1+gelu dec", mk(_C.gemm, xq, wq, yq, out2=yq2, bias=bq, epilogue=_C.EPI_GELU), 0, 2.0 * md * nd_ * kd))
        wq2, auxq = rn(kd, nd_, dtype=BF, scale=0.02), rn(md, nd_, dtype=BF)
        K.append(("gemm fc2 dgrad gelu' dec", mk(_C.gemm, xq, wq2, yq, b_mn=True, aux=auxq, epilogue=_C.EPI_GELU_BWD), 0, 2.0 * md * nd_ * kd))
        wp, yp = rn(kd, kd, dtype=BF, scale=0.02), torch.empty(md, kd, device=DEV, dtype=BF)
        K.append(("gemm q-proj dec 512x512", mk(_C.gemm, xq, wp, yp, bias=rn(kd)), 0, 2.0 * md * kd * kd))
        K.append(("gemm dgrad dec 512x512", mk(_C.gemm, xq, wp, yp, b_mn=True), 0, 2.0 * md * kd * kd))
        resq = rn(md, kd)
        K.append(("gemm proj+res dec 512x512", mk(_C.gemm, xq, wp, resq, bias=rn(kd), residual=resq), 0, 2.0 * md * kd * kd))
        xs, ws, ys = rn(147456, 64, dtype=BF), rn(256, 64, dtype=BF), torch.empty(147456, 256, device=DEV, dtype=BF)
        ys2, bs = torch.empty(147456, 256, device=DEV, dtype=BF), rn(256)
        K.append(("gemm stem fc1+gelu 147456x256x64", mk(_C.gemm, xs, ws, ys, out2=ys2, bias=bs, epilogue=_C.EPI_GELU), 147456 * (64 + 512) * 2, 2.0 * 147456 * 256 * 64))
    if "misc" in only:
        xc, oc = rn(10960, 3072, dtype=BF), torch.zeros(3072, device=DEV)
        K.append(("colsum 10960x3072", mk(_C.colsum, xc, oc), 10960 * 3072 * 2, 0))
        xc2, oc2 = rn(32848, 512, dtype=BF), torch.zeros(512, device=DEV)
        K.append(("colsum 32848x512", mk(_C.colsum, xc2, oc2), 32848 * 512 * 2, 0))
    if "misc" in only:  # input pipeline: RandZoom + ScaleIntensity + padding of one batch (SAX int16, LAX uint8)
        for tag, shape, dt in (("sax 192x192x16", (B, 1, 192, 192, 16), torch.int16), ("lax 192x192", (B, 1, 192, 192), torch.uint8)):
            raw = torch.randint(0, 255, shape, device=DEV, dtype=torch.int32).to(dt)
            nd = len(shape) - 2
            extent = torch.tensor([list(shape[2:]) + [1] * (3 - nd)] * B, dtype=torch.int32, device=DEV)
            zoom = (0.9 + 0.2 * torch.rand(B, device=DEV)).contiguous()
            out = torch.empty(shape, device=DEV)
            keys = torch.empty(2 * B, dtype=torch.int32, device=DEV)
            n = raw.numel()
            K.append((f"zoom+scale+pad {tag}", mk(_C.zoom_intensity, raw, extent, zoom, out, keys), n * (raw.element_size() + 12), 0))
            lo, hi = torch.zeros(B, device=DEV), torch.full((B,), 255.0, device=DEV)
            K.append((f"scale_intensity {tag}", mk(_C.scale_intensity, raw, lo, hi, out), n * (raw.element_size() + 4), 0))
    if "dw" in only:
        grid, f, c, nk = (12, 12, 16), (4, 4, 1), 64, 576
        n = 12 * 12 * 16
        mask = torch.ones(B, n, dtype=torch.bool, device=DEV)
        g = torch.Generator(device=DEV).manual_seed(0)
        for i in range(B):
            mask[i, torch.randperm(n, device=DEV, generator=g)[:nk]] = False
        keep, drop, slot = _C.mask_to_index(mask, nk)
        x = rn(B * nk * 16, c, dtype=BF)
        w = rn(c, 1, 5, 5, 5, dtype=BF, scale=0.2)
        out = torch.empty_like(x)
        bias = rn(c)
        K.append(("dwconv fwd SAX L0", mk(_C.dwconv_tokens, x, out, w, bias, mask, slot, keep, grid, f), 0, 2.0 * B * nk * 16 * 125 * c))
        dw, db = torch.zeros(c, 1, 5, 5, 5, device=DEV), torch.zeros(c, device=DEV)
        K.append(("dwconv wgrad SAX L0", mk(_C.dwconv_tokens_wgrad, x, out, dw, db, mask, slot, keep, grid, f), 0, 2.0 * B * nk * 16 * 125 * c))
    return K


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", default="ln,attn,gemm,dw,misc")
    a = ap.parse_args()
    kernels = build(set(a.only.split(",")))
    for name, fn, by, fl in kernels:  # warm-up (not profiled when ncu is given -s)
        fn()
    torch.cuda.synchronize()
    if not a.time:
        for name, fn, by, fl in kernels:
            fn()
        torch.cuda.synchronize()
        return
    for name, fn, by, fl in kernels:
        ts = []
        reps = 4
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(1_500_000)  # let the host run ahead: launches are queued, not timed
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        t = sorted(ts)[len(ts) // 2]
        extra = f"{by / t / 1e6:8.0f} GB/s" if by else ""
        extra += f"{fl / t / 1e9:8.0f} TF/s" if fl else ""
        print(f"{name:36s} {t * 1e3:9.1f} us {extra}")


if __name__ == "__main__":
    main()
