"""Micro-benchmark of cb_gemm_bf16 against torch.matmul (cuBLAS) on the dominant shapes of the ViT-B MAE step
(tools/gemm_shapes.py).  Run on the GPU box: python tools/perf_gemm.py [--bn 0]"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=12, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    # (name, M, N, K) forward orientation y[M,N] = x[M,K] w[N,K]^T ; count per step
    shapes = [("enc.fc1 gelu", 10960, 3072, 768, 12), ("enc.fc2 res", 10960, 768, 3072, 12), ("enc.qkv", 10960, 2304, 768, 12),
              ("enc.proj res", 10960, 768, 768, 12), ("dec.fc1 gelu", 32848, 2048, 512, 8), ("dec.fc2 res", 32848, 512, 2048, 8),
              ("dec.q/proj", 32848, 512, 512, 16), ("dec.kv_all", 10944, 8192, 512, 1)]
    if a.quick:
        shapes = shapes[:2]
    print(f"{'shape':14s} {'M':>6s} {'N':>5s} {'K':>5s} | {'fwd us':>7s} {'TF/s':>6s} {'cublas':>6s} | {'dgrad':>7s} {'TF/s':>6s} {'cublas':>6s} | {'wgrad':>7s} {'TF/s':>6s} {'cublas':>6s} | step ms")
    total = 0.0
    for name, m, n, k, cnt in shapes:
        x = torch.randn(m, k, device=DEV).bfloat16()
        w = (torch.randn(n, k, device=DEV) * 0.02).bfloat16()
        dy = torch.randn(m, n, device=DEV).bfloat16()
        bias = torch.randn(n, device=DEV)
        y = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        y2 = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        y32 = torch.empty(m, n, device=DEV)
        res = torch.randn(m, n, device=DEV)
        dx = torch.empty(m, k, device=DEV, dtype=torch.bfloat16)
        aux = torch.randn(m, k, device=DEV).bfloat16()
        dw = torch.zeros(n, k, device=DEV)
        fl = 2.0 * m * n * k
        if "gelu" in name:
            f_fwd = lambda: _C.gemm(x, w, y, out2=y2, bias=bias, epilogue=_C.EPI_GELU, block_n=a.bn)
            f_dg = lambda: _C.gemm(dy, w, dx, b_mn=True, block_n=a.bn)
        elif "res" in name:
            f_fwd = lambda: _C.gemm(x, w, y32, bias=bias, residual=res, block_n=a.bn)
            # dgrad of fc2 / proj feeds GELU' (fc2) -- time the GELU' variant for fc2
            if "fc2" in name:
                f_dg = lambda: _C.gemm(dy, w, dx, b_mn=True, aux=aux, epilogue=_C.EPI_GELU_BWD, block_n=a.bn)
            else:
                f_dg = lambda: _C.gemm(dy, w, dx, b_mn=True, block_n=a.bn)
        else:
            f_fwd = lambda: _C.gemm(x, w, y, bias=bias, block_n=a.bn)
            f_dg = lambda: _C.gemm(dy, w, dx, b_mn=True, block_n=a.bn)
        t_f = timeit(f_fwd)
        t_c = timeit(lambda: torch.matmul(x, w.t()))
        t_d = timeit(f_dg)
        t_dc = timeit(lambda: torch.matmul(dy, w))
        t_w = timeit(lambda: _C.gemm(dy, x, dw, a_mn=True, b_mn=True, accumulate=True))
        t_wc = timeit(lambda: torch.matmul(dy.t(), x))
        tf = lambda t: fl / t / 1e9  # noqa: E731
        step = cnt * (t_f + t_d + t_w)
        total += step
        print(f"{name:14s} {m:6d} {n:5d} {k:5d} | {t_f*1e3:7.1f} {tf(t_f):6.0f} {tf(t_c):6.0f} | {t_d*1e3:7.1f} {tf(t_d):6.0f} {tf(t_dc):6.0f} | {t_w*1e3:7.1f} {tf(t_w):6.0f} {tf(t_wc):6.0f} | {step:6.2f}")
    print(f"sum over the step's dominant GEMMs: {total:.2f} ms")


if __name__ == "__main__":
    main()
