"""Micro-benchmark of cb_gemm_bf16 against torch.matmul (cuBLAS) on the shapes of the ViT-B MAE step.
Run on the GPU box: python tools/perf_gemm.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=20, warm=5):
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
    B = 16
    shapes = []
    for tag, tok, d in [("enc", 769, 768), ("dec", 2305, 512)]:
        m = B * tok
        shapes += [(f"{tag}.qkv/fc-like", m, 3 * d, d), (f"{tag}.proj", m, d, d), (f"{tag}.fc1", m, 4 * d, d),
                   (f"{tag}.fc2", m, d, 4 * d)]
    print(f"{'shape':22s} {'M':>6s} {'N':>5s} {'K':>5s} | {'fwd us':>8s} {'TF/s':>7s} {'cublas':>7s} | {'dgrad':>7s} {'TF/s':>6s} | {'wgrad':>7s} {'TF/s':>6s} {'cublas':>7s}")
    for name, m, n, k in shapes:
        x = torch.randn(m, k, device=DEV).bfloat16()
        w = (torch.randn(n, k, device=DEV) * 0.02).bfloat16()
        dy = torch.randn(m, n, device=DEV).bfloat16()
        y = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        dx = torch.empty(m, k, device=DEV, dtype=torch.bfloat16)
        dw = torch.zeros(n, k, device=DEV)
        fl = 2.0 * m * n * k
        t_f = timeit(lambda: _C.gemm(x, w, y))
        t_c = timeit(lambda: torch.matmul(x, w.t()))
        t_d = timeit(lambda: _C.gemm(dy, w, dx, b_mn=True))
        t_w = timeit(lambda: _C.gemm(dy, x, dw, a_mn=True, b_mn=True, accumulate=True))
        t_wc = timeit(lambda: torch.matmul(dy.t(), x))
        tf = lambda t: fl / t / 1e9  # noqa: E731
        print(f"{name:22s} {m:6d} {n:5d} {k:5d} | {t_f*1e3:8.1f} {tf(t_f):7.0f} {tf(t_c):7.0f} | {t_d*1e3:7.1f} {tf(t_d):6.0f} | {t_w*1e3:7.1f} {tf(t_w):6.0f} {tf(t_wc):7.0f}")
    # epilogue variants on the fc1 shape
    m, n, k = B * 769, 3072, 768
    x = torch.randn(m, k, device=DEV).bfloat16()
    w = (torch.randn(n, k, device=DEV) * 0.02).bfloat16()
    bias = torch.randn(n, device=DEV)
    pre = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    act = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    t = timeit(lambda: _C.gemm(x, w, pre, out2=act, bias=bias, epilogue=_C.EPI_GELU))
    print(f"fc1 + bias + GELU (2 outputs): {t*1e3:.1f} us  {2.0*m*n*k/t/1e9:.0f} TF/s")
    res = torch.randn(m, k, device=DEV)
    h = torch.randn(m, n, device=DEV).bfloat16()
    w2 = (torch.randn(k, n, device=DEV) * 0.02).bfloat16()
    b2 = torch.randn(k, device=DEV)
    t = timeit(lambda: _C.gemm(h, w2, res, bias=b2, residual=res))
    print(f"fc2 + bias + fp32 residual in place: {t*1e3:.1f} us  {2.0*m*n*k/t/1e9:.0f} TF/s")
    for bn in (64, 128, 256):
        t = timeit(lambda: _C.gemm(x, w, pre, block_n=bn))
        print(f"fc1 plain block_n={bn}: {t*1e3:.1f} us  {2.0*m*n*k/t/1e9:.0f} TF/s")


if __name__ == "__main__":
    main()
