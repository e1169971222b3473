"""Epilogue cost isolation on one GEMM shape (M=10960, N=3072, K=768, dgrad orientation)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C
from tools.perf_gemm import timeit
DEV = "cuda"
m, n, k = (32848, 2048, 512) if "--dec" in sys.argv else (10960, 3072, 768)
print(f"M {m} N {n} K {k}")
dy = torch.randn(m, k, device=DEV).bfloat16()
w = (torch.randn(k, n, device=DEV) * 0.02).bfloat16()
wk = (torch.randn(n, k, device=DEV) * 0.02).bfloat16()
dx = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
dx2 = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
dx32 = torch.empty(m, n, device=DEV)
aux = torch.randn(m, n, device=DEV).bfloat16()
res = torch.randn(m, n, device=DEV)
bias = torch.randn(n, device=DEV)
fl = 2.0 * m * n * k
cases = {
    "plain bf16 (B mn-major)": lambda: _C.gemm(dy, w, dx, b_mn=True),
    "plain bf16 (B k-major)": lambda: _C.gemm(dy, wk, dx),
    "+bias": lambda: _C.gemm(dy, wk, dx, bias=bias),
    "+bias fp32 out": lambda: _C.gemm(dy, wk, dx32, bias=bias),
    "+residual, bf16 out": lambda: _C.gemm(dy, wk, dx, residual=res),
    "+residual, fp32 out": lambda: _C.gemm(dy, wk, dx32, residual=res),
    "fp32 out + bf16 shadow": lambda: _C.gemm(dy, wk, dx32, out2=dx),
    "GELU (2 bf16 outs)": lambda: _C.gemm(dy, wk, dx, out2=dx2, bias=bias, epilogue=_C.EPI_GELU),
    "GELU (act only)": lambda: _C.gemm(dy, wk, None, out2=dx2, bias=bias, epilogue=_C.EPI_GELU),
    "GELU' (aux)": lambda: _C.gemm(dy, w, dx, b_mn=True, aux=aux, epilogue=_C.EPI_GELU_BWD),
}
for name, fn in cases.items():
    t = timeit(fn)
    print(f"{name:28s} {t*1e3:7.1f} us {fl/t/1e9:6.0f} TF/s")
