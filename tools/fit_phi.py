"""Fit of the epilogue normal-CDF approximation used by csrc/common.cuh::phi_fast:
Phi(x) ~= 1 / (1 + exp(-x * P(x^2))), P of degree 4 in x^2, minimax by iteratively re-weighted least squares.
Prints the coefficients (and their -log2(e)-folded form used with ex2) and checks the error in float32 arithmetic."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

x = np.linspace(-6.0, 6.0, 40001)
phi = 0.5 * (1 + erf(x / np.sqrt(2)))


def res(c):
    u = np.clip(x * sum(c[i] * x ** (2 * i) for i in range(len(c))), -80, 80)
    return 1 / (1 + np.exp(-u)) - phi


c = np.array([1.5957691, 0.0713548, -6e-4, 0, 0], float)
w = np.ones_like(x)
best = None
for _ in range(80):
    c = least_squares(lambda c: res(c) * w, c, x_scale=np.abs(c) + 1e-7).x
    e = np.abs(res(c))
    if best is None or e.max() < best[0]:
        best = (e.max(), c.copy())
    w = w * (1 + 4 * e / e.max())
    w /= w.mean()
c = best[1]
print("coefficients", c, "max |Phi err|", best[0])
f = (-c * np.log2(np.e)).astype(np.float32)
print("folded (ex2 form)", [f"{v:.9g}" for v in f])
xs = np.linspace(-12, 12, 200001).astype(np.float32)
x2 = xs * xs
p = f[4]
for k in (3, 2, 1, 0):
    p = p * x2 + f[k]
with np.errstate(over="ignore"):
    approx = 1 / (1 + np.exp2(xs * p))
exact = 0.5 * (1 + erf(xs.astype(np.float64) / np.sqrt(2)))
print("float32 check on [-12, 12]: max |Phi err|", np.abs(approx - exact).max(), " max |gelu err|", np.abs(xs * (approx - exact)).max())

# --check-grad: the tanh form used by gelu_grad_fast (Q = -ln2/2 * P) over every bf16 input in [-30, 30], tanh.approx
# modelled at its worst-case relative error 2^-11
q = (c * 0.5).astype(np.float32)
print("tanh-form coefficients (gelu_grad_fast)", [f"{v:.9g}" for v in q])
u16 = np.arange(0, 65536, dtype=np.uint32)
xb = (u16 << 16).view(np.float32)
xb = xb[np.isfinite(xb) & (np.abs(xb) <= 30)]
xd = xb.astype(np.float64)
gp = 0.5 * (1 + erf(xd / np.sqrt(2))) + xd * np.exp(-0.5 * xd ** 2) / np.sqrt(2 * np.pi)
x2 = xb * xb
qq = q[4]
for k in (3, 2, 1, 0):
    qq = qq * x2 + q[k]
phi_t = 0.5 + 0.5 * np.tanh((xb * qq).astype(np.float64)) * (1 + 2.0 ** -11)
gq = phi_t + xd * 0.3989422804014327 * np.exp2(-0.72134752044448170 * xd ** 2)
print("GELU' (tanh form): max abs err", np.abs(gq - gp).max(), " rms", np.sqrt(((gq - gp) ** 2).mean()))
