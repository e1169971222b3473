"""In-situ timing of the encoder MLP forward sequence (LN -> fc1+GELU -> fc2+residual) x 12, back to back, events per kernel."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C
DEV = "cuda"; BF = torch.bfloat16
m, d, h = 10960, 768, 3072
x = torch.randn(m, d, device=DEV)
g, b = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
w1 = (torch.randn(h, d, device=DEV) * 0.02).to(BF); b1 = torch.randn(h, device=DEV)
w2 = (torch.randn(d, h, device=DEV) * 0.02).to(BF); b2 = torch.randn(d, device=DEV)
L = 12
bufs = [dict(h16=torch.empty(m, d, device=DEV, dtype=BF), pre=torch.empty(m, h, device=DEV, dtype=BF), act=torch.empty(m, h, device=DEV, dtype=BF),
             mean=torch.empty(m, device=DEV), rstd=torch.empty(m, device=DEV), x2=torch.empty(m, d, device=DEV)) for _ in range(L)]
def run(ev=None):
    cur = x
    for i in range(L):
        t = bufs[i]
        def rec(name, fn):
            if ev is None: fn(); return
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); ev.append((name, s, e))
        rec("ln", lambda: _C.layernorm_fwd(cur, g, b, 1e-5, y16=t["h16"], mean=t["mean"], rstd=t["rstd"]))
        rec("fc1", lambda: _C.gemm(t["h16"], w1, t["pre"], out2=t["act"], bias=b1, epilogue=_C.EPI_GELU))
        rec("fc2", lambda: _C.gemm(t["act"], w2, t["x2"], bias=b2, residual=cur))
        cur = t["x2"]
for pdl in (0, 1):
    _C.set_pdl(bool(pdl))
    for _ in range(2): run()
    torch.cuda.synchronize()
    ev = []
    torch.cuda._sleep(200_000_000)
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); run(ev); e0.record(); torch.cuda.synchronize()
    agg = {}
    for n, s, e in ev: agg.setdefault(n, []).append(s.elapsed_time(e) * 1e3)
    print("pdl", pdl, "total %.1f us" % (s0.elapsed_time(e0) * 1e3), {k: "%.1f" % (sum(v) / len(v)) for k, v in agg.items()}, "first fc1 %.1f last %.1f" % (agg["fc1"][0], agg["fc1"][-1]))
    torch.cuda._sleep(200_000_000)
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); run(None); e0.record(); torch.cuda.synchronize()
    print("   no per-kernel events: total %.1f us" % (s0.elapsed_time(e0) * 1e3))
