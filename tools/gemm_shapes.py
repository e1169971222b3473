"""List every GEMM / attention launch of one MAE training step (shapes only, no math): host logic is run on CPU with
the C-ABI wrappers replaced by recorders.  Used to size kernel work; not part of the product."""
import collections, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import model_kwargs, synthetic_batch
from cinema_b200 import _C, CineMA, engine
from tests import emu_c

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rec = collections.OrderedDict()
def gemm(a, b, out, *, a_mn=False, b_mn=False, accumulate=False, out2=None, bias=None, residual=None, aux=None, epilogue=0, alpha=1.0, split_k=0, block_n=0):
    m, k = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    n = b.shape[1] if b_mn else b.shape[0]
    key = ("gemm", m, n, k, int(a_mn), int(b_mn), int(accumulate), epilogue, "f32" if (out is not None and out.dtype == torch.float32) else "bf16", residual is not None)
    rec[key] = rec.get(key, 0) + 1
def attention_fwd(q, k, v, o, lse, scale):
    key = ("attn_fwd", tuple(q.shape), k.shape[1]); rec[key] = rec.get(key, 0) + 1
def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dq_acc, scale):
    key = ("attn_bwd", tuple(q.shape), k.shape[1]); rec[key] = rec.get(key, 0) + 1
def ln_fwd(x, *a, **k):
    key = ("ln_fwd", tuple(x.shape)); rec[key] = rec.get(key, 0) + 1
def ln_bwd(dy, x, *a, **k):
    key = ("ln_bwd", tuple(x.shape), str(dy.dtype)); rec[key] = rec.get(key, 0) + 1
def colsum(x, out):
    key = ("colsum", tuple(x.shape)); rec[key] = rec.get(key, 0) + 1
keepers = {"mask_to_index", "expand_token_index", "device_info"}
for name in emu_c.ALL:
    if name in keepers:
        setattr(_C, name, getattr(emu_c, name))
    else:
        setattr(_C, name, lambda *a, **k: None)
_C.gemm, _C.attention_fwd, _C.attention_bwd, _C.layernorm_fwd, _C.layernorm_bwd, _C.colsum = gemm, attention_fwd, attention_bwd, ln_fwd, ln_bwd, colsum
engine.check_head_dim = lambda d: None
kw = model_kwargs("base", (192, 192, 16), (192, 192))
model = CineMA(**kw); model.train()
batch = synthetic_batch(kw, B, 0, False)
loss, *_ = model(batch, 0.75)
loss.backward()
tot = 0
for k, c in sorted(rec.items(), key=lambda kv: -(2.0 * kv[0][1] * kv[0][2] * kv[0][3] * kv[1]) if kv[0][0] == "gemm" else 0):
    if k[0] == "gemm":
        fl = 2.0 * k[1] * k[2] * k[3] * c; tot += fl
        print(f"{c:4d} x gemm M={k[1]:7d} N={k[2]:6d} K={k[3]:7d} a_mn={k[4]} b_mn={k[5]} acc={k[6]} epi={k[7]} out={k[8]} res={k[9]}  {fl/1e9:9.1f} GF")
    else:
        print(c, "x", k)
print("total gemm GF", tot / 1e9)
