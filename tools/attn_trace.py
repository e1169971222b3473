"""In-kernel clock trace of the second-generation attention kernels (cb_attention_trace): one CTA in the middle of the grid
stamps clock64() at its phase boundaries; this prints the per-iteration deltas.   python tools/attn_trace.py [fwd|bwd] [d]

Slot layout.  forward (j = key tile < 8): 16 j + {0 loop top, 1 s_full ok, 2 S in registers, 3 row max done, 4 pv_done ok,
5 exp2 + P stored} (softmax warp 0), + {8 s_free ok, 9 S(j+1) issued, 10 p_full ok, 11 P.V issued} (MMA warp);
240 kernel entry, 241 setup done, 242 last tile done, 243 last pv_done ok, 244 epilogue stored, 245 exit.
backward (i = query tile < 24): 32 i + 8 hf + {0 top, 1 s_full ok, 2 S^T loaded, 3 dP^T loaded + vectors, 4 math done,
5 pdk / ds waits ok, 6 stores done} (compute warp 0 of half hf), + {16 S/dP(i) issued (issuer A), 17 / 19 ps_ready[0 / 1] ok,
18 / 20 dV dK issued, 21 dq_free ok, 22 dQ issued} (MMA warp), + {24 dq_full ok, 25 drained} (drain warp 0);
1000 entry, 1001 setup done, 1002 loop done, 1003 dK / dV stored."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C  # noqa: E402

DEV, BF = "cuda", torch.bfloat16


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    B, nq, nk, h = (16, 2053, 684, 16) if d == 32 else (16, 685, 685, 12)
    q, k, v = (torch.randn(B, n, h, d, device=DEV).to(BF) for n in (nq, nk, nk))
    o, lse = torch.empty_like(q), torch.empty(B, h, nq, device=DEV)
    do = torch.randn_like(q)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta, dqa = _C.attention_bwd_workspace(B, h, nq, d, DEV)
    sc = d ** -0.5
    buf = torch.zeros(1024, dtype=torch.int64, device=DEV)
    for _ in range(2):
        _C.attention_fwd(q, k, v, o, lse, sc)
        _C.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dqa, sc)
    torch.cuda.synchronize()
    assert _C.lib().cb_attention_trace(buf.data_ptr()) == 0
    if which == "fwd":
        _C.attention_fwd(q, k, v, o, lse, sc)
    else:
        _C.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, delta, dqa, sc)
    torch.cuda.synchronize()
    _C.lib().cb_attention_trace(None)
    t = buf.cpu().tolist()
    if which == "fwd":
        base = t[240]
        print(f"forward d{d}: setup {t[241] - base}, last tile done {t[242] - base}, epilogue stored {t[244] - base}, exit {t[245] - base}")
        names = {0: "top", 1: "s_full ok", 2: "S loaded", 3: "max done", 4: "pv_done ok", 5: "exp+P stored", 8: "mma: s_free ok",
                 9: "mma: S(j+1) issued", 10: "mma: p_full ok", 11: "mma: PV issued"}
        for j in range(8):
            row = {s: t[16 * j + s] for s in names if t[16 * j + s]}
            if not row:
                break
            print(f"  j={j}: " + "  ".join(f"{names[s]} {val - base}" for s, val in sorted(row.items(), key=lambda kv: kv[1])))
    else:
        base = t[1000]
        print(f"backward d{d}: setup {t[1001] - base}, loop done {t[1002] - base}, dK/dV stored {t[1003] - base}")
        names = {0: "h0 top", 1: "h0 s_full", 2: "h0 S ld", 3: "h0 dP ld", 4: "h0 math", 5: "h0 waits", 6: "h0 stored",
                 8: "h1 top", 9: "h1 s_full", 10: "h1 S ld", 11: "h1 dP ld", 12: "h1 math", 13: "h1 waits", 14: "h1 stored",
                 16: "mmaA SdP(i)", 17: "mma ps0 ok", 18: "mma dVdK0", 19: "mma ps1 ok", 20: "mma dVdK1", 21: "mma dq_free", 22: "mma dQ",
                 24: "drain dq_full", 25: "drain done"}
        for i in range(24):
            row = {s: t[32 * i + s] for s in names if t[32 * i + s]}
            if not row:
                break
            print(f"  i={i}: " + "  ".join(f"{names[s]} {val - base}" for s, val in sorted(row.items(), key=lambda kv: kv[1])))
            if i >= int(sys.argv[3]) if len(sys.argv) > 3 else False:
                break


if __name__ == "__main__":
    main()
