"""In-kernel clock trace of the tcgen05 GEMM (cb_gemm_trace): the leader CTA of one pair in the middle of the grid stamps
clock64() at the phase boundaries of its producer warp, its MMA warp and its first epilogue warp for its first eight tiles.
    python tools/gemm_trace.py M N K [plain|bias|gelu|gelu_bwd|res32|wgrad]
Slot layout (ti = this worker's tile counter < 8): 64 ti + kb: producer got smem stage for k-block kb (< 16);
64 ti + 16 + kb: MMA warp saw k-block kb's data; 64 ti + 32: MMA warp got the accumulator stage; 64 ti + 33: last k-block issued
and committed; 64 ti + 40 / 41: epilogue warp starts / ends waiting for the accumulator; 64 ti + 42 + c: chunk c done;
64 ti + 50: accumulator released.  1000 entry, 1001 set-up done (after griddepcontrol.wait), 1002 roles done, 1003 after the
final cluster barrier."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cinema_b200 import _C  # noqa: E402

DEV, BF = "cuda", torch.bfloat16


def main():
    m, n, k = (int(a) for a in sys.argv[1:4])
    mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
    x = torch.randn(m, k, device=DEV).to(BF)
    w = (torch.randn(n, k, device=DEV) * 0.02).to(BF)
    bias = torch.randn(n, device=DEV)
    y, y2 = torch.empty(m, n, device=DEV, dtype=BF), torch.empty(m, n, device=DEV, dtype=BF)
    aux = torch.randn(m, n, device=DEV).to(BF)
    res = torch.randn(m, n, device=DEV)
    if mode == "wgrad":  # dW[N, K] += dy[M, N]^T x[M, K]: contraction over the M tokens
        dy, dw = torch.randn(m, n, device=DEV).to(BF), torch.zeros(n, k, device=DEV)
        fn = lambda: _C.gemm(dy, x, dw, a_mn=True, b_mn=True, accumulate=True)  # noqa: E731
    else:
        fn = {"plain": lambda: _C.gemm(x, w, y), "bias": lambda: _C.gemm(x, w, y, bias=bias),
              "gelu": lambda: _C.gemm(x, w, y, out2=y2, bias=bias, epilogue=_C.EPI_GELU),
              "gelu_bwd": lambda: _C.gemm(x, w, y, aux=aux, epilogue=_C.EPI_GELU_BWD),
              "res32": lambda: _C.gemm(x, w, res, bias=bias, residual=res)}[mode]
    buf = torch.zeros(1024, dtype=torch.int64, device=DEV)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"== M {m} N {n} K {k} {mode}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    assert _C.lib().cb_gemm_trace(buf.data_ptr()) == 0
    fn()
    torch.cuda.synchronize()
    _C.lib().cb_gemm_trace(None)
    t = buf.cpu().tolist()
    base = t[1000]
    print(f"set-up done {t[1001] - base}, roles done {t[1002] - base}, exit barrier {t[1003] - base}  (clocks from kernel entry)")
    for ti in range(8):
        s = t[64 * ti:64 * ti + 64]
        if not s[32]:
            break
        prod = [v - base for v in s[0:16] if v]
        mma = [v - base for v in s[16:32] if v]
        chunks = [v - base for v in s[42:50] if v]
        print(f"  tile {ti}: producer kb {prod[0]} .. {prod[-1]} | mma: acc free {s[32] - base}, kb data {mma[0]} .. {mma[-1]}, committed {s[33] - base}"
              f" | epilogue: wait {s[40] - base} -> {s[41] - base}, chunks {chunks}, released {s[50] - base}")
        if len(mma) > 1:
            d = [b - a for a, b in zip(mma, mma[1:])]
            print(f"           mma k-block intervals: {d}")


if __name__ == "__main__":
    main()
