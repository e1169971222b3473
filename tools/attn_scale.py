"""Attention forward / backward time against the number of key tiles and query tiles (head_dim 32 and 64, B = 16): the
slope gives the cost per 128 x 128 tile over the whole grid, the intercept the fixed cost per CTA
(profiles/r01_attention_clock_trace.md).   python tools/attn_scale.py"""
import sys, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from cinema_b200 import _C
DEV='cuda'; BF=torch.bfloat16
def run(nq, nk, h, d, B=16):
    q=torch.randn(B,nq,h,d,device=DEV).to(BF); k=torch.randn(B,nk,h,d,device=DEV).to(BF); v=torch.randn(B,nk,h,d,device=DEV).to(BF)
    o=torch.empty_like(q); lse=torch.empty(B,h,nq,device=DEV)
    do=torch.randn_like(q); dq=torch.empty_like(q); dk=torch.empty_like(k); dv=torch.empty_like(v)
    delta,dqa=_C.attention_bwd_workspace(B,h,nq,d,DEV)
    sc=d**-0.5
    res=[]
    for fn in (lambda: _C.attention_fwd(q,k,v,o,lse,sc), lambda: _C.attention_bwd(q,k,v,o,do,lse,dq,dk,dv,delta,dqa,sc)):
        for _ in range(2): fn()
        ts=[]
        for _ in range(5):
            torch.cuda._sleep(1_000_000)
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); 
            for _ in range(3): fn()
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)/3)
        res.append(sorted(ts)[2]*1e3)
    return res
for (nq,nk,h,d) in [(2048,128,16,32),(2048,256,16,32),(2048,512,16,32),(2048,1024,16,32),(2048,2048,16,32),(256,2048,16,32),(1024,2048,16,32),(768,128,12,64),(768,768,12,64),(768,1536,12,64)]:
    f,b=run(nq,nk,h,d)
    print(f"nq={nq:5d} nk={nk:5d} h={h} d={d}: fwd {f:8.1f} us  bwd {b:8.1f} us")
