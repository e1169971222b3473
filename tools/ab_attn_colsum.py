import sys, torch
sys.path.insert(0, '.')
from cinema_b200 import _C
DEV='cuda'; BF=torch.bfloat16
def run(B,nq,nk,h,d, fused_layout):
    if fused_layout:
        qkv=torch.randn(B,nq,3,h,d,device=DEV).to(BF); q,k,v=qkv[:,:,0],qkv[:,:,1],qkv[:,:,2]
        dqkv=torch.empty_like(qkv); dq,dk,dv=dqkv[:,:,0],dqkv[:,:,1],dqkv[:,:,2]
    else:
        q=torch.randn(B,nq,h,d,device=DEV).to(BF); k=torch.randn(B,nk,h,d,device=DEV).to(BF); v=torch.randn(B,nk,h,d,device=DEV).to(BF)
        dq,dk,dv=torch.empty_like(q),torch.empty_like(k),torch.empty_like(v)
    o=torch.empty(B,nq,h,d,device=DEV,dtype=BF); lse=torch.empty(B,h,nq,device=DEV); do=torch.randn(B,nq,h,d,device=DEV).to(BF)
    delta,dqa=_C.attention_bwd_workspace(B,h,nq,d,DEV); sc=d**-0.5
    _C.attention_fwd(q,k,v,o,lse,sc)
    cs=[torch.zeros(h*d,device=DEV) for _ in range(3)]
    for name,args in (("plain",()),("dq only",(cs[0],None,None)),("dq+dk+dv",tuple(cs))):
        fn=lambda: _C.attention_bwd(q,k,v,o,do,lse,dq,dk,dv,delta,dqa,sc,*args)
        for _ in range(3): fn()
        ts=[]
        for _ in range(7):
            torch.cuda._sleep(1_000_000); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): fn()
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)/4)
        print(f"d{d} fused_layout={fused_layout} {name:10s} {sorted(ts)[3]*1e3:8.1f} us (delta + memset + bwd + convert)")
run(16,685,685,12,64,True); run(16,685,685,12,64,False); run(16,2053,684,16,32,False)
