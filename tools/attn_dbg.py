import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
from cinema_b200 import _C
DEV='cuda'; BF=torch.bfloat16
B,nq,nk,h,d=16,2053,684,16,32
q=torch.randn(B,nq,h,d,device=DEV).to(BF); k=torch.randn(B,nk,h,d,device=DEV).to(BF); v=torch.randn(B,nk,h,d,device=DEV).to(BF)
o=torch.empty_like(q); lse=torch.empty(B,h,nq,device=DEV)
do=torch.randn_like(q); dq=torch.empty_like(q); dk=torch.empty_like(k); dv=torch.empty_like(v)
delta,dqa=_C.attention_bwd_workspace(B,h,nq,d,DEV)
_C.attention_fwd(q,k,v,o,lse,d**-0.5)
for _ in range(3): _C.attention_bwd(q,k,v,o,do,lse,dq,dk,dv,delta,dqa,d**-0.5)
torch.cuda.synchronize()
buf=(ctypes.c_longlong*512)()
_C.lib().cb_attn_dbg_read(buf,512)
t=list(buf)
base=t[0]
names={0:'cw loop top',1:'s_full ok',2:'S loaded',3:'dP loaded+sdp_free',4:'math done',5:'mma2_done ok',6:'dq drained',7:'stores+ps_ready',8:'mma: S/dP(i+1) issued',9:'mma: ps_ready ok',10:'mma: dq_free ok',11:'mma: dV,dK,dQ issued'}
for i in range(2,6):
    print('--- i',i)
    for sl in range(12):
        val=t[16*i+sl]
        if val: print(f'   {names[sl]:26s} {val-base:8d}')
