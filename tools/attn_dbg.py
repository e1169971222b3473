import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
from cinema_b200 import _C
DEV='cuda'; BF=torch.bfloat16
B,nq,nk,h,d=16,2053,684,16,32
q=torch.randn(B,nq,h,d,device=DEV).to(BF); k=torch.randn(B,nk,h,d,device=DEV).to(BF); v=torch.randn(B,nk,h,d,device=DEV).to(BF)
o=torch.empty_like(q); lse=torch.empty(B,h,nq,device=DEV)
for _ in range(3): _C.attention_fwd(q,k,v,o,lse,d**-0.5)
torch.cuda.synchronize()
buf=(ctypes.c_longlong*128)()
_C.lib().cb_attn_dbg_read(buf,128)
t=list(buf)
base=t[0]
names={0:'wg0 loop top',1:'s_full ok',2:'ldtm done',3:'max done',4:'pv_done ok',5:'exp+store done',6:'p_full arrived',8:'mma: before v_full',9:'mma: p_full0 ok',10:'mma: PV0 issued',11:'mma: p_full1 ok',12:'mma: PV1 issued'}
for j in range(6):
    print('--- j',j)
    for sl in (0,1,2,3,4,5,6,8,9,10,11,12):
        val=t[16*j+sl]
        if val: print(f'   {names[sl]:22s} {val-base:8d}')
