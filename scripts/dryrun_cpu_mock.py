import sys, types, torch
import torch.nn.functional as F
m = types.ModuleType("multimae_b200.kernels")
def gemm(A,B,*,a_mn=False,b_mn=False,bias=None,act=0,residual=None,dgelu_z=None,preact=None,out_f32=None,out_bf16=None,accumulate=False,split_k=1,alpha=1.0):
    Af = A.float().t() if a_mn else A.float(); Bf = B.float().t() if b_mn else B.float()
    v = alpha*(Af @ Bf.t())
    if bias is not None: v = v + bias
    if preact is not None: preact.copy_(v)
    if act==1: v = F.gelu(v)
    if dgelu_z is not None:
        z = dgelu_z.float().requires_grad_(True); F.gelu(z).sum().backward(); v = v*z.grad
    if residual is not None: v = v+residual
    if out_f32 is not None:
        if accumulate or split_k>1: out_f32.add_(v)
        else: out_f32.copy_(v)
    if out_bf16 is not None: out_bf16.copy_(v)
m.gemm=gemm
m.cast_bf16=lambda x,dst=None: x.to(torch.bfloat16)
def cast_colsum(src,dst=None,colsum=None):
    if dst is not None: dst.copy_(src)
    if colsum is not None: colsum.add_(src.sum(0))
m.cast_colsum=cast_colsum
m.colsum_bf16=lambda s,c: c.add_(s.float().sum(0))
m.transpose_bf16=lambda s,d=None: s.t().contiguous()
def layernorm_fwd(x,g,b,eps=1e-6,out_bf16=True,out_f32=False):
    y=F.layer_norm(x,(x.shape[1],),g,b,eps); mean=x.mean(1); rstd=(x.var(1,unbiased=False)+eps).rsqrt()
    return y.to(torch.bfloat16), y, mean, rstd
m.layernorm_fwd=layernorm_fwd
def layernorm_bwd(dy,x,mean,rstd,gamma,dgamma,dbeta,dx_resid=None,dx=None):
    xr=x.clone().requires_grad_(True); g=gamma.clone().requires_grad_(True); b=torch.zeros_like(gamma).requires_grad_(True)
    F.layer_norm(xr,(x.shape[1],),g,b,1e-6).backward(dy.float())
    if dgamma is not None: dgamma.add_(g.grad); dbeta.add_(b.grad)
    return xr.grad + (dx_resid if dx_resid is not None else 0)
m.layernorm_bwd=layernorm_bwd
def _att(q,k,v,B,H,Nq,Nk,dh,scale):
    qf=q.float().reshape(B,Nq,H,dh).transpose(1,2); kf=k.float().reshape(B,Nk,H,dh).transpose(1,2); vf=v.float().reshape(B,Nk,H,dh).transpose(1,2)
    return qf,kf,vf
def attention_fwd(q,k,v,B,H,Nq,Nk,dh,scale,out=None):
    qf,kf,vf=_att(q,k,v,B,H,Nq,Nk,dh,scale); s=(qf@kf.transpose(-2,-1))*scale
    o=(torch.softmax(s,-1)@vf).transpose(1,2).reshape(B*Nq,H*dh).to(torch.bfloat16)
    return o, torch.logsumexp(s,-1)
m.attention_fwd=attention_fwd
def attention_bwd(q,k,v,o,do,lse,dq,dk,dv,B,H,Nq,Nk,dh,scale):
    qf,kf,vf=[t.detach().requires_grad_(True) for t in _att(q,k,v,B,H,Nq,Nk,dh,scale)]
    s=(qf@kf.transpose(-2,-1))*scale
    (torch.softmax(s,-1)@vf).transpose(1,2).reshape(B*Nq,H*dh).backward(do.float())
    dq.copy_(qf.grad.transpose(1,2).reshape(B*Nq,-1)); dk.copy_(kf.grad.transpose(1,2).reshape(B*Nk,-1)); dv.copy_(vf.grad.transpose(1,2).reshape(B*Nk,-1))
m.attention_bwd=attention_bwd
class _FakeLib:
    def mmae_gemm_set_variant(self, v): return 0
    def mmae_gemm_set_tma_store(self, v): return 0
    def mmae_set_pdl(self, v): return 0
    def mmae_set_wgrad_stream(self, v): return 0
    def mmae_weight_mirror_register(self, *a): return 0
    def mmae_attention_set_tc(self, v): return 0
libmod = types.ModuleType("multimae_b200._lib"); libmod.lib = lambda: _FakeLib()
pkg = types.ModuleType("multimae_b200"); pkg.kernels = m; pkg._lib = libmod; pkg.__path__=[]
sys.modules["multimae_b200._lib"]=libmod
sys.modules["multimae_b200"]=pkg; sys.modules["multimae_b200.kernels"]=m
_dev = torch.device
torch.device = lambda *a, **k: _dev("cpu")
class _E:
    def __init__(s,**k): pass
    def record(s): pass
    def elapsed_time(s,o): return 1.0
torch.cuda.Event=_E; torch.cuda.synchronize=lambda: None
src=open("scripts/gpu_check_gemm.py").read()
# shrink timing shapes for CPU
src=src.replace("12672","128").replace("25088","128").replace("12544","128").replace("(2560, 2304, 768)","(256, 264, 128)").replace("B_, H_, N_, dh_ = 128, 12, 99, 64","B_, H_, N_, dh_ = 2, 2, 9, 64").replace("iters=20","iters=1")
exec(compile(src,"gpu_check_gemm.py","exec"))
