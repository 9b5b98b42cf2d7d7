"""Role cycle counters of the tcgen05 convolution (MFLOW_CONV_DEBUG=8, CTA (0,0)); level-0 3x3 at batch 256."""
import sys, os, ctypes
os.environ['MFLOW_CONV_DEBUG'] = str(int(os.environ.get('MFLOW_CONV_DEBUG', '0')) | 8)
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow.b200.cabi import ConvDesc, lib, ptr, stream, check
dev=torch.device('cuda')
HW=int(sys.argv[1]) if len(sys.argv)>1 else 32; k=int(sys.argv[2]) if len(sys.argv)>2 else 3; co=int(sys.argv[3]) if len(sys.argv)>3 else 100
B=256; ci=100
x=torch.randn(B,ci,HW,HW,device=dev); w=torch.randn(co,ci,k,k,device=dev)*0.1; b=torch.randn(co,device=dev); res=torch.randn(B,co,HW,HW,device=dev)
n_pad=(co+127)//128*128; k_pad=(ci+31)//32*32
wt=torch.zeros(k*k,n_pad,k_pad,device=dev); wt[:,:co,:ci]=w.permute(2,3,0,1).reshape(k*k,co,ci)
hi=(wt.view(torch.int32)&-8192).view(torch.float32).contiguous(); lo=(((wt-hi).view(torch.int32)+4096)&-8192).view(torch.float32).contiguous()
d=ConvDesc(batch=B,c_in=ci,c_out=co,height=HW,width=HW,kernel=k,relu_in=1,relu_out=0,n_pad=n_pad,k_pad=k_pad)
y=torch.empty(B,co,HW,HW,device=dev)
names=['mma total','mma wait b_full','mma wait a_full','mma wait acc_empty','xf total','xf wait raw_full','xf wait a_empty','xf transform','xf wait acc_full','xf epilogue total','tma total','tma wait raw_empty','tma wait b_empty','probe: 1 raw box','probe: n_raw raw boxes','probe: TMA weight tile']
out=(ctypes.c_ulonglong*16)()
L=lib(); L.mflow_debug_conv_profile.restype=ctypes.c_int; L.mflow_debug_conv_profile.argtypes=[ctypes.c_void_p]
for aux in (None, res):
    for _ in range(3): check(L.mflow_conv2d_tf32x3(ctypes.byref(d),ptr(x),ptr(hi),ptr(lo),ptr(b),ptr(aux),None,ptr(y),stream()),'conv')
    torch.cuda.synchronize(); L.mflow_debug_conv_profile(out)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); check(L.mflow_conv2d_tf32x3(ctypes.byref(d),ptr(x),ptr(hi),ptr(lo),ptr(b),ptr(aux),None,ptr(y),stream()),'conv'); e1.record()
    torch.cuda.synchronize(); L.mflow_debug_conv_profile(out)
    print(f'--- aux={"residual" if aux is not None else "none"}: {e0.elapsed_time(e1)*1000:.0f} us')
    for n,v in zip(names,out): print(f'  {n:26s} {v/1e3:9.2f} kcycles')
