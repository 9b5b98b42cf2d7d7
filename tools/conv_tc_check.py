"""Standalone check of the tcgen05 3xTF32 convolution kernel against fp64 F.conv2d."""
import sys, os, ctypes
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch, torch.nn.functional as F
from manifold_flow.b200 import cabi
from manifold_flow.b200.cabi import ConvDesc, lib, ptr, stream, check
if os.environ.get('MFLOW_LIB'): cabi._LIB_PATH = os.environ['MFLOW_LIB']  # tooling only: time another build of the library
dev=torch.device('cuda')
def prep(w):
    co,ci,kh,kw=w.shape
    n_pad=(co+127)//128*128; k_pad=(ci+31)//32*32
    wt=torch.zeros(kh*kw,n_pad,k_pad,device=w.device)
    wt[:,:co,:ci]=w.permute(2,3,0,1).reshape(kh*kw,co,ci)
    hi=(wt.view(torch.int32)&-8192).view(torch.float32)
    lo=(((wt-hi).view(torch.int32)+4096)&-8192).view(torch.float32)
    return hi.contiguous(),lo.contiguous(),n_pad,k_pad
def conv_tc(x,w,b=None,res=None,relu_in=False,relu_out=False):
    hi,lo,n_pad,k_pad=prep(w)
    B,ci,H,W=x.shape; co=w.shape[0]
    d=ConvDesc(batch=B,c_in=ci,c_out=co,height=H,width=W,kernel=w.shape[2],relu_in=int(relu_in),relu_out=int(relu_out),n_pad=n_pad,k_pad=k_pad)
    y=torch.full((B,co,H,W),float('nan'),device=x.device)
    check(lib().mflow_conv2d_tf32x3(ctypes.byref(d),ptr(x),ptr(hi),ptr(lo),ptr(b),ptr(res),None,ptr(y),stream()),'conv')
    return y
torch.manual_seed(0)
cases=[] if (os.environ.get('MFLOW_CONV_DEBUG') or os.environ.get('MFLOW_CONV_TIMING_ONLY')) else [(2,100,100,32,3),(2,100,100,16,3),(4,100,100,8,3),(9,100,100,4,3),(2,6,100,32,1),(3,100,192,32,1),(2,100,384,16,1),(2,48,100,4,1),(1,8,8,32,3)]
if len(sys.argv)>1: cases=cases[:int(sys.argv[1])]
for (B,ci,co,HW,k) in cases:
    x=torch.randn(B,ci,HW,HW,device=dev); w=torch.randn(co,ci,k,k,device=dev)*0.1; b=torch.randn(co,device=dev); r=torch.randn(B,co,HW,HW,device=dev)
    try:
        y=conv_tc(x,w,b,r,relu_in=True,relu_out=False)
        torch.cuda.synchronize()
    except Exception as e:
        print('case',(B,ci,co,HW,k),'FAILED',repr(e)[:200]); break
    ref=F.conv2d(F.relu(x.double()),w.double(),b.double(),padding=k//2)+r.double()
    err=(y.double()-ref).abs().max().item(); rel=((y.double()-ref).norm()/ref.norm()).item()
    y1=F.conv2d(F.relu(x),w,b,padding=k//2)+r
    print(f'case B={B} ci={ci} co={co} HW={HW} k={k}: max abs err {err:.3e} rel L2 {rel:.3e} nan={torch.isnan(y).sum().item()} | cudnn fp32 rel {((y1.double()-ref).norm()/ref.norm()).item():.3e}')

# timing of the level-0 3x3 convolution at batch 256 against cuDNN fp32 (TF32 off)
torch.backends.cudnn.allow_tf32=False
B=256; x=torch.randn(B,100,32,32,device=dev); w=torch.randn(100,100,3,3,device=dev)*0.1; b=torch.randn(100,device=dev)
hi,lo,n_pad,k_pad=prep(w)
d=ConvDesc(batch=B,c_in=100,c_out=100,height=32,width=32,kernel=3,relu_in=1,relu_out=0,n_pad=n_pad,k_pad=k_pad)
y=torch.empty(B,100,32,32,device=dev)
def run_tc(): check(lib().mflow_conv2d_tf32x3(ctypes.byref(d),ptr(x),ptr(hi),ptr(lo),ptr(b),None,None,ptr(y),stream()),'conv')
def run_cudnn(): return F.conv2d(F.relu(x),w,b,padding=1)
for name,fn in (('tcgen05 3xTF32',run_tc),('cuDNN fp32 (+relu kernel)',run_cudnn)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
    print(f'{name}: {ms*1e3:.0f} us  -> {2*B*1024*900*100/ms/1e9:.1f} TFLOP/s (useful fp32-equivalent)')
res=torch.randn(B,100,32,32,device=dev); gate=torch.randn(B,100,32,32,device=dev)
def run_v(r,g,relu):
    d.relu_in=int(relu)
    check(lib().mflow_conv2d_tf32x3(ctypes.byref(d),ptr(x),ptr(hi),ptr(lo),ptr(b),ptr(r),ptr(g),ptr(y),stream()),'conv')
for name,(r,g,relu) in (('plain',(None,None,True)),('+residual',(res,None,True)),('+gate (dgrad)',(None,gate,False))):
    for _ in range(3): run_v(r,g,relu)
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run_v(r,g,relu)
    e1.record(); torch.cuda.synchronize(); print(f'variant {name}: {e0.elapsed_time(e1)/10*1e3:.0f} us')
# cold-cache variant: rotate over 6 different input tensors (630 MB) so x never sits in L2
xs=[torch.randn(B,100,32,32,device=dev) for _ in range(6)]
i=[0]
def run_cold():
    i[0]+=1
    check(lib().mflow_conv2d_tf32x3(ctypes.byref(d),ptr(xs[i[0]%6]),ptr(hi),ptr(lo),ptr(b),None,None,ptr(y),stream()),'conv')
d.relu_in=1
for _ in range(3): run_cold()
torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(12): run_cold()
e1.record(); torch.cuda.synchronize(); print(f'variant cold inputs: {e0.elapsed_time(e1)/12*1e3:.0f} us')
