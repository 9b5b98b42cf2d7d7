"""Standalone check + timing of the tcgen05 3xTF32 weight-gradient kernel against fp64 autograd and cuDNN fp32."""
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch, torch.nn.functional as F
from manifold_flow.b200 import ops
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
dev=torch.device('cuda'); torch.manual_seed(0)
def ref64(g,x,k,relu):
    w=torch.zeros(g.shape[1],x.shape[1],k,k,device=dev,dtype=torch.float64,requires_grad=True)
    xin=(F.relu(x) if relu else x).double()
    y=F.conv2d(xin,w,padding=k//2); y.backward(g.double()); return w.grad
cases=[(2,100,100,32,3,True),(3,100,100,16,3,True),(5,100,100,8,3,False),(2,6,100,32,1,False),(2,100,192,32,1,True),(3,100,384,16,1,True),(2,24,100,8,1,False),(1,8,8,32,3,False),(2,130,40,16,1,False),(6,100,100,4,3,True),(5,48,100,4,1,False),(3,100,1536,4,1,True)]
if len(sys.argv)>1: cases=[cases[int(sys.argv[1])]]
TIMING=len(sys.argv)<=1
for (B,ci,co,HW,k,relu) in cases:
    x=torch.randn(B,ci,HW,HW,device=dev); g=torch.randn(B,co,HW,HW,device=dev)
    try:
        gw=ops.conv_wgrad_tc(g,x,k,relu); torch.cuda.synchronize()
    except Exception as e:
        print('case',(B,ci,co,HW,k),'FAILED',repr(e)[:300]); break
    want=ref64(g,x,k,relu)
    err=(gw.double()-want).abs().max().item(); rel=((gw.double()-want).norm()/want.norm()).item()
    xin=F.relu(x) if relu else x
    lib=torch.ops.aten.convolution_backward(g,xin,torch.zeros(co,ci,k,k,device=dev),None,[1,1],[k//2]*2,[1,1],False,[0,0],1,[False,True,False])[1]
    print(f'case B={B} ci={ci} co={co} HW={HW} k={k} relu={relu}: max abs err {err:.3e} rel L2 {rel:.3e} nan={torch.isnan(gw).sum().item()} | library fp32 rel {((lib.double()-want).norm()/want.norm()).item():.3e}')
def timeit(fn,iters=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/iters*1e3
for (B,ci,co,HW,k) in [] if not TIMING else [(256,100,100,32,3),(256,100,192,32,1),(256,6,100,32,1),(256,100,100,16,3),(256,100,384,16,1),(256,100,100,8,3),(256,100,100,4,3),(256,100,1536,4,1)]:
    x=torch.randn(B,ci,HW,HW,device=dev); g=torch.randn(B,co,HW,HW,device=dev); w0=torch.zeros(co,ci,k,k,device=dev)
    t_tc=timeit(lambda: ops.conv_wgrad_tc(g,x,k,True))
    t_lib=timeit(lambda: torch.ops.aten.convolution_backward(g,F.relu(x),w0,None,[1,1],[k//2]*2,[1,1],False,[0,0],1,[False,True,False]))
    fl=2*B*HW*HW*ci*co*k*k
    print(f'wgrad B={B} {ci}->{co} {HW}x{HW} k={k}: tcgen05 {t_tc:.0f} us ({fl/t_tc/1e6:.1f} TFLOP/s useful) | library fp32 (+relu) {t_lib:.0f} us')
