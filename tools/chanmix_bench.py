"""chanmix (fused ActNorm / permutation / LU 1x1 conv) forward: CUDA-event timing, algorithmic GB/s (8 B per element)."""
import sys, os, json
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow.b200 import ops
dev=torch.device('cuda'); peak=6541.5
for (b,c,h) in ((256,12,32),(256,24,16),(256,48,8),(256,96,4)):
    nbuf=max(2,int(300e6/(b*c*h*h*8)))
    xs=[torch.randn(b,c,h,h,device=dev) for _ in range(nbuf)]
    m=torch.randn(c,c,device=dev); bias=torch.randn(c,device=dev)
    ref=torch.einsum('ck,bkhw->bchw',m.double(),xs[0].double())+bias.double()[None,:,None,None]
    err=((ops.chanmix(xs[0],m,bias).double()-ref).norm()/ref.norm()).item()
    # device time: the 20 launches are captured into a CUDA graph (eager calls are bound by ~30 us of host work each)
    side=torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): ops.chanmix(xs[0],m,bias)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(20): ops.chanmix(xs[i%nbuf],m,bias)
    gr.replay(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); t=e0.elapsed_time(e1)/20
    by=b*c*h*h*8
    print(f'[{b},{c},{h},{h}]: {t*1e3:.1f} us  {by/t/1e6:.0f} GB/s ({by/t/1e6/peak:.2f} of {peak:.0f})  rel err vs fp64 {err:.1e}')
