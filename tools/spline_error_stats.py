"""Error statistics of the spline kernels against the oracle in fp32 (same algorithm on CPU) and fp64."""
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests'): sys.path.insert(0, os.path.join(R,p))
import numpy as np, torch, mflow_oracle as o
from manifold_flow.b200 import ops
dev=torch.device('cuda')
def stats(tag, a, b):
    d=(a.double()-b.double()).abs()
    rel=d/(b.double().abs()+1e-30)
    print(f'   {tag}: max abs {d.max().item():.3e}  p99.9 abs {d.flatten().kthvalue(max(1,int(d.numel()*0.999)))[0].item():.3e}  mean {d.mean().item():.3e}  #>1e-5*(1+|ref|): {(d>1e-5*(1+b.double().abs())).sum().item()}/{d.numel()}')
for k,t,scale in ((8,3.0,1.5),(11,10.0,1.5),(11,10.0,0.3),(10,6.0,30.0)):
    g=torch.Generator().manual_seed(5+k)
    n=40013
    x=torch.randn(n,generator=g)*(t/2)
    params=torch.randn(n,3*k-1,generator=g)*scale
    for inv in (False,True):
        r64=o.rqs(x.double(),params[:,:k].double(),params[:,k:2*k].double(),params[:,2*k:].double(),inv,t,return_bins=True)
        r32=o.rqs(x,params[:,:k],params[:,k:2*k],params[:,2*k:],inv,t,return_bins=True)
        y,lad,bins=ops.rqs_elementwise(x.to(dev),params.to(dev),k,t,inverse=inv,return_bins=True)
        y,lad,bins=y.cpu(),lad.cpu(),bins.cpu().long()
        print(f'K={k} T={t} scale={scale} inv={inv}: bin mismatches vs ref32 {(bins!=r32[2]).sum().item()} vs ref64 {(bins!=r64[2]).sum().item()}; ref32 vs ref64 {(r32[2]!=r64[2]).sum().item()}; bitwise-equal y vs ref32: {(y==r32[0]).float().mean().item():.4f}, lad {(lad==r32[1]).float().mean().item():.4f}')
        stats('y   gpu-ref32',y,r32[0]); stats('y   gpu-ref64',y,r64[0]); stats('y   ref32-ref64',r32[0],r64[0])
        stats('lad gpu-ref32',lad,r32[1]); stats('lad gpu-ref64',lad,r64[1]); stats('lad ref32-ref64',r32[1],r64[1])
