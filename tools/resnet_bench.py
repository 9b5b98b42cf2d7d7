"""Vector conditioner (ResidualNet, reference nn/resnet.py:43-72) forward + backward: tensor-core planes path vs cuBLAS path."""
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow import nn as nn_
from manifold_flow.b200.config import config
dev=torch.device('cuda')
def timeit(fn,iters=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/iters
for rows,n_in,n_out,ctx,name in ((1<<16,20,640,None,'cfg3 outer'),(1<<16,7,224,2,'cfg3 inner'),(1<<20,1,23,None,'cfg1'),(1<<14,33,1024,1,'cfg4 inner')):
    torch.manual_seed(0)
    net=nn_.ResidualNet(n_in,n_out,hidden_features=100,context_features=ctx,num_blocks=2).to(dev)
    x=torch.randn(rows,n_in,device=dev,requires_grad=True); c=None if ctx is None else torch.randn(rows,ctx,device=dev); g=torch.randn(rows,n_out,device=dev)
    def step():
        net.zero_grad(set_to_none=True); x.grad=None
        net(x,c).backward(g)
    res={}
    for thr in (1,1<<30):
        config.linear_tc_min_rows=thr; res[thr]=timeit(step)
    print(f'{name}: rows {rows} {n_in}->{n_out} ctx {ctx}: fwd+bwd tensor cores {res[1]:.3f} ms | cuBLAS fp32 {res[1<<30]:.3f} ms')
