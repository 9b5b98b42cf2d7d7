"""Per-level timing of ConvResidualNet forward+backward: tensor-core path vs cuDNN fp32."""
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow import nn as nn_
from manifold_flow.b200 import config, ops
dev=torch.device('cuda'); B=int(sys.argv[1]) if len(sys.argv)>1 else 256
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/iters
for (cid,hw) in ((6,32),(12,16),(24,8),(48,4)):
    torch.manual_seed(0)
    net=nn_.ConvResidualNet(cid,cid*32,hidden_channels=100,num_blocks=2).to(dev)
    x=torch.randn(B,cid,hw,hw,device=dev,requires_grad=True); gy=torch.randn(B,cid*32,hw,hw,device=dev)
    def step():
        net.zero_grad(set_to_none=True); y=net(x); y.backward(gy)
    def fwd():
        with torch.no_grad(): net(x)
    res={}
    for flag in (True,False):
        config.conv_tensor_cores=flag
        res[flag]=(timeit(fwd),timeit(step))
    config.conv_tensor_cores=True
    print(f'level HW={hw} C_id={cid} B={B}: fwd TC {res[True][0]:.3f} ms vs cuDNN {res[False][0]:.3f} ms | fwd+bwd TC {res[True][1]:.3f} ms vs cuDNN {res[False][1]:.3f} ms')
