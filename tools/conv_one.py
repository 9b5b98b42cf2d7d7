import sys, os, ctypes
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow.b200 import ops
dev=torch.device('cuda'); B=256
HW=int(sys.argv[1]) if len(sys.argv)>1 else 32   # feature-map width: 32 (level 0) ... 4
x=torch.randn(B,100,HW,HW,device=dev); w=torch.randn(100,100,3,3,device=dev)*0.1; b=torch.randn(100,device=dev)
for _ in range(4):
    y=ops.conv_tc(x,w,b,relu_in=True)
torch.cuda.synchronize(); print('ok', float(y.abs().mean()))
