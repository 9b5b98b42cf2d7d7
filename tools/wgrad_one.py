import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow.b200 import ops
dev=torch.device('cuda'); B=256
HW=int(sys.argv[1]) if len(sys.argv)>1 else 32   # feature-map width: 32 (level 0) ... 4
x=torch.randn(B,100,HW,HW,device=dev); g=torch.randn(B,100,HW,HW,device=dev)
for _ in range(4):
    gw=ops.conv_wgrad_tc(g,x,3,True)
torch.cuda.synchronize(); print('ok', float(gw.abs().mean()))
