"""One bench step (cfg4, log_prob + backward) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ..."""
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
import golden_util as gu
from manifold_flow.b200 import factory
import bench
b=int(sys.argv[1]) if len(sys.argv)>1 else 256
dev=torch.device('cuda')
torch.manual_seed(gu.SEED)
model=factory.create_model('cfg4').to(dev); model.train()
x,c=bench.synthetic_batch('cfg4',b,0); x,c=x.to(dev),c.to(dev)
with torch.no_grad(): model(x,mode='projection',context=c)
def step():
    model.zero_grad(set_to_none=True)
    lp=model(x,mode='mf-fixed-manifold',context=c)[1]
    (-lp.mean()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
