"""torch.profiler summary of one bench step (device time per kernel) -> stdout table."""
import sys, os, json
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from torch.profiler import profile, ProfilerActivity
import golden_util as gu
from manifold_flow.b200 import factory, cabi
import bench
b=int(sys.argv[1]) if len(sys.argv)>1 else 64
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
n0=cabi.launch_count
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print('mflow launches per step', cabi.launch_count-n0)
ka=prof.key_averages()
rows=[(e.key, e.count, e.device_time_total/1e3) for e in ka if e.device_time_total>0 and e.device_type.name=='CUDA']
rows.sort(key=lambda r:-r[2])
tot=sum(r[2] for r in rows)
print(f'total device kernel time {tot:.2f} ms, kernels launched {sum(r[1] for r in rows)}')
for k,cnt,ms in rows[:45]:
    print(f'{ms:9.3f} ms {100*ms/tot:5.1f}% x{cnt:5d}  {k[:150]}')
import time
t=time.perf_counter(); 
for _ in range(5): step()
torch.cuda.synchronize(); print('wall ms/step', (time.perf_counter()-t)/5*1e3)
