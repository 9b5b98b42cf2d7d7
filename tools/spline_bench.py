"""Micro-benchmark of the fused RQ-spline kernels (SURVEY 8(d) shapes): CUDA-event timing of
mflow_rqs_apply / mflow_rqs_backward alone, algorithmic GB/s against the measured HBM peak."""
import sys, os, json
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow.b200 import ops
dev=torch.device('cuda')
peak=6541.5
try: peak=json.load(open(os.path.join(R,'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception: pass
quick = len(sys.argv)>1 and sys.argv[1]=='quick'
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
K,T=11,10.0; P=3*K-1
shapes=[(256,12,32,32)] if quick else [(256,12,32,32),(256,24,16,16),(256,48,8,8),(256,96,4,4),(25,12,32,32)]
g=torch.Generator(device=dev).manual_seed(0)
for (b,c,h,w) in shapes:
    ct=c//2
    # several distinct buffers so that consecutive iterations do not hit in the 126 MB L2
    nbuf=max(1,int(400e6/(b*ct*P*h*w*4))+1) if not quick else 2
    xs=[torch.randn(b,c,h,w,device=dev,generator=g)*5 for _ in range(nbuf)]
    ps=[torch.randn(b,ct*P,h,w,device=dev,generator=g) for _ in range(nbuf)]
    geom=ops.SplineGeometry(ct,0,1,c-ct,ct,1)
    n=b*ct*h*w
    i=[0]
    def fwd(inv=False):
        j=i[0]%nbuf; i[0]+=1
        return ops.rqs_coupling(xs[j],ps[j],geom,K,T,1e-3,1e-3,1e-3,10.0,inv)
    t_f=timeit(lambda: fwd(False)); t_i=timeit(lambda: fwd(True))
    xg=[x.clone().requires_grad_(True) for x in xs]; pg=[p.clone().requires_grad_(True) for p in ps]
    outs=[ops.rqs_coupling(xg[j],pg[j],geom,K,T,1e-3,1e-3,1e-3,10.0,False) for j in range(nbuf)]
    gy=torch.randn_like(xs[0]); gl=torch.randn(b,device=dev)
    def bwd():
        j=i[0]%nbuf; i[0]+=1
        torch.autograd.grad([outs[j][0],outs[j][1]],[xg[j],pg[j]],[gy,gl],retain_graph=True)
    t_b=timeit(bwd)
    bf=n*4*(P+2); bb=n*4*(2*P+3)
    print(f'[{b},{c},{h},{w}] K={K}: fwd {t_f*1e3:.1f} us {bf/t_f/1e6:.0f} GB/s ({bf/t_f/1e6/peak:.2f}) | inv {t_i*1e3:.1f} us {bf/t_i/1e6:.0f} GB/s ({bf/t_i/1e6/peak:.2f}) | bwd {t_b*1e3:.1f} us {bb/t_b/1e6:.0f} GB/s ({bb/t_b/1e6/peak:.2f})  [bufs {nbuf}]')
if not quick:
    for K2,nt,D in ((8,1,2),(10,1,3),(11,20,40),(11,32,64)):
        P2=3*K2-1; b=1<<20 if nt<=2 else 1<<17
        x=torch.randn(b,D,device=dev)*2; p=torch.randn(b,nt*P2,device=dev)
        geom=ops.SplineGeometry(nt,0,2,D-nt,1,2) if nt*2<=D else ops.SplineGeometry(nt,0,1,D-nt,nt,1)
        t=timeit(lambda: ops.rqs_coupling(x,p,geom,K2,3.0,1e-3,1e-3,1e-3,10.0,False))
        n=b*nt; bf=n*4*(P2+2)
        print(f'rows [B={b},n_t={nt}] K={K2}: fwd {t*1e3:.1f} us {bf/t/1e6:.0f} GB/s ({bf/t/1e6/peak:.2f})')
