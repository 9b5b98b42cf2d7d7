import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests'): sys.path.insert(0, os.path.join(R,p))
import numpy as np, torch, mflow_oracle as o, golden_util as gu
from test_oracle_golden import _layer_spec
from test_gpu_layers import load_case
from manifold_flow.b200 import ops
L=gu.load_npz('layers.npz')
dev=torch.device('cuda')
for name in ['vcoup_D64_K11_T10_ctx1_odd','vcoup_D64_K11_T10_ctx1_even','vcoup_D40_K11_T10_ctx0_odd']:
    layer, sd, x, ctx = load_case(name, dev)
    y=torch.from_numpy(L[name+'/y'])
    K=11;T=10.0
    idf=sd['identity_features']; tf=sd['transform_features']
    cpu_params=o.residual_net(sd,'transform_net.',y[:,idf],ctx)
    with torch.no_grad():
        gpu_params=layer.transform_net(y.to(dev)[:,idf], None if ctx is None else ctx.to(dev)).cpu()
    print(name,'params max abs diff', (cpu_params-gpu_params).abs().max().item(), 'max', cpu_params.abs().max().item())
    P=3*K-1
    pv=cpu_params.reshape(y.shape[0], len(tf), P)
    xe=y[:,tf].reshape(-1)
    pe=pv.reshape(-1,P).clone(); pe[:,:2*K]/=4.0  # hidden 16
    yy,ll,bb=ops.rqs_elementwise(xe.to(dev), pe.to(dev), K, T, inverse=True, return_bins=True)
    ry,rl,rb=o.rqs(xe.double(), pe[:,:K].double(), pe[:,K:2*K].double(), pe[:,2*K:].double(), True, T, return_bins=True)
    ry32,rl32,rb32=o.rqs(xe, pe[:,:K], pe[:,K:2*K], pe[:,2*K:], True, T, return_bins=True)
    dl=(ll.cpu().double()-rl).abs(); dl32=(rl32.double()-rl).abs()
    i=int(dl.argmax())
    print('  kernel-only: max lad err vs fp64', dl.max().item(), 'ref32 err', dl32.max().item(), 'at', i, 'bins', bb[i].item(), rb[i].item(), 'x', xe[i].item(), 'lad', rl[i].item(), 'y err', (yy.cpu().double()-ry).abs().max().item())
    print('  params at worst', pe[i])
