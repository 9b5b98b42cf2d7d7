import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import numpy as np, torch
import golden_util as gu, mflow_oracle as oracle, model_util as mu
from manifold_flow.b200 import ops, config
import manifold_flow.transforms.coupling as coupling_mod
name=sys.argv[1] if len(sys.argv)>1 else 'cfg4s'
dev=torch.device('cuda')
model,_,sd=mu.seeded_product_model(name); model.to(dev)
x,ctx=gu.config_inputs(name)
om=mu.oracle_model(name)
def oracle_grads(dtype):
    cast=lambda v: v.detach().clone().to(dtype) if torch.is_floating_point(v) else v
    s={k:(cast(v).requires_grad_(True) if torch.is_floating_point(v) and k.split('.')[-1] not in ('_shift','_scale') else cast(v)) for k,v in sd.items()}
    if om['kind']=='image_flow': lp=oracle.flow_forward(om,s,cast(x),None,decode=False)[1]
    else: lp=oracle.mflow_forward(om,s,cast(x),'mf-fixed-manifold',None if ctx is None else cast(ctx))[1]
    (-lp.mean()).backward()
    return {k:v.grad for k,v in s.items() if torch.is_floating_point(v) and v.requires_grad and v.grad is not None}
g64=oracle_grads(torch.float64); g32=oracle_grads(torch.float32)
def product_grads():
    model.zero_grad(set_to_none=True)
    xd=x.to(dev); cd=None if ctx is None else ctx.to(dev)
    lp = model(xd)[1] if om['kind']=='image_flow' else model(xd,mode='mf-fixed-manifold',context=cd)[1]
    (-lp.mean()).backward()
    return {k:p.grad.cpu() for k,p in model.named_parameters() if p.grad is not None}
def report(tag,g):
    errs=[]
    for k,v in g64.items():
        if k not in g: continue
        sc=float(v.norm())+1e-30
        errs.append((float((g[k].double()-v).norm())/sc, float((g32[k].double()-v).norm())/sc, k))
    errs.sort(reverse=True)
    e=np.array([a for a,_,_ in errs]); r=np.array([b for _,b,_ in errs])
    print(f'{tag}: rel-L2 grad error vs fp64: median {np.median(e):.2e} p90 {np.quantile(e,0.9):.2e} max {e.max():.2e} | oracle fp32 CPU: median {np.median(r):.2e} p90 {np.quantile(r,0.9):.2e} max {r.max():.2e}')
    for a,b,k in errs[:6]: print(f'      {a:.2e} (ref {b:.2e}) {k}')
report('product', product_grads())
# variant: spline via oracle autograd on the GPU (isolates the spline kernels)
orig=ops.rqs_coupling
def rqs_via_oracle(xx,params,geom,K,T,mw,mh,md,div,inverse):
    b=xx.shape[0]; ct=geom.n_transform
    tr=slice(geom.t_offset,None,geom.t_step); idn=slice(geom.i_offset,None,geom.i_step)
    xt=xx[:,tr][:,:ct]
    if xx.dim()==4:
        hh,ww=xx.shape[2:]; pv=params.reshape(b,ct,-1,hh,ww).permute(0,1,3,4,2)
    else: pv=params.reshape(b,ct,-1)
    yt,lad=oracle.rqs(xt,pv[...,:K]/div,pv[...,K:2*K]/div,pv[...,2*K:],inverse,T)
    out=torch.empty_like(xx); idx_t=torch.arange(xx.shape[1],device=xx.device)[tr][:ct]; idx_i=torch.arange(xx.shape[1],device=xx.device)[idn][:geom.n_identity]
    out=out.index_copy(1,idx_t,yt).index_copy(1,idx_i,xx[:,idn][:,:geom.n_identity])
    return out, oracle.sum_except_batch(lad)
coupling_mod.ops.rqs_coupling=rqs_via_oracle
report('spline=oracle-autograd-on-GPU', product_grads())
coupling_mod.ops.rqs_coupling=orig
config.fuse_linear_glue=False
report('unfused glue', product_grads())
config.fuse_linear_glue=True
