"""Per-step timeline of one CTA of the tcgen05 convolution (MFLOW_CONV_DEBUG=8): cycles relative to kernel start."""
import sys, os, ctypes
os.environ['MFLOW_CONV_DEBUG'] = str(8 + 256 * int(sys.argv[3] if len(sys.argv) > 3 else 0))   # traced CTA index in the high bits
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('manifold-flow_b200','oracle','tests',''): sys.path.insert(0, os.path.join(R,p))
import torch
from manifold_flow.b200.cabi import ConvDesc, lib, ptr, stream, check
dev=torch.device('cuda')
B=int(sys.argv[1]) if len(sys.argv)>1 else 8; HW=int(sys.argv[2]) if len(sys.argv)>2 else 4
ci=co=100; k=3
x=torch.randn(B,ci,HW,HW,device=dev); w=torch.randn(co,ci,k,k,device=dev)*0.1; b=torch.randn(co,device=dev)
n_pad=128; k_pad=128
wt=torch.zeros(k*k,n_pad,k_pad,device=dev); wt[:,:co,:ci]=w.permute(2,3,0,1).reshape(k*k,co,ci)
hi=(wt.view(torch.int32)&-8192).view(torch.float32).contiguous(); lo=(((wt-hi).view(torch.int32)+4096)&-8192).view(torch.float32).contiguous()
d=ConvDesc(batch=B,c_in=ci,c_out=co,height=HW,width=HW,kernel=k,relu_in=1,relu_out=0,n_pad=n_pad,k_pad=k_pad)
y=torch.empty(B,co,HW,HW,device=dev)
L=lib(); L.mflow_debug_conv_trace.restype=ctypes.c_int; L.mflow_debug_conv_trace.argtypes=[ctypes.c_void_p]
for _ in range(3): check(L.mflow_conv2d_tf32x3(ctypes.byref(d),ptr(x),ptr(hi),ptr(lo),ptr(b),None,None,ptr(y),stream()),'conv')
torch.cuda.synchronize()
out=(ctypes.c_longlong*320)(); L.mflow_debug_conv_trace(out)
t=[list(out[i*64:(i+1)*64]) for i in range(5)]
t0=t[4][0]
print(f'B={B} HW={HW}: acc complete at {t[4][1]-t0}, kernel end at {t[4][2]-t0} cycles')
print('raw box ready per kb:', [v-t0 for v in t[3][:4]])
print('step: xf_start xf_done (dur) | mma_issue | gap to previous issue')
prev=None
for s in range(36):
    a,bb,c=t[0][s]-t0,t[1][s]-t0,t[2][s]-t0
    print(f'{s:3d}: {a:7d} {bb:7d} ({bb-a:5d}) | {c:7d} | {"" if prev is None else c-prev}')
    prev=c

# per-SM CTA timeline (globaltimer, ns)
import collections
L.mflow_debug_conv_ctas.restype=ctypes.c_int; L.mflow_debug_conv_ctas.argtypes=[ctypes.c_void_p]
buf=(ctypes.c_ulonglong*(4096*3))(); L.mflow_debug_conv_ctas(buf)
n_cta=min(4096, (B*HW*HW+127)//128)
recs=[(buf[3*i],buf[3*i+1],buf[3*i+2]) for i in range(n_cta)]
t_min=min(r[0] for r in recs); t_max=max(r[1] for r in recs)
dur=[r[1]-r[0] for r in recs]
print(f'CTAs {n_cta}: kernel span {(t_max-t_min)/1e3:.1f} us; CTA duration mean {sum(dur)/len(dur)/1e3:.2f} us, min {min(dur)/1e3:.2f}, max {max(dur)/1e3:.2f}')
by_sm=collections.defaultdict(list)
for s0,e0,sm in recs: by_sm[sm].append((s0-t_min,e0-t_min))
sm0=sorted(by_sm)[0]
print(f'SM {sm0}: {len(by_sm[sm0])} CTAs; (start,end) us:', [(round(a/1e3,1),round(b/1e3,1)) for a,b in sorted(by_sm[sm0])])
busy=[]
for sm,l in by_sm.items():
    l.sort(); busy.append(sum(b-a for a,b in l))
print(f'mean CTA-time per SM {sum(busy)/len(busy)/1e3:.1f} us over {len(by_sm)} SMs (2 CTAs run concurrently)')
