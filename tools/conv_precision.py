import torch, torch.nn.functional as F
dev='cuda'
torch.manual_seed(0)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32=tf32; torch.backends.cuda.matmul.allow_tf32=tf32
    for shape in ((2,100,32,32),(64,100,32,32),(2,100,4,4)):
        x=torch.randn(*shape,device=dev); w=torch.randn(100,100,3,3,device=dev)*0.05; b=torch.randn(100,device=dev)
        g=torch.randn(*shape,device=dev)
        def run(dt):
            xx=x.detach().clone().to(dt).requires_grad_(True); ww=w.detach().clone().to(dt).requires_grad_(True); bb=b.detach().clone().to(dt).requires_grad_(True)
            y=F.conv2d(F.relu(xx),ww,bb,padding=1); y.backward(g.to(dt))
            return y.detach(),xx.grad,ww.grad,bb.grad
        r32=run(torch.float32); r64=run(torch.float64)
        rel=lambda a,b_: float((a.double()-b_).norm()/b_.norm())
        print(f'tf32={tf32} shape={shape}: fwd {rel(r32[0],r64[0]):.2e} dgrad {rel(r32[1],r64[1]):.2e} wgrad {rel(r32[2],r64[2]):.2e} bgrad {rel(r32[3],r64[3]):.2e}', torch.backends.cudnn.allow_tf32)
try:
    print('fp32_precision conv:', torch.backends.cudnn.conv.fp32_precision, 'matmul:', torch.backends.cuda.matmul.fp32_precision)
except Exception as e: print('no new api', e)
