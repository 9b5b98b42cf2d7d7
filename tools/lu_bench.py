"""Which library kernels the large LULinear (3840 features, batch 256) launches, per operation."""
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda'); n, B = 3840, 256
torch.manual_seed(0)
L = (torch.eye(n, device=dev) + 0.01 * torch.randn(n, n, device=dev).tril(-1)).requires_grad_()
U = (torch.eye(n, device=dev) + 0.01 * torch.randn(n, n, device=dev).triu(1)).requires_grad_()
x = torch.randn(B, n, device=dev, requires_grad=True)
def fwd(): return F.linear(F.linear(x, U), L)
def inv():
    o = torch.linalg.solve_triangular(L, x.t(), upper=False, unitriangular=True)
    return torch.linalg.solve_triangular(U, o, upper=True).t()
for name, fn in (('forward (2 x F.linear)', fwd), ('inverse (2 x solve_triangular)', inv)):
    for with_bwd in (False, True):
        def run():
            y = fn()
            if with_bwd: y.backward(torch.ones_like(y)); L.grad = U.grad = x.grad = None
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        print(f'=== {name}{" + backward" if with_bwd else ""}: {e0.elapsed_time(e1)/5:.3f} ms')
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run(); torch.cuda.synchronize()
        for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
            print(f'   {ev.device_time_total/1e3:8.3f} ms x{ev.count:3d}  {ev.key[:110]}')

# blocked substitution (manifold_flow.b200.ops.tri_solve) against the library solve
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R,'manifold-flow_b200'))
from manifold_flow.b200 import ops
xt = x.detach().t().contiguous()
for blk in (128, 256, 512):
    def inv_blocked():
        o = ops.tri_solve(L, xt, upper=False, unit=True, block=blk)
        return ops.tri_solve(U, o, upper=True, block=blk)
    ref = torch.linalg.solve_triangular(U.detach().double(), torch.linalg.solve_triangular(L.detach().double(), xt.double(), upper=False, unitriangular=True), upper=True)
    got = inv_blocked()
    lib = torch.linalg.solve_triangular(U, torch.linalg.solve_triangular(L, xt, upper=False, unitriangular=True), upper=True)
    print(f'block {blk}: rel err vs fp64 {((got.double()-ref).norm()/ref.norm()).item():.2e} (library {((lib.double()-ref).norm()/ref.norm()).item():.2e})')
    for with_bwd in (False, True):
        def run():
            y = inv_blocked()
            if with_bwd: y.backward(torch.ones_like(y)); L.grad = U.grad = None
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        print(f'   blocked inverse{" + backward" if with_bwd else ""}: {e0.elapsed_time(e1)/5:.3f} ms (eager, launch-bound on the host; the bench replays a graph)')
