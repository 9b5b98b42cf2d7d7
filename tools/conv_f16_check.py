"""fp16-split vs TF32-split tcgen05 convolution / weight-gradient kernels: accuracy against fp64 (including operands far
outside fp16's exponent range: gradients of 1e-7, activations of 1e4) and time per launch at the bench shapes."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch
import torch.nn.functional as F

from manifold_flow.b200 import ops
from manifold_flow.b200.config import config

dev = torch.device("cuda")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def rel(a, b):
    return float((a.double() - b).norm() / (b.norm() + 1e-300))


cases = [(2, 100, 100, 32, 3), (3, 100, 100, 16, 3), (5, 100, 100, 8, 3), (9, 100, 100, 4, 3), (2, 6, 100, 32, 1), (2, 100, 192, 32, 1),
         (3, 100, 1536, 4, 1), (2, 48, 100, 4, 1), (2, 16, 16, 8, 3), (1, 1, 1, 8, 3), (2, 130, 40, 16, 1)]
if quick:
    cases = cases[:2]
print("== accuracy vs fp64 (rel L2): y, grad_x, grad_w, grad_b; x_scale / g_scale multiply the operands")
for fmt in ("fp16", "tf32"):
    config.conv_operand_format = fmt
    for (b, ci, co, hw, k) in cases:
        for xs, gs in ((1.0, 1.0), (1e4, 1e-7), (1e-6, 1e3)):
            g = torch.Generator(device="cuda").manual_seed(b + ci + co + hw)
            x = torch.randn(b, ci, hw, hw, generator=g, device=dev) * xs
            w = torch.randn(co, ci, k, k, generator=g, device=dev) * (1.0 / (ci * k * k) ** 0.5)
            bias = torch.randn(co, generator=g, device=dev) * xs
            gy = torch.randn(b, co, hw, hw, generator=g, device=dev) * gs

            def run(dtype, fn):
                leaves = [t.detach().clone().to(dtype).requires_grad_(True) for t in (x, w, bias)]
                y = fn(*leaves)
                y.backward(gy.to(dtype))
                return [y.detach()] + [t.grad for t in leaves]

            ref = run(torch.float64, lambda xx, ww, bb: F.conv2d(F.relu(xx), ww, bb, padding=k // 2))
            try:
                got = run(torch.float32, lambda xx, ww, bb: ops.conv_tc(xx, ww, bb, relu_in=True))
                torch.cuda.synchronize()
            except Exception as exc:
                print(fmt, (b, ci, co, hw, k), "FAILED", repr(exc)[:300])
                sys.exit(1)
            errs = [rel(a, r) for a, r in zip(got, ref)]
            nan = sum(int(torch.isnan(a).sum()) for a in got)
            print(f"{fmt} B={b} {ci}->{co} {hw}x{hw} k={k} xs={xs:g} gs={gs:g}: " + " ".join(f"{e:.2e}" for e in errs) + f" nan={nan}")
            if xs != 1.0 and quick:
                break

print("== time per launch (CUDA events, 20 launches over 6 rotating inputs), batch 256")


def timeit(fn, n=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = 256
shapes = [(100, 100, 32, 3), (100, 100, 16, 3), (100, 100, 8, 3), (100, 100, 4, 3), (6, 100, 32, 1), (100, 192, 32, 1), (100, 1536, 4, 1)]
if quick:
    shapes = shapes[:1]
for (ci, co, hw, k) in shapes:
    xs_ = [torch.randn(B, ci, hw, hw, device=dev) for _ in range(6)]
    gs_ = [torch.randn(B, co, hw, hw, device=dev) for _ in range(6)]
    w = torch.randn(co, ci, k, k, device=dev) * 0.05
    bias = torch.randn(co, device=dev)
    line = f"{ci}->{co} {hw}x{hw} k={k}:"
    for fmt in ("fp16", "tf32"):
        config.conv_operand_format = fmt
        with torch.no_grad():
            amax = [ops.tensor_amax(t) for t in xs_] if fmt == "fp16" else [None] * 6
            gam = [ops.tensor_amax(t) for t in gs_] if fmt == "fp16" else [None] * 6
            split = ops._weight_cache.get(w, False, None, fmt)
            t_f = timeit(lambda i: ops._conv_tc_launch(xs_[i % 6], split, co, k, bias, None, None, True, amax[i % 6]))
            t_w = timeit(lambda i: ops.conv_wgrad_tc(gs_[i % 6], xs_[i % 6], k, True, with_bias=True, x_amax=amax[i % 6], g_amax=gam[i % 6])) \
                if ops.conv_wgrad_supported(xs_[0], co, k) else float("nan")
            t_a = timeit(lambda i: ops.tensor_amax(xs_[i % 6], fresh=True))
        fl = 2 * B * hw * hw * ci * co * k * k
        line += f"  [{fmt}] conv {t_f:.0f} us ({fl / t_f / 1e6:.0f} TF/s) wgrad {t_w:.0f} us ({fl / t_w / 1e6:.0f} TF/s) absmax(x) {t_a:.0f} us"
    print(line)
