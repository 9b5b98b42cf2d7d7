"""A few launches of the rows x planes spline kernels at the cfg3 bench shape (65536 rows, 40 features, 20 transformed,
K = 11, parameters [64, 640, 32, 32]) for ncu, plus CUDA-event timings of the same launches against the rows layout."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch

from manifold_flow.b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
rows, d, ct, k, w = 65536, 40, 20, 11, 32
p = 3 * k - 1
geom = ops.SplineGeometry(ct, 0, 2, d - ct, 1, 2)
x = torch.randn(rows, d, device=dev, generator=g) * 4
par_rows = torch.randn(rows, ct * p, device=dev, generator=g) * 3
par_planes = par_rows.view(rows // (w * w), w * w, ct * p).transpose(1, 2).contiguous().view(-1, ct * p, w, w)
args = (geom, k, 10.0, 1e-3, 1e-3, 1e-3, 10.0)


def run(par, inverse, backward):
    xx, pp = x.clone().requires_grad_(backward), par.clone().requires_grad_(backward)
    y, lad = ops.rqs_coupling(xx, pp, *args, inverse)
    if backward:
        (y.sum() + lad.sum()).backward()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    for _ in range(2):
        run(par_planes, False, True)
        with torch.no_grad():
            run(par_planes, True, False)
    torch.cuda.synchronize()
else:
    n = rows * ct
    for name, par in (("rows", par_rows), ("rows x planes", par_planes)):
        with torch.no_grad():
            tf, ti = timed(lambda: run(par, False, False)), timed(lambda: run(par, True, False))
        print(f"{name:14s} fwd {tf:7.1f} us = {n * 4 * (p + 2) / tf / 1e3:6.0f} GB/s   inv {ti:7.1f} us = {n * 4 * (p + 2) / ti / 1e3:6.0f} GB/s"
              "   (times include two clone() launches of the inputs)")
print("ok")
