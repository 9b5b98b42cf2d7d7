"""A few launches of the fused spline kernels (level-0 shape of the 64x64x3 models at batch 256: [256, 12, 32, 32], 6
transformed channels x 32 parameters) and of the fused vector conditioner (256 rows, 33 -> 100 -> 1024) for ncu."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch

from manifold_flow import nn as nn_
from manifold_flow.b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
geom = ops.SplineGeometry(6, 0, 1, 6, 6, 1)
for i in range(3):
    x = (torch.randn(256, 12, 32, 32, device=dev, generator=g) * 5).requires_grad_(True)
    p = (torch.randn(256, 6 * 32, 32, 32, device=dev, generator=g) * 3).requires_grad_(True)
    y, lad = ops.rqs_coupling(x, p, geom, 11, 10.0, 1e-3, 1e-3, 1e-3, 10.0, False)
    (y.sum() + lad.sum()).backward()
    with torch.no_grad():
        ops.rqs_coupling(y.detach(), p.detach(), geom, 11, 10.0, 1e-3, 1e-3, 1e-3, 10.0, True)
net = nn_.ResidualNet(32, 1024, hidden_features=100, context_features=1, num_blocks=2).to(dev)
for i in range(3):
    xv = torch.randn(256, 64, device=dev, requires_grad=True)
    out = net(xv[:, 1::2], torch.randn(256, 1, device=dev))
    out.sum().backward()
torch.cuda.synchronize()
print("ok")
