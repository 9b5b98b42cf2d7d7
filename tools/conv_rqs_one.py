"""A few launches of the final 1x1 convolution with the spline in its epilogue (level 0 of the 64x64x3 models at batch
256: hidden [256, 100, 32, 32] -> 6 transformed channels x 32 parameters) for ncu:
ncu --set full --clock-control none --import-source on -k regex:conv_split3 -s 2 -c 1 python tools/conv_rqs_one.py"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch

from manifold_flow.b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
B = 256
hidden = torch.randn(B, 100, 32, 32, device=dev)
w = torch.randn(192, 100, 1, 1, device=dev) * 0.3
b = torch.randn(192, device=dev)
x = torch.randn(B, 12, 32, 32, device=dev) * 4
geom = ops.SplineGeometry(6, 0, 1, 6, 6, 1)
with torch.no_grad():
    for _ in range(4):
        y, lad = ops.conv_rqs(hidden, w, b, x, geom, 11, 10.0, 1e-3, 1e-3, 1e-3, 10.0, False)
    torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(lad.mean()))
