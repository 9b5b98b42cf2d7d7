"""A few launches of one conditioner convolution (and its weight gradient) at batch 256 for ncu: 
python tools/conv_prof.py [fp16|tf32] [hw] [kernel]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch

from manifold_flow.b200 import ops
from manifold_flow.b200.config import config

fmt = sys.argv[1] if len(sys.argv) > 1 else "fp16"
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
k = int(sys.argv[3]) if len(sys.argv) > 3 else 3
config.conv_operand_format = fmt
dev = torch.device("cuda")
B, ci, co = 256, 100, 100
xs = [torch.randn(B, ci, hw, hw, device=dev) for _ in range(3)]
gs = [torch.randn(B, co, hw, hw, device=dev) for _ in range(3)]
w = torch.randn(co, ci, k, k, device=dev) * 0.05
bias = torch.randn(co, device=dev)
with torch.no_grad():
    split = ops._weight_cache.get(w, False, None, fmt)
    for i in range(3):
        ax = ops.tensor_amax(xs[i]) if fmt == "fp16" else None
        ag = ops.tensor_amax(gs[i]) if fmt == "fp16" else None
        ops._conv_tc_launch(xs[i], split, co, k, bias, None, None, True, ax)
        ops.conv_wgrad_tc(gs[i], xs[i], k, True, with_bias=True, x_amax=ax, g_amax=ag)
torch.cuda.synchronize()
print("ok")
