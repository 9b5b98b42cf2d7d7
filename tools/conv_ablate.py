"""Dev-only: time the level-0 3x3 fp16 convolution with parts of the kernel switched off (libmflow_b200_ablate.so,
built by `make -C manifold-flow_b200/csrc ablate`).  Results are wrong by design; only the times mean something."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch

from manifold_flow.b200 import cabi, ops
from manifold_flow.b200.config import config

config.conv_operand_format = "fp16"
dev = torch.device("cuda")
B = 256
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
xs = [torch.randn(B, 100, hw, hw, device=dev) for _ in range(6)]
w = torch.randn(100, 100, 3, 3, device=dev) * 0.05
bias = torch.randn(100, device=dev)
with torch.no_grad(), cabi.library_variant("ablate"):
    split = ops._weight_cache.get(w, False, None, "fp16")
    amax = [ops.tensor_amax(t) for t in xs]
    for bits in (0, 1, 2, 4, 16, 2 | 16, 1 | 4, 1 | 2 | 4 | 16):
        os.environ["MFLOW_CONV_ABLATE"] = str(bits)
        for i in range(3):
            ops._conv_tc_launch(xs[i], split, 100, 3, bias, None, None, True, amax[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            ops._conv_tc_launch(xs[i % 6], split, 100, 3, bias, None, None, True, amax[i % 6])
        e1.record()
        torch.cuda.synchronize()
        print(f"ablate={bits:2d}: {e0.elapsed_time(e1) / 20 * 1e3:.0f} us   (1 no B_lo product, 2 no gather loads, 4 MMA N=64, 16 no TMEM stores)")
