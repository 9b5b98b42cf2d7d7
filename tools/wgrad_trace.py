"""Dev-only: per-block timeline of one CTA of the fp16 weight-gradient kernel (libmflow_b200_ablate.so)."""
import ctypes
import os
import sys

import numpy as np

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
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x = torch.randn(B, 100, hw, hw, device=dev)
g = torch.randn(B, 100, hw, hw, device=dev)
with torch.no_grad(), cabi.library_variant("ablate") as lib:
    ax, ag = ops.tensor_amax(x), ops.tensor_amax(g)
    for _ in range(3):
        ops.conv_wgrad_tc(g, x, k, True, with_bias=True, x_amax=ax, g_amax=ag)
    torch.cuda.synchronize()
    st = np.zeros((8, 64), dtype=np.int64)
    fn = lib.mflow_debug_wgrad_steps
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p]
    fn(st.ctypes.data)
    t0 = st[0, 0]
    print("block: tma_raw_slot_free tma_q_slot_free | qsplit_arrived qsplit_done | p_raw_arrived p_tap0_done | mma_q_ready mma_issued   (cycles)")
    for i in range(40):
        print("  %2d: %6d %6d | %6d %6d | %6d %6d | %6d %6d" % (i, st[0, i] - t0, st[1, i] - t0, st[2, i] - t0, st[3, i] - t0, st[4, i] - t0, st[7, i] - t0,
                                                              st[5, i] - t0, st[6, i] - t0))
