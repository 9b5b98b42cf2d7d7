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
    for bits, nb, nraw, stag in ((0, 0, 0, 0), (0, 5, 2, 0), (0, 6, 2, 0), (0, 8, 2, 0), (5, 8, 2, 0), (2 + 16, 8, 2, 0), (23, 8, 2, 0)):
        os.environ["MFLOW_CONV_ABLATE"] = str(bits)
        os.environ["MFLOW_CONV_STAGGER"] = str(stag)
        os.environ["MFLOW_CONV_NB"] = str(nb)
        os.environ["MFLOW_CONV_NRAW"] = str(nraw)
        for i in range(3):
            ops._conv_tc_launch(xs[i], split, 100, 3, bias, None, None, True, amax[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            ops._conv_tc_launch(xs[i % 6], split, 100, 3, bias, None, None, True, amax[i % 6])
        e1.record()
        torch.cuda.synchronize()
        print(f"ablate={bits:2d} n_b={nb} n_raw={nraw} stagger={stag} ns: {e0.elapsed_time(e1) / 20 * 1e3:.0f} us   (1 no B_lo product, 2 no gather loads, 4 MMA N=64, 16 no TMEM stores)")

# per-CTA timeline of one launch (ablate build)
import ctypes
import numpy as np
with torch.no_grad(), cabi.library_variant("ablate") as lib:
    for bits, stag in ((0, 0),):
        os.environ["MFLOW_CONV_ABLATE"] = str(bits)
        os.environ["MFLOW_CONV_STAGGER"] = str(stag)
        os.environ["MFLOW_CONV_NB"] = "8"
        os.environ["MFLOW_CONV_NRAW"] = "2"
        ops._conv_tc_launch(xs[0], split, 100, 3, bias, None, None, True, amax[0])
        torch.cuda.synchronize()
        n = B * hw * hw // 128
        buf = np.zeros((min(n, 4096), 9), dtype=np.int64)
        fn = lib.mflow_debug_conv_timeline
        fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]
        fn(buf.ctypes.data, buf.shape[0])
        d = buf[:, 1:6] - buf[:, 0:5]
        print(f"ablate={bits} stagger={stag}: per-CTA cycles (median over {buf.shape[0]} CTAs): setup {np.median(d[:,0]):.0f}, to first MMA {np.median(d[:,1]):.0f}, "
              f"main loop {np.median(d[:,2]):.0f}, epilogue {np.median(d[:,3]):.0f}, teardown {np.median(d[:,4]):.0f}, total {np.median(buf[:,5]-buf[:,0]):.0f}; "
              f"wall per CTA {np.median(buf[:,7]-buf[:,6]):.0f} ns")
        # gaps between consecutive CTAs of one SM slot
        for sm in (0, 1):
            rows = buf[buf[:, 8] == sm]
            rows = rows[np.argsort(rows[:, 6])]
            print("  SM", sm, "CTA start/end (us since first):", [(round((r[6]-rows[0,6])/1e3,1), round((r[7]-rows[0,6])/1e3,1)) for r in rows[:10]])
        st = np.zeros((7, 64), dtype=np.int64)
        fn2 = lib.mflow_debug_conv_steps
        fn2.restype, fn2.argtypes = ctypes.c_int, [ctypes.c_void_p]
        fn2(st.ctypes.data)
        t0 = st[0, 0]
        print("  step: xf_wait_begin xf_stage_free xf_done | mma_a_full mma_b_full mma_issued | tma_slot_free   (cycles since the first)")
        for i in range(36):
            print("  %2d: %6d %6d %6d | %6d %6d %6d | %6d" % ((i,) + tuple(int(st[r, i] - t0) for r in range(7))))
