"""Back-to-back launches of the 1x1 weight-gradient kernel (preceded by a convolution, as in the ResidualNet backward) on
large inputs; every result must be bit-identical to the first.  With `raw3` the dev variant libmflow_b200_raw3.so (built
with -DMFLOW_WGRAD_FORCE_RAW3: the three-slot raw ring used before the fix; `make -C manifold-flow_b200/csrc raw3`) runs
instead, to reproduce the hazard.
usage: python tools/wgrad_repeat.py [raw3] [reps] [batch]"""
import contextlib
import os
import sys
import threading
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch  # noqa: E402

from manifold_flow.b200 import cabi, ops  # noqa: E402

variant = "raw3" if "raw3" in sys.argv else None
nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
reps = nums[0] if nums else 2000
batch = nums[1] if len(nums) > 1 else 256
progress = [0, time.time()]


def watchdog():
    while True:
        time.sleep(2)
        if time.time() - progress[1] > 40:
            print(f"HANG ({variant or 'fixed'}): no progress for 40 s after {progress[0]} synchronised chunks", flush=True)
            os._exit(7)


threading.Thread(target=watchdog, daemon=True).start()
dev = torch.device("cuda")
torch.manual_seed(3)
x = torch.randn(batch, 100, 32, 32, device=dev)
g = torch.randn(batch, 100, 32, 32, device=dev)
w = torch.randn(100, 100, 1, 1, device=dev) * 0.1
ctx = cabi.library_variant(variant) if variant else contextlib.nullcontext()
with ctx:
    first = ops.conv_wgrad_tc(g, x, 1, True)
    bad = torch.zeros((), dtype=torch.int32, device=dev)
    try:
        for chunk in range(reps // 100):
            for _ in range(100):
                y = ops.conv_tc(g, w)                       # the data gradient that precedes the weight gradient
                bad += (ops.conv_wgrad_tc(g, x, 1, True) != first).any()
            torch.cuda.synchronize()
            progress[0], progress[1] = chunk + 1, time.time()
        print(f"{variant or 'fixed'}: {reps} launches at batch {batch}, mismatching results: {int(bad)}")
    except Exception as exc:  # noqa: BLE001
        print(f"{variant or 'fixed'}: {type(exc).__name__} after {progress[0]} chunks: {str(exc).splitlines()[0][:120]}")
        os._exit(8)
