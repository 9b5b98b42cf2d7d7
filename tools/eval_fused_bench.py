"""Evaluation throughput of the image models (no autograd) with the spline in the final convolution's epilogue
(config.fuse_spline_epilogue) against the two-launch path.  usage: python tools/eval_fused_bench.py [cfg4|cfg5] [batch]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch

import bench
import golden_util as gu
from manifold_flow.b200 import config, factory

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 256
C = bench.CONFIGS[cfg]
dev = torch.device("cuda")
torch.manual_seed(gu.SEED)
model = factory.create_model(cfg).to(dev)
model.train()
x, c = bench.synthetic_batch(cfg, b, 0)
x, c = x.to(dev), (None if c is None else c.to(dev))
with torch.no_grad():
    (model.log_prob(x, context=c) if C["kind"] == "train_flow" else model(x, mode="projection", context=c))
model.eval()


def run():
    with torch.no_grad():
        return model.log_prob(x, context=c) if C["kind"] == "train_flow" else model(x, mode=C["mode"], context=c)[1]


out = {}
for fused in (False, True):
    config.fuse_spline_epilogue = fused
    for _ in range(3):
        lp = run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        lp = run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[fused] = (ms, lp.clone())
    print(f"{cfg} batch {b} evaluation log_prob, spline {'in the conv epilogue' if fused else 'as its own launch'}: {ms:.3f} ms / pass = {b / ms * 1e3:.0f} samples/s")
d = (out[True][1] - out[False][1]).abs().max().item()
print(f"max |log_prob difference| between the two paths: {d:.3g} (|log_prob| up to {out[False][1].abs().max().item():.3g})")
