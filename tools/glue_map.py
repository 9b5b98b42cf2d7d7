"""Where do the non-libmflow (ATen / library) kernel launches of one step come from?  torch.profiler with Python stacks:
every CPU-side op that launched device kernels is attributed to the innermost frame inside this repo (forward and custom
backward functions) or to the autograd node that ran it (engine-side ops).
usage: python tools/glue_map.py [cfg4] [256] > gpurun_out/glue_map.txt"""
import collections
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "oracle", "tests", ""):
    sys.path.insert(0, os.path.join(R, p))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import golden_util as gu  # noqa: E402
from manifold_flow.b200 import factory  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
C = bench.CONFIGS[cfg]
b = int(sys.argv[2]) if len(sys.argv) > 2 else C["batch"]
dev = torch.device("cuda")
torch.manual_seed(gu.SEED)
model = factory.create_model(cfg).to(dev)
model.train()
x, c = bench.synthetic_batch(cfg, b, 0)
x, c = x.to(dev), (None if c is None else c.to(dev))


def forward():
    if C["kind"] == "train_flow":
        return model.log_prob(x, context=c)
    return model(x, mode=C["mode"], context=c)[1]


with torch.no_grad():
    if C["kind"] == "train_flow":
        forward()
    else:
        model(x, mode="projection", context=c)


def step():
    model.zero_grad(set_to_none=True)
    (-forward().mean()).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()

agg = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
for e in prof.events():
    if e.device_type.name != "CPU" or not e.kernels:
        continue
    # an op's own kernels only (children report theirs)
    child_kernels = sum(len(ch.kernels) for ch in e.cpu_children)
    own = len(e.kernels) - child_kernels
    if own <= 0:
        continue
    own_us = sum(k.duration for k in e.kernels) - sum(sum(k.duration for k in ch.kernels) for ch in e.cpu_children)
    where = None
    for fr in (e.stack or []):
        if "manifold_flow" in fr or "bench.py" in fr or "glue_map" in fr:
            where = fr.split("manifold-flow_b200/")[-1].strip()
            break
    if where is None:
        p = e.cpu_parent
        while p is not None:
            if p.name.startswith("autograd::engine::evaluate_function") or "Backward" in p.name:
                where = "engine: " + p.name.replace("autograd::engine::evaluate_function: ", "")
                break
            p = p.cpu_parent
    names = collections.Counter(k.name.split("<")[0].split("(")[0][-60:] for k in e.kernels)
    rec = agg[(e.name, where or "?")]
    rec[0] += own
    rec[1] += own_us
    rec[2].update(names)

rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
tot_n = sum(v[0] for _, v in rows)
tot_us = sum(v[1] for _, v in rows)
print(f"# {cfg} batch {b}: {tot_n} kernel launches, {tot_us / 1e3:.2f} ms device time in one eager step")
ours = sum(v[0] for (name, _), v in rows if any("mflow" in k for k in v[2]))
print(f"# launched from ops whose kernels are libmflow_b200's: {ours}")
print("| launches | us | op | from | kernels |\n|---:|---:|---|---|---|")
for (name, where), (n, us, ks) in rows[:120]:
    kn = ", ".join(f"{k}x{v}" for k, v in ks.most_common(2))
    print(f"| {n} | {us:.0f} | {name} | {where} | {kn[:90]} |")
