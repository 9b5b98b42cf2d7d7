#!/usr/bin/env python
"""Golden spline cases (tests/golden/spline.npz, frozen from the live reference) through the default build and the
exact-operation-order build (libmflow_b200_exact.so): element-wise error against the reference's fp32 output on the
elements where the reference's own fp32 result is well conditioned (within 2e-6 relative of its fp64 result)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("manifold-flow_b200", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import golden_util as gu  # noqa: E402
from manifold_flow.b200 import cabi  # noqa: E402
from manifold_flow.transforms.splines import unconstrained_rational_quadratic_spline as spline  # noqa: E402

S = gu.load_npz("spline.npz")
dev = torch.device("cuda")
t_ = lambda a: torch.from_numpy(np.asarray(a))
print(f"{'case':28s} {'dir':3s} {'build':7s} {'key':3s} {'well':>5s} {'>1e-5':>5s} {'max_rel(well)':>13s} {'max_rel(all)':>12s} {'ref32 max_rel(all)':>18s} {'bit-equal':>9s}")
for name in [str(c) for c in S["cases"]]:
    t = float(name.split("_")[1][1:])
    x, w, h, d = (t_(S[f"{name}/{k}"]).to(dev) for k in "xwhd")
    for inverse in (False, True):
        tag = "inv" if inverse else "fwd"
        res = {"default": spline(x, w, h, d, inverse=inverse, tails="linear", tail_bound=t)}
        with cabi.library_variant("exact"):
            res["exact"] = spline(x, w, h, d, inverse=inverse, tails="linear", tail_bound=t)
        for build, (y, lad) in res.items():
            for key, got in (("y", y), ("lad", lad)):
                ref32, ref64 = t_(S[f"{name}/{tag}_{key}"]).double(), t_(S[f"{name}/{tag}_{key}64"])
                scale = ref64.abs().clamp_min(1.0)
                well = (ref32 - ref64).abs() <= 2e-6 * scale
                err = (got.cpu().double() - ref32).abs() / scale
                eref = (ref32 - ref64).abs() / scale
                same = float((got.cpu() == t_(S[f"{name}/{tag}_{key}"])).float().mean())
                print(f"{name:28s} {tag:3s} {build:7s} {key:3s} {int(well.sum()):5d} {int((err[well] > 1e-5).sum()):5d} {float(err[well].max()):13.2e} "
                      f"{float(err.max()):12.2e} {float(eref.max()):18.2e} {same:9.2f}")
