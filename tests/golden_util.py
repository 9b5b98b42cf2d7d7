"""Helpers shared by ``oracle/make_golden.py`` (reference side, build container only) and the
tests (oracle / product side).  Pure torch + numpy + hashlib; imports neither the reference
nor the product package."""
from __future__ import annotations

import hashlib
import json
import os
from argparse import Namespace

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 1357  # reference default --seed, experiments/train.py:93


def tensor_sha(t: torch.Tensor) -> str:
    a = t.detach().cpu().contiguous().numpy()
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()[:16]


def state_dict_hashes(sd) -> dict:
    return {k: tensor_sha(v) for k, v in sd.items()}


def perturb_state_dict(sd, seed: int = 2468) -> dict:
    """Deterministic, key-order-independent perturbation that moves a freshly initialised
    model away from its near-identity start (zero LU entries, 1e-3 residual layers) while
    keeping every linear layer well conditioned.  Applied identically to the reference's and
    to the product's ``state_dict`` so both sides evaluate the same non-trivial weights."""
    out = {}
    for key in sorted(sd.keys()):
        v = sd[key]
        if not torch.is_floating_point(v) or key.endswith("_scale") or key.endswith("_shift"):
            out[key] = v.clone()
            continue
        g = torch.Generator().manual_seed(seed + int(hashlib.sha256(key.encode()).hexdigest()[:8], 16) % 100003)
        noise = torch.randn(v.shape, generator=g, dtype=torch.float32).to(v.dtype)
        if key.endswith("lower_entries") or key.endswith("upper_entries"):
            n = int(round((1 + np.sqrt(1 + 8 * v.numel())) / 2))
            scale = 0.2 / np.sqrt(max(n, 1))
        elif key.endswith("unconstrained_upper_diag") or key.endswith("log_scale"):
            scale = 0.1
        elif key.endswith("shift") or key.endswith("bias"):
            scale = 0.05
        else:
            scale = 0.02
        out[key] = v + scale * noise
    return {k: out[k] for k in sd.keys()}


# BASELINE.json configs as the argparse namespace experiments/architectures/create_model.py reads
# (defaults: experiments/train.py:41-65; overrides: the cited experiments/configs/*.config)
_DEFAULT_ARGS = dict(
    algorithm="mf", specified=False, outertransform="rq-coupling", innertransform="rq-coupling", lineartransform="permutation",
    outerlayers=5, innerlayers=5, conditionalouter=False, dropout=0.0, pieepsilon=0.01, pieclip=None, encoderblocks=5,
    encoderhidden=100, splinerange=3.0, splinebins=8, levels=3, actnorm=False, batchnorm=False, linlayers=2,
    linchannelfactor=2, intermediatensf=False, decoderblocks=5, decoderhidden=100, datadim=3, modellatentdim=2,
)

CONFIGS = {
    "cfg1": dict(args=dict(datadim=3, modellatentdim=2), image=False, data_dim=3, parameter_dim=None, batch=8),
    "cfg2": dict(args=dict(datadim=3, modellatentdim=2, splinerange=6.0, splinebins=10), image=False, data_dim=3, parameter_dim=1,
                 batch=8),
    "cfg3": dict(args=dict(datadim=40, modellatentdim=14, outerlayers=20, innerlayers=15, lineartransform="lu", splinerange=10.0,
                           splinebins=11), image=False, data_dim=40, parameter_dim=2, batch=6),
    "cfg4": dict(args=dict(datadim=(3, 64, 64), modellatentdim=64, outerlayers=20, innerlayers=8, levels=4, linlayers=2,
                           linchannelfactor=2, lineartransform="lu", splinerange=10.0, splinebins=11, actnorm=True),
                 image=True, data_dim=(3, 64, 64), parameter_dim=1, batch=2),
    "cfg5": dict(args=dict(algorithm="flow", datadim=(3, 64, 64), modellatentdim=512, outerlayers=20, innerlayers=8, levels=4,
                           splinerange=10.0, splinebins=11, actnorm=True),
                 image=True, data_dim=(3, 64, 64), parameter_dim=None, batch=2),
    # reduced image variants (same code path, 16x16 inputs, 2 levels) that run in well under a second
    "cfg4s": dict(args=dict(datadim=(3, 16, 16), modellatentdim=16, outerlayers=4, innerlayers=2, levels=2, linlayers=2,
                            linchannelfactor=2, lineartransform="lu", splinerange=10.0, splinebins=11, actnorm=True),
                  image=True, data_dim=(3, 16, 16), parameter_dim=1, batch=3),
    "cfg5s": dict(args=dict(algorithm="flow", datadim=(3, 16, 16), modellatentdim=16, outerlayers=2, innerlayers=2, levels=2,
                            splinerange=10.0, splinebins=11, actnorm=True),
                  image=True, data_dim=(3, 16, 16), parameter_dim=None, batch=3),
}


def config_args(name: str) -> Namespace:
    d = dict(_DEFAULT_ARGS)
    d.update(CONFIGS[name]["args"])
    return Namespace(**d)


class FakeSimulator:
    """The three methods ``create_model`` calls on a simulator (create_model.py:15-62)."""

    def __init__(self, name: str):
        self._c = CONFIGS[name]

    def is_image(self):
        return self._c["image"]

    def data_dim(self):
        return self._c["data_dim"]

    def parameter_dim(self):
        return self._c["parameter_dim"]


def config_inputs(name: str, batch: int | None = None, seed: int = 4321, dtype=torch.float32):
    """Synthetic inputs of SURVEY.md section 8(d)."""
    c = CONFIGS[name]
    b = batch or c["batch"]
    g = torch.Generator().manual_seed(seed)
    if c["image"]:
        x = torch.rand((b,) + tuple(c["data_dim"]), generator=g, dtype=torch.float32) * 256.0
    else:
        x = torch.randn((b, c["data_dim"]), generator=g, dtype=torch.float32)
    ctx = None
    if c["parameter_dim"]:
        ctx = torch.randn((b, c["parameter_dim"]), generator=g, dtype=torch.float32)
    return x.to(dtype), (None if ctx is None else ctx.to(dtype))


def load_json(name: str):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def load_npz(name: str):
    return dict(np.load(os.path.join(GOLDEN_DIR, name)))
