"""CPU oracle for the manifold-flow neural-spline coupling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``manifold-flow_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the timed CPU arm.

What it is
----------
A *functional* restatement of the reference's algorithm for the path
``BASELINE.json:north_star`` names: it walks a small architecture description (a
"spec" tree of tuples) and a flat ``state_dict`` that uses the reference's own key
names, and evaluates every layer with plain ``torch`` ops on whatever device / dtype
the inputs live on (CPU fp32 and fp64 in the tests).  There are no ``nn.Module``
classes, no boolean-mask gathers and no host syncs; every function cites the reference
lines whose arithmetic (and arithmetic *order*) it follows, because in fp32 the order
is part of the contract.

Third-party arithmetic: the reference executes every FLOP through PyTorch ATen
(``environment.yml`` pins ``pytorch>=1.4``; no lock file).  The oracle calls the same
ATen ops (softmax, cumsum, softplus, conv2d, linear, triangular solve) of the torch
installed here (2.11.0), so on CPU it reproduces the reference to the last bit wherever
the restatement is faithful.

Parity pin
----------
The reference ships no tests or golden vectors.  The oracle is
therefore pinned against the *live reference* imported from ``/root/reference`` in the
build container: ``oracle/make_golden.py`` runs the reference, freezes inputs/outputs
(and per-tensor SHA-256 of the seeded ``state_dict``) under ``tests/golden/`` and
``tests/test_oracle_golden.py`` replays them through this file on every CPU test run.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

MIN_BIN_WIDTH = 1e-3  # splines/rational_quadratic.py:11
MIN_BIN_HEIGHT = 1e-3  # :12
MIN_DERIVATIVE = 1e-3  # :13


# --------------------------------------------------------------------------------------
# (a) rational-quadratic spline            splines/rational_quadratic.py:16-168
# --------------------------------------------------------------------------------------
def rqs_knots(unnormalized: Tensor, lo: float, hi: float, min_size: float) -> Tuple[Tensor, Tensor]:
    """Knot positions [..., K+1] and bin sizes [..., K] for one axis of the spline.

    rational_quadratic.py:102-109 (widths) and :113-120 (heights): softmax, affine floor,
    cumsum, left zero pad, rescale as mul-then-add, both end knots overwritten exactly,
    and the bin sizes re-derived as differences of the rounded knots.
    """
    k = unnormalized.shape[-1]
    frac = F.softmax(unnormalized, dim=-1)
    frac = min_size + (1 - min_size * k) * frac
    knots = F.pad(torch.cumsum(frac, dim=-1), (1, 0), value=0.0)
    knots = (hi - lo) * knots + lo
    knots[..., 0] = lo
    knots[..., -1] = hi
    return knots, knots[..., 1:] - knots[..., :-1]


def search_bins(knots: Tensor, values: Tensor, eps: float = 1e-6) -> Tensor:
    """utils/various.py:472-474 - the last knot is bumped by ``eps`` (in the tensor's own
    precision), then bin = #(value >= knot) - 1."""
    bumped = knots.clone()
    bumped[..., -1] += eps
    return torch.sum(values[..., None] >= bumped, dim=-1) - 1


def rqs_padded_derivatives(unnormalized_derivatives: Tensor, min_derivative: float) -> Tensor:
    """rational_quadratic.py:38-41 - pad K-1 interior logits to K+1 with the constant whose
    softplus is 1 - min_derivative (float64 on the host, rounded when stored)."""
    padded = F.pad(unnormalized_derivatives, (1, 1))
    edge = np.log(np.exp(1 - min_derivative) - 1)
    padded[..., 0] = edge
    padded[..., -1] = edge
    return padded


def rqs_from_knots(
    x: Tensor, cw: Tensor, w: Tensor, ch: Tensor, h: Tensor, d: Tensor, inverse: bool
) -> Tuple[Tensor, Tensor, Tensor]:
    """Evaluate the monotone RQ spline given finished knots.  Returns (y, logabsdet, bin).

    rational_quadratic.py:122-168.  ``x`` must already lie inside [cw[0], cw[-1]].
    """
    idx = search_bins(ch if inverse else cw, x)  # :122-125
    g = idx.clamp(0, w.shape[-1] - 1)[..., None]
    xk = cw.gather(-1, g)[..., 0]
    wk = w.gather(-1, g)[..., 0]
    yk = ch.gather(-1, g)[..., 0]
    sk = (h / w).gather(-1, g)[..., 0]  # :131-132, delta computed for every bin first
    dk = d.gather(-1, g)[..., 0]
    dk1 = d[..., 1:].gather(-1, g)[..., 0]
    hk = h.gather(-1, g)[..., 0]

    if inverse:  # :139-156
        a = (x - yk) * (dk + dk1 - 2 * sk) + hk * (sk - dk)
        b = hk * dk - (x - yk) * (dk + dk1 - 2 * sk)
        c = -sk * (x - yk)
        disc = torch.clamp(b.pow(2) - 4 * a * c, min=0.0)
        root = (2 * c) / (-b - torch.sqrt(disc))
        out = root * wk + xk
        tt = root * (1 - root)
        den = sk + ((dk + dk1 - 2 * sk) * tt)
        dnum = sk.pow(2) * (dk1 * root.pow(2) + 2 * sk * tt + dk * (1 - root).pow(2))
        return out, -(torch.log(dnum) - 2 * torch.log(den)), idx
    # :158-168
    theta = (x - xk) / wk
    tt = theta * (1 - theta)
    num = hk * (sk * theta.pow(2) + dk * tt)
    den = sk + ((dk + dk1 - 2 * sk) * tt)
    out = yk + num / den
    dnum = sk.pow(2) * (dk1 * theta.pow(2) + 2 * sk * tt + dk * (1 - theta).pow(2))
    return out, torch.log(dnum) - 2 * torch.log(den), idx


def rqs_box(
    x: Tensor,
    unnormalized_widths: Tensor,
    unnormalized_heights: Tensor,
    unnormalized_derivatives: Tensor,
    inverse: bool = False,
    left: float = 0.0,
    right: float = 1.0,
    bottom: float = 0.0,
    top: float = 1.0,
    min_bin_width: float = MIN_BIN_WIDTH,
    min_bin_height: float = MIN_BIN_HEIGHT,
    min_derivative: float = MIN_DERIVATIVE,
):
    """``rational_quadratic_spline`` on the box [left, right] x [bottom, top]; rational_quadratic.py:67-168.
    All K+1 knot derivatives are parameters (:111).  The caller guarantees the domain (:85-86)."""
    d = min_derivative + F.softplus(unnormalized_derivatives)  # :111
    cw, w = rqs_knots(unnormalized_widths, left, right, min_bin_width)
    ch, h = rqs_knots(unnormalized_heights, bottom, top, min_bin_height)
    y, lad, _ = rqs_from_knots(x, cw, w, ch, h, d, inverse)
    return y, lad


def rqs(
    x: Tensor,
    unnormalized_widths: Tensor,
    unnormalized_heights: Tensor,
    unnormalized_derivatives: Tensor,
    inverse: bool = False,
    tail_bound: float = 1.0,
    min_bin_width: float = MIN_BIN_WIDTH,
    min_bin_height: float = MIN_BIN_HEIGHT,
    min_derivative: float = MIN_DERIVATIVE,
    return_bins: bool = False,
):
    """``unconstrained_rational_quadratic_spline(..., tails="linear")`` per element.

    rational_quadratic.py:16-64.  The reference evaluates the spline only on the masked
    subset ``-T <= x <= T`` and copies the rest through with logabsdet 0 (:31,:43-44).  Every
    op of the inner function is element-/row-local, so evaluating all rows on a clamped copy
    of ``x`` and selecting afterwards gives the same numbers without host syncs.
    """
    t = float(tail_bound)
    inside = (x >= -t) & (x <= t)
    x_in = torch.where(inside, x, torch.zeros_like(x))
    d = min_derivative + F.softplus(rqs_padded_derivatives(unnormalized_derivatives, min_derivative))  # :111
    cw, w = rqs_knots(unnormalized_widths, -t, t, min_bin_width)
    ch, h = rqs_knots(unnormalized_heights, -t, t, min_bin_height)
    y, lad, idx = rqs_from_knots(x_in, cw, w, ch, h, d, inverse)
    y = torch.where(inside, y, x)
    lad = torch.where(inside, lad, torch.zeros_like(lad))
    if return_bins:
        return y, lad, torch.where(inside, idx, torch.full_like(idx, -1))
    return y, lad


def sum_except_batch(x: Tensor) -> Tensor:
    """utils/various.py:329-334."""
    return torch.sum(x, dim=list(range(1, x.dim()))) if x.dim() > 1 else x


# --------------------------------------------------------------------------------------
# (b) conditioners                          nn/resnet.py
# --------------------------------------------------------------------------------------
def residual_net(sd: StateDict, p: str, x: Tensor, context: Optional[Tensor], num_blocks: int = 2) -> Tensor:
    """ResidualNet.forward, nn/resnet.py:64-72 with ResidualBlock.forward :27-40
    (batch norm off, dropout 0: every BASELINE config)."""
    if context is not None:
        x = torch.cat((x, context), dim=1)
    t = F.linear(x, sd[p + "initial_layer.weight"], sd[p + "initial_layer.bias"])
    for i in range(num_blocks):
        q = f"{p}blocks.{i}."
        r = F.linear(F.relu(t), sd[q + "linear_layers.0.weight"], sd[q + "linear_layers.0.bias"])
        r = F.linear(F.relu(r), sd[q + "linear_layers.1.weight"], sd[q + "linear_layers.1.bias"])
        if context is not None:
            gate = F.linear(context, sd[q + "context_layer.weight"], sd[q + "context_layer.bias"])
            r = F.glu(torch.cat((r, gate), dim=1), dim=1)
        t = t + r
    return F.linear(t, sd[p + "final_layer.weight"], sd[p + "final_layer.bias"])


def conv_residual_net(sd: StateDict, p: str, x: Tensor, context: Optional[Tensor], num_blocks: int = 2) -> Tensor:
    """ConvResidualNet.forward, nn/resnet.py:128-142 with ConvResidualBlock.forward :91-104."""
    if context is not None and context.dim() < x.dim():
        context = context[:, :, None, None].expand(-1, -1, x.shape[2], x.shape[3])  # :130-133
    if context is not None:
        x = torch.cat((x, context), dim=1)
    t = F.conv2d(x, sd[p + "initial_layer.weight"], sd[p + "initial_layer.bias"])
    for i in range(num_blocks):
        q = f"{p}blocks.{i}."
        r = F.conv2d(F.relu(t), sd[q + "conv_layers.0.weight"], sd[q + "conv_layers.0.bias"], padding=1)
        r = F.conv2d(F.relu(r), sd[q + "conv_layers.1.weight"], sd[q + "conv_layers.1.bias"], padding=1)
        if context is not None:
            gate = F.conv2d(context, sd[q + "context_layer.weight"], sd[q + "context_layer.bias"])
            r = F.glu(torch.cat((r, gate), dim=1), dim=1)
        t = t + r
    return F.conv2d(t, sd[p + "final_layer.weight"], sd[p + "final_layer.bias"])


# --------------------------------------------------------------------------------------
# (a4-a6) RQ coupling layer                 transforms/coupling.py
# --------------------------------------------------------------------------------------
def rq_coupling(sd: StateDict, p: str, x: Tensor, context: Optional[Tensor], inverse: bool, cfg: dict):
    """CouplingTransform.forward/inverse (coupling.py:58-129 / 131-185) specialised to
    PiecewiseRationalQuadraticCouplingTransform (:458-533) with linear tails."""
    k, t, hidden = cfg["num_bins"], cfg["tail_bound"], cfg["hidden"]
    id_idx, tr_idx = sd[p + "identity_features"], sd[p + "transform_features"]
    x_id, x_tr = x[:, id_idx, ...], x[:, tr_idx, ...]  # :65-66
    net = conv_residual_net if x.dim() == 4 else residual_net
    params = net(sd, p + "transform_net.", x_id, context)  # :108
    if x.dim() == 4:  # :266-268
        b, c, hh, ww = x_tr.shape
        params = params.reshape(b, c, -1, hh, ww).permute(0, 1, 3, 4, 2)
    else:  # :270-272
        params = params.reshape(x_tr.shape[0], x_tr.shape[1], -1)
    scale = np.sqrt(hidden)  # :506-511, an in-place division by a float64 numpy scalar
    uw = params[..., :k] / scale
    uh = params[..., k : 2 * k] / scale
    ud = params[..., 2 * k :]
    y_tr, lad = rqs(x_tr, uw, uh, ud, inverse=inverse, tail_bound=t)
    out = torch.empty_like(x)  # :116-118
    out[:, id_idx, ...] = x_id
    out[:, tr_idx, ...] = y_tr
    return out, sum_except_batch(lad)  # :279


# --------------------------------------------------------------------------------------
# (a9-a14) linear glue
# --------------------------------------------------------------------------------------
def lu_matrices(sd: StateDict, p: str, eps: float = 1e-3) -> Tuple[Tensor, Tensor, Tensor]:
    """LULinear._create_lower_upper (lu.py:44-54) and upper_diag (:118-120)."""
    diag_raw = sd[p + "unconstrained_upper_diag"]
    n = diag_raw.shape[0]
    lower = diag_raw.new_zeros(n, n)
    upper = diag_raw.new_zeros(n, n)
    li = np.tril_indices(n, k=-1)
    ui = np.triu_indices(n, k=1)
    di = np.diag_indices(n)
    lower[li[0], li[1]] = sd[p + "lower_entries"]
    lower[di[0], di[1]] = 1.0
    upper[ui[0], ui[1]] = sd[p + "upper_entries"]
    udiag = F.softplus(diag_raw) + eps
    upper[di[0], di[1]] = udiag
    return lower, upper, udiag


def lu_linear(sd: StateDict, p: str, x: Tensor, inverse: bool):
    """LULinear.forward_no_cache / inverse_no_cache (lu.py:56-95)."""
    lower, upper, udiag = lu_matrices(sd, p)
    logdet = torch.sum(torch.log(udiag))  # :122-128
    if not inverse:
        y = F.linear(F.linear(x, upper), lower, sd[p + "bias"])
        return y, logdet * x.new_ones(y.shape[0])
    y = (x - sd[p + "bias"]).t()
    y = torch.linalg.solve_triangular(lower, y, upper=False, unitriangular=True)
    y = torch.linalg.solve_triangular(upper, y, upper=True, unitriangular=False)
    y = y.t()
    return y, (-logdet) * x.new_ones(y.shape[0])


def permute(x: Tensor, perm: Tensor, inverse: bool, dim: int = 1) -> Tensor:
    """Permutation.forward/inverse (permutations.py:27-29,73,77-81)."""
    return torch.index_select(x, dim, torch.argsort(perm) if inverse else perm)


def one_by_one_conv(sd: StateDict, p: str, x: Tensor, inverse: bool):
    """OneByOneConvolution (conv.py:16-55): channel permutation, then the LU map applied to
    every pixel as a row of a [B*H*W, C] matrix; logabsdet summed over H*W copies."""
    perm = sd[p + "permutation._permutation"]
    b, c, hh, ww = x.shape
    if not inverse:
        x = permute(x, perm, False)
    rows = x.permute(0, 2, 3, 1).reshape(b * hh * ww, c)
    rows, lad = lu_linear(sd, p, rows, inverse)
    y = rows.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    lad = sum_except_batch(lad.reshape(b, hh, ww))
    if inverse:
        y = permute(y, perm, True)
    return y, lad


def actnorm(sd: StateDict, p: str, x: Tensor, inverse: bool):
    """ActNorm.forward/inverse with initialised parameters (normalization.py:99-131)."""
    log_scale, shift = sd[p + "log_scale"], sd[p + "shift"]
    shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
    scale = torch.exp(log_scale).view(shape)
    hw = x.shape[2] * x.shape[3] if x.dim() == 4 else 1
    ones = x.new_ones(x.shape[0])
    if inverse:
        return (x - shift.view(shape)) / scale, -hw * torch.sum(log_scale) * ones
    return scale * x + shift.view(shape), hw * torch.sum(log_scale) * ones


def actnorm_init(x: Tensor) -> Tuple[Tensor, Tensor]:
    """ActNorm._initialize (normalization.py:134-147) -> (log_scale, shift)."""
    if x.dim() == 4:
        x = x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])
    std = x.std(dim=0)
    mu = (x / std).mean(dim=0)
    return -torch.log(std), -mu


def squeeze(x: Tensor, inverse: bool, f: int = 2) -> Tensor:
    """SqueezeTransform (reshape.py:29-57)."""
    b, c, hh, ww = x.shape
    if not inverse:
        x = x.view(b, c, hh // f, f, ww // f, f).permute(0, 1, 3, 5, 2, 4).contiguous()
        return x.view(b, c * f * f, hh // f, ww // f)
    x = x.reshape(b, c // (f * f), f, f, hh, ww).permute(0, 1, 4, 2, 5, 3).contiguous()
    return x.view(b, c // (f * f), hh * f, ww * f)


def affine_scalar(sd: StateDict, p: str, x: Tensor, inverse: bool):
    """AffineScalarTransform (standard.py:41-57)."""
    scale, shift = sd[p + "_scale"], sd[p + "_shift"]
    num_dims = torch.prod(torch.tensor(x.shape[1:]), dtype=torch.float)
    log_scale = torch.log(torch.abs(scale))
    if inverse:
        return (x - shift) / scale, x.new_ones(x.shape[0]) * (-log_scale * num_dims).to(x.dtype)
    return x * scale + shift, x.new_ones(x.shape[0]) * (log_scale * num_dims).to(x.dtype)


def log_tanh(x: Tensor, inverse: bool, cut_point: float = 1.0):
    """LogTanh (nonlinearities.py:38-95), mask-free."""
    inv_cut = np.tanh(cut_point)
    alpha = (1 - np.tanh(np.tanh(cut_point))) / cut_point
    beta = np.exp((np.tanh(cut_point) - alpha * np.log(cut_point)) / alpha)
    if not inverse:
        right, left = x > cut_point, x < -cut_point
        mid = ~(right | left)
        xm = torch.where(mid, x, torch.zeros_like(x))
        xr = torch.where(right, x, torch.ones_like(x))
        xl = torch.where(left, x, -torch.ones_like(x))
        ym = torch.tanh(xm)
        y = torch.where(mid, ym, torch.where(right, alpha * torch.log(beta * xr), alpha * -torch.log(-beta * xl)))
        lad = torch.where(mid, torch.log(1 - ym ** 2), torch.where(right, torch.log(alpha / xr), torch.log(-alpha / xl)))
        return y, sum_except_batch(lad)
    right, left = x > inv_cut, x < -inv_cut
    mid = ~(right | left)
    xm = torch.where(mid, x, torch.zeros_like(x))
    xr = torch.where(right, x, torch.zeros_like(x))
    xl = torch.where(left, x, torch.zeros_like(x))
    y = torch.where(
        mid, 0.5 * torch.log((1 + xm) / (1 - xm)), torch.where(right, torch.exp(xr / alpha) / beta, -torch.exp(-xl / alpha) / beta)
    )
    lab = float(np.log(alpha * beta))
    lad = torch.where(mid, -torch.log(1 - xm ** 2), torch.where(right, -lab + xr / alpha, -lab - xl / alpha))
    return y, sum_except_batch(lad)


# --------------------------------------------------------------------------------------
# spec tree walker                           transforms/base.py, partial.py
# --------------------------------------------------------------------------------------
def apply(spec, sd: StateDict, p: str, x: Tensor, context: Optional[Tensor], inverse: bool):
    """Evaluate one node of the spec tree; returns (outputs, logabsdet[B])."""
    kind = spec[0]
    zeros = lambda: x.new_zeros(x.shape[0])
    if kind == "composite":  # CompositeTransform._cascade, base.py:71-76
        children = list(enumerate(spec[1]))
        total = zeros()
        for i, child in (reversed(children) if inverse else children):
            x, lad = apply(child, sd, f"{p}_transforms.{i}.", x, context, inverse)
            total = total + lad
        return x, total
    if kind == "multiscale":
        return _multiscale(spec, sd, p, x, context, inverse)
    if kind == "rq_coupling":
        ctx = context if spec[1].get("context", False) else None
        return rq_coupling(sd, p, x, ctx, inverse, spec[1])
    if kind == "lu":
        return lu_linear(sd, p, x, inverse)
    if kind == "conv1x1":
        return one_by_one_conv(sd, p, x, inverse)
    if kind == "perm":
        return permute(x, sd[p + "_permutation"], inverse), zeros()
    if kind == "actnorm":
        return actnorm(sd, p, x, inverse)
    if kind == "squeeze":
        return squeeze(x, inverse), zeros()
    if kind == "affine_scalar":
        return affine_scalar(sd, p, x, inverse)
    if kind == "logtanh":
        return log_tanh(x, inverse, spec[1])
    if kind == "identity":
        return x, zeros()
    if kind == "partial":  # PartialTransform, partial.py:42-88
        id_idx, tr_idx = sd[p + "identity_features"], sd[p + "transform_features"]
        y_tr, lad = apply(spec[1], sd, p + "transform.", x[:, tr_idx, ...], context, inverse)
        out = torch.empty_like(x)
        out[:, id_idx, ...] = x[:, id_idx, ...]
        out[:, tr_idx, ...] = y_tr
        return out, lad
    raise ValueError(f"unknown spec node {kind!r}")


def _multiscale(spec, sd, p, x, context, inverse):
    """MultiscaleCompositeTransform.forward (base.py:147-196) / inverse (:198-245)."""
    _, children, out_shapes = spec
    n = len(children)
    b = x.shape[0]
    total = x.new_zeros(b)
    if not inverse:
        pieces, hidden = [], x
        for i, child in enumerate(children):
            y, lad = apply(child, sd, f"{p}_transforms.{i}.", hidden, context, False)
            total = total + lad
            if i < n - 1:
                out, hidden = torch.chunk(y, 2, dim=1)
            else:
                out = y
            pieces.append(out.reshape(b, -1))
        return torch.cat(pieces, dim=-1), total
    sizes = [int(np.prod(s)) for s in out_shapes]
    offs = np.concatenate(([0], np.cumsum(sizes)))
    chunks = [x[:, offs[i] : offs[i + 1]].reshape(-1, *out_shapes[i]) for i in range(n)]
    hidden, lad = apply(children[-1], sd, f"{p}_transforms.{n - 1}.", chunks[-1], context, True)
    total = total + lad
    for i in range(n - 2, -1, -1):
        hidden, lad = apply(children[i], sd, f"{p}_transforms.{i}.", torch.cat([chunks[i], hidden], dim=1), context, True)
        total = total + lad
    return hidden, total


# --------------------------------------------------------------------------------------
# spec builders - the shapes experiments/architectures/* fixes
# --------------------------------------------------------------------------------------
def vector_transform_spec(dim, flow_steps, linear="permutation", num_bins=8, tail_bound=3.0, context=False, hidden=100):
    """create_vector_transform, architectures/vector_transforms.py:139-177."""

    def lin():
        if linear == "permutation":
            return ("perm",)
        if linear == "lu":
            return ("composite", [("perm",), ("lu",)])
        raise ValueError(linear)

    cfg = dict(num_bins=num_bins, tail_bound=float(tail_bound), hidden=hidden, context=context)
    steps = [("composite", [lin(), ("rq_coupling", cfg)]) for _ in range(flow_steps)]
    return ("composite", steps + [lin()])


def mlt_channel_mask(features, channels_per_level, resolution):
    """utils/various.py:453-469."""
    mask = np.zeros(features, dtype=bool)
    pos, total, res = 0, features, resolution
    for ch in channels_per_level:
        total //= 2
        res //= 2
        mask[pos : pos + ch * res * res] = True
        pos += total
    return mask


def image_transform_spec(
    c, h, w, levels, steps_per_level, num_bins, tail_bound, postprocessing, use_actnorm=True, context=False, hidden=100,
    postprocessing_layers=2,
):
    """create_image_transform, architectures/image_transforms.py:116-218, multi-scale, glow
    pre-processing; post-processing 'partial_mlp' (:240-258) or 'permutation' (:277-280)."""
    cfg = dict(num_bins=num_bins, tail_bound=float(tail_bound), hidden=hidden, context=context)
    children, out_shapes = [], []
    for level in range(levels):
        c, h, w = c * 4, h // 2, w // 2
        step = ([("actnorm",)] if use_actnorm else []) + [("conv1x1",), ("rq_coupling", cfg)]
        children.append(("composite", [("squeeze",)] + [("composite", list(step)) for _ in range(steps_per_level)] + [("conv1x1",)]))
        if level < levels - 1:
            out_shapes.append(((c + 1) // 2, h, w))
            c = c // 2
        else:
            out_shapes.append((c, h, w))
    if postprocessing == "partial_mlp":
        inner = [("lu",)]
        for _ in range(postprocessing_layers - 1):
            inner += [("logtanh", 1.0), ("lu",)]
        post = ("composite", [("partial", ("composite", inner)), ("perm",)])
    elif postprocessing == "permutation":
        post = ("perm",)
    else:
        raise ValueError(postprocessing)
    return ("composite", [("affine_scalar",), ("multiscale", children, out_shapes), post])


# --------------------------------------------------------------------------------------
# (a15-a17) model wrappers                  flows/manifold_flow.py, flows/flow.py
# --------------------------------------------------------------------------------------
def standard_normal_log_prob(u: Tensor) -> Tensor:
    """StandardNormal._log_prob (distributions/normal.py:18-23)."""
    n = int(np.prod(u.shape[1:]))
    return -0.5 * sum_except_batch(u ** 2) - 0.5 * n * np.log(2 * np.pi)


def rescaled_normal_log_prob(v: Tensor, std: float, clip: Optional[float]) -> Tensor:
    """RescaledNormal._log_prob (distributions/normal.py:52-59)."""
    m = int(np.prod(v.shape[1:]))
    if clip is not None:
        v = torch.clamp(v, -clip, clip)
    log_z = 0.5 * m * np.log(2 * np.pi) + m * np.log(std)
    return -0.5 * sum_except_batch(v ** 2) / std ** 2 - log_z


def mflow_forward(
    model: dict, sd: StateDict, x: Tensor, mode: str = "mf", context: Optional[Tensor] = None, return_hidden: bool = False
):
    """ManifoldFlow.forward (flows/manifold_flow.py:40-64): encode, decode, log_prob.

    ``model`` = dict(outer=spec, inner=spec, latent_dim=n, data_shape=tuple, pie_epsilon,
    clip_pie, conditional_outer).  Modes: projection, pie, mf-fixed-manifold, slice, mf.
    """
    if mode == "mf":
        out = mflow_log_prob_mf(model, sd, x, context)
        return out + (None,) if return_hidden else out
    n = model["latent_dim"]
    eps = model.get("pie_epsilon", 1e-2)
    clip = model.get("clip_pie", None)
    clip = None if not clip else clip * eps  # manifold_flow.py:28
    octx = context if model.get("conditional_outer", False) else None
    b = x.shape[0]
    # _encode :96-102
    h, ld_outer = apply(model["outer"], sd, "outer_transform.", x, octx, False)
    h = h.reshape(b, -1)
    h_man, h_orth = h[:, :n], h[:, n:]  # ProjectionSplit.forward, projections.py:25-35
    u, ld_inner = apply(model["inner"], sd, "inner_transform.", h_man, context, False)
    # _decode :104-122
    h_dec, inv_ld_inner = apply(model["inner"], sd, "inner_transform.", u, context, True)
    h_full = torch.cat((h_dec, h_dec.new_zeros(b, h.shape[1] - n)), dim=1)  # projections.py:37-47
    # ManifoldFlow builds ProjectionSplit from the *total* dims (ints), so it stays in "vector" mode
    # and the flat [B, D] tensor goes straight into outer_transform.inverse (manifold_flow.py:30)
    x_reco, inv_ld_outer = apply(model["outer"], sd, "outer_transform.", h_full, octx, True)
    # _log_prob :124-157
    if mode in ("pie", "mf-fixed-manifold"):
        log_prob = standard_normal_log_prob(u) + rescaled_normal_log_prob(h_orth, eps, clip) + ld_outer + ld_inner
    elif mode == "slice":
        log_prob = standard_normal_log_prob(u) + rescaled_normal_log_prob(torch.zeros_like(h_orth), eps, clip)
        log_prob = log_prob - inv_ld_outer - inv_ld_inner
    elif mode == "projection":
        log_prob = None
    else:
        raise ValueError(mode)
    if return_hidden:
        return x_reco, log_prob, u, torch.cat((h_man, h_orth), -1)
    return x_reco, log_prob, u


def batch_jacobian_wrt(outputs: Tensor, inputs: Tensor) -> Tensor:
    """[B, D_out, D_in] per-sample Jacobian.  Samples are independent, so one reverse pass per
    output coordinate of the batch sum yields every sample's row (the reference loops over
    B * D_out scalars instead: utils/various.py:22-45,68-87)."""
    rows = []
    for i in range(outputs.shape[1]):
        (g,) = torch.autograd.grad(outputs[:, i].sum(), inputs, retain_graph=True, create_graph=False)
        rows.append(g)
    return torch.stack(rows, dim=1)


def permutation_jacobian_reference(perm: Tensor, inverse: bool, batch: int, dtype) -> Tensor:
    """The matrix Permutation._permute returns for full_jacobian=True (permutations.py:38-71):
    ``jacobian[permutation, arange] = 1``.  NOTE: for outputs[j] = inputs[permutation[j]] the true
    Jacobian has ones at [j, permutation[j]]; the reference's matrix is its transpose, which only
    coincides for involutions.  ``mode="mf"`` likelihoods of the reference are computed with this
    matrix, so parity requires reproducing it [probed: 3-cycle permutation, autograd vs reference]."""
    idx = torch.argsort(perm) if inverse else perm
    n = perm.numel()
    jac = torch.zeros(n, n, dtype=dtype, device=perm.device)
    jac[idx, torch.arange(n, device=perm.device)] = 1.0
    return jac.unsqueeze(0).expand(batch, -1, -1)


def apply_with_jacobian(spec, sd: StateDict, p: str, x: Tensor, context: Optional[Tensor], inverse: bool):
    """(outputs, jacobian[B, D, D]) the way CompositeTransform._cascade(full_jacobian=True) builds it
    (base.py:52-69): per-layer Jacobians multiplied left to right.  Leaf layers: autograd Jacobian of
    the layer (what CouplingTransform does via batch_jacobian, coupling.py:71-105,141-170, and what
    LULinear states analytically, lu.py:69-71,89-91); permutations: the reference's own matrix."""
    kind = spec[0]
    if kind == "composite":
        children = list(enumerate(spec[1]))
        total = None
        for i, child in (reversed(children) if inverse else children):
            x, jac = apply_with_jacobian(child, sd, f"{p}_transforms.{i}.", x, context, inverse)
            total = jac if total is None else torch.bmm(jac, total)
        return x, total
    if kind == "perm":
        perm = sd[p + "_permutation"]
        return permute(x, perm, inverse), permutation_jacobian_reference(perm, inverse, x.shape[0], x.dtype)
    xin = x.detach().requires_grad_(True)
    y, _ = apply(spec, sd, p, xin, context, inverse)
    return y.detach(), batch_jacobian_wrt(y, xin)


def mflow_log_prob_mf(model: dict, sd: StateDict, x: Tensor, context: Optional[Tensor] = None):
    """mode="mf" exactly as the reference composes it (manifold_flow.py:115-120,140-147):
    J = chain of per-layer Jacobians of outer^{-1} restricted to the first n columns,
    log p = log N(u) - 1/2 log det(J^T J) - inv_log_det_inner.  Returns (x_reco, log_prob, u)."""
    n = model["latent_dim"]
    octx = context if model.get("conditional_outer", False) else None
    b = x.shape[0]
    with torch.no_grad():
        h, _ = apply(model["outer"], sd, "outer_transform.", x, octx, False)
        h = h.reshape(b, -1)
        u, _ = apply(model["inner"], sd, "inner_transform.", h[:, :n], context, False)
        h_dec, inv_ld_inner = apply(model["inner"], sd, "inner_transform.", u, context, True)
        h_full = torch.cat((h_dec, h_dec.new_zeros(b, h.shape[1] - n)), dim=1)
    x_reco, jac = apply_with_jacobian(model["outer"], sd, "outer_transform.", h_full, octx, True)
    jac = jac[:, :, :n]
    jtj = torch.bmm(jac.transpose(-2, -1), jac)
    log_prob = standard_normal_log_prob(u) - 0.5 * torch.slogdet(jtj)[1] - inv_ld_inner
    return x_reco.detach(), log_prob.detach(), u.detach()


def flow_forward(model: dict, sd: StateDict, x: Tensor, context: Optional[Tensor] = None, decode: bool = True):
    """Flow.forward (flows/flow.py:26-39) / Flow.log_prob (:53-63 when decode=False)."""
    u, log_det = apply(model["transform"], sd, "transform.", x, context, False)
    log_prob = standard_normal_log_prob(u) + log_det
    x_reco = None
    if decode:
        x_reco, _ = apply(model["transform"], sd, "transform.", u, context, True)
    return x_reco, log_prob, u


# --------------------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------------------
def config_model(name: str, **over) -> dict:
    """Architecture dicts for cfg1..cfg5 (and reduced image variants for quick tests)."""
    if name == "cfg1":  # spherical_gaussian 3->2, train.py defaults :41-60
        m = dict(kind="mf", data_shape=(3,), latent_dim=2, outer_layers=5, inner_layers=5, linear="permutation", num_bins=8,
                 tail_bound=3.0, context_features=0, pie_epsilon=0.01, clip_pie=None)
    elif name == "cfg2":  # power, configs/train_power_march.config
        m = dict(kind="mf", data_shape=(3,), latent_dim=2, outer_layers=5, inner_layers=5, linear="permutation", num_bins=10,
                 tail_bound=6.0, context_features=1, pie_epsilon=0.01, clip_pie=None)
    elif name == "cfg3":  # lhc40d, configs/train_mf_lhc_june.config
        m = dict(kind="mf", data_shape=(40,), latent_dim=14, outer_layers=20, inner_layers=15, linear="lu", num_bins=11,
                 tail_bound=10.0, context_features=2, pie_epsilon=0.01, clip_pie=None)
    elif name == "cfg4":  # gan64d, configs/train_mf_gan64d_april.config
        m = dict(kind="image_mf", data_shape=(3, 64, 64), latent_dim=64, levels=4, outer_layers=20, inner_layers=8, linear="lu",
                 num_bins=11, tail_bound=10.0, context_features=1, pie_epsilon=0.01, clip_pie=None, actnorm=True)
    elif name == "cfg5":  # celeba ambient flow, configs/train_flow_celeba_april.config
        m = dict(kind="image_flow", data_shape=(3, 64, 64), levels=4, outer_layers=20, inner_layers=8, num_bins=11, tail_bound=10.0,
                 context_features=0, actnorm=True)
    else:
        raise ValueError(name)
    m.update(over)
    return build_model(m)


def build_model(m: dict) -> dict:
    ctx = m.get("context_features", 0) > 0
    if m["kind"] == "mf":
        d = m["data_shape"][0]
        m["outer"] = vector_transform_spec(d, m["outer_layers"], m["linear"], m["num_bins"], m["tail_bound"], context=False)
        m["inner"] = vector_transform_spec(m["latent_dim"], m["inner_layers"], m["linear"], m["num_bins"], m["tail_bound"], context=ctx)
    elif m["kind"] == "image_mf":
        c, h, w = m["data_shape"]
        m["outer"] = image_transform_spec(c, h, w, m["levels"], m["outer_layers"] // m["levels"], m["num_bins"], m["tail_bound"],
                                          "partial_mlp", use_actnorm=m.get("actnorm", True), context=False)
        m["inner"] = vector_transform_spec(m["latent_dim"], m["inner_layers"], m["linear"], m["num_bins"], m["tail_bound"], context=ctx)
    elif m["kind"] == "image_flow":
        c, h, w = m["data_shape"]
        steps = (m["inner_layers"] + m["outer_layers"]) // m["levels"]
        m["transform"] = image_transform_spec(c, h, w, m["levels"], steps, m["num_bins"], m["tail_bound"], "permutation",
                                              use_actnorm=m.get("actnorm", True), context=ctx)
    else:
        raise ValueError(m["kind"])
    return m
