#!/usr/bin/env python
"""Freeze golden vectors from the LIVE reference (build container only).

TEST INFRASTRUCTURE.  Imports the unmodified reference from ``/root/reference`` (read-only),
runs its own PyTorch CPU path on seeded inputs and writes small fixtures to ``tests/golden/``:

* ``spline.npz``      - ``unconstrained_rational_quadratic_spline`` / ``searchsorted`` outputs
* ``layers.npz``      - single transforms (couplings, ActNorm, 1x1 conv, LU, ...) with their weights
* ``models.json/npz`` - cfg1..cfg5 models: SHA-256 of the seeded ``state_dict`` (weights are not
  stored; both the reference and the product rebuild them from the seed), outputs of every
  supported mode after a deterministic perturbation of the weights, selected gradients.

The reference needs three shims on numpy >= 1.24 (``np.int``, ``np.float``, ``np.product``) and stub
``matplotlib`` / ``configargparse`` modules; nothing in it is modified.

Usage:  python oracle/make_golden.py            (about a minute on 8 cores)
"""
import json
import os
import sys
import types
import warnings

import numpy as np

np.int = int
np.float = float
np.product = np.prod
for _name in ("matplotlib", "matplotlib.pyplot", "configargparse"):
    sys.modules.setdefault(_name, types.ModuleType(_name))

import torch  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/experiments")
sys.path.insert(0, os.path.join(REPO, "tests"))
warnings.filterwarnings("ignore")

import golden_util as gu  # noqa: E402
from manifold_flow import nn as nn_  # noqa: E402  (the reference)
from manifold_flow import transforms  # noqa: E402
from manifold_flow.transforms.splines import rational_quadratic as ref_rq  # noqa: E402
from manifold_flow.utils import various  # noqa: E402

OUT = gu.GOLDEN_DIR
os.makedirs(OUT, exist_ok=True)


def npy(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ spline-level
def spline_cases():
    out = {}
    g = torch.Generator().manual_seed(99)
    n = 64
    case_list = []
    for k, t in ((8, 3.0), (10, 6.0), (11, 10.0), (8, 10.0), (11, 3.0)):
        for kind in ("normal", "zeros", "saturated"):
            name = f"K{k}_T{int(t)}_{kind}"
            case_list.append(name)
            if kind == "normal":
                w, h, d = (torch.randn(n, k, generator=g), torch.randn(n, k, generator=g), torch.randn(n, k - 1, generator=g))
            elif kind == "zeros":
                w, h, d = torch.zeros(n, k), torch.zeros(n, k), torch.zeros(n, k - 1)
            else:
                w, h, d = (30 * torch.randn(n, k, generator=g), 30 * torch.randn(n, k, generator=g), 8 * torch.randn(n, k - 1, generator=g))
            x = torch.randn(n, generator=g) * (t / 2)
            x[0], x[1], x[2], x[3] = t, -t, t + 1e-3, -t - 1e-3  # closed interval ends and just outside
            x[4], x[5] = 0.0, float(np.nextafter(np.float32(t), np.float32(0)))
            out[f"{name}/x"], out[f"{name}/w"], out[f"{name}/h"], out[f"{name}/d"] = npy(x), npy(w), npy(h), npy(d)
            for inv in (False, True):
                tag = "inv" if inv else "fwd"
                y, lad = ref_rq.unconstrained_rational_quadratic_spline(
                    x.clone(), w.clone(), h.clone(), d.clone(), inverse=inv, tails="linear", tail_bound=t
                )
                y64, lad64 = ref_rq.unconstrained_rational_quadratic_spline(
                    x.double(), w.double(), h.double(), d.double(), inverse=inv, tails="linear", tail_bound=t
                )
                out[f"{name}/{tag}_y"], out[f"{name}/{tag}_lad"] = npy(y), npy(lad)
                out[f"{name}/{tag}_y64"], out[f"{name}/{tag}_lad64"] = npy(y64), npy(lad64)
    # searchsorted on given knots: values on knots, +-1 ulp, at the ends, past the bumped end
    for t in (3.0, 10.0):
        k = 8
        knots = torch.sort(torch.rand(16, k + 1, generator=g) * 2 * t - t, dim=-1)[0]
        knots[:, 0], knots[:, -1] = -t, t
        vals = []
        for j in range(k + 1):
            kj = npy(knots[:, j]).astype(np.float32)
            vals += [kj, np.nextafter(kj, np.float32(-1e9)), np.nextafter(kj, np.float32(1e9))]
        vals = np.clip(np.stack(vals, axis=1), -t, np.float32(t))  # [16, 3(k+1)]
        vals[:, -1] = np.float32(t)
        v = torch.from_numpy(vals.astype(np.float32))
        kk = knots[:, None, :].expand(-1, v.shape[1], -1).contiguous()
        bins = various.searchsorted(kk.clone(), v)
        out[f"search_T{int(t)}/knots"], out[f"search_T{int(t)}/values"], out[f"search_T{int(t)}/bins"] = npy(knots), npy(v), npy(bins)
    out["cases"] = np.array(case_list)
    np.savez_compressed(os.path.join(OUT, "spline.npz"), **out)
    print("spline.npz:", len(case_list), "cases")


def spline_box_cases():
    """``rational_quadratic_spline`` on arbitrary boxes (rational_quadratic.py:67-168): outputs, logabsdet (fp32 / fp64)
    and the reference's autograd gradients of sum(y) + sum(logabsdet).  Own file and own generator so that the
    fixtures above stay bit-identical."""
    out, names = {}, []
    g = torch.Generator().manual_seed(4242)
    n = 96
    for k, (left, right, bottom, top) in ((8, (0.0, 1.0, 0.0, 1.0)), (5, (-2.0, 3.0, -1.0, 0.5)), (11, (-10.0, 10.0, -10.0, 10.0)),
                                            (3, (1.0, 4.0, -7.0, 9.0)), (16, (-0.5, 0.5, 2.0, 2.5))):
        name = f"box_K{k}_{left:g}_{right:g}_{bottom:g}_{top:g}"
        names.append(name)
        w, h, d = (torch.randn(n, k, generator=g) * 1.5, torch.randn(n, k, generator=g) * 1.5, torch.randn(n, k + 1, generator=g) * 2)
        x = left + (right - left) * torch.rand(n, generator=g)
        x[0], x[1] = left, right
        out[f"{name}/box"] = np.array([left, right, bottom, top], dtype=np.float64)
        out[f"{name}/x"], out[f"{name}/w"], out[f"{name}/h"], out[f"{name}/d"] = npy(x), npy(w), npy(h), npy(d)
        for dt, suffix in ((torch.float32, ""), (torch.float64, "64")):
            xx, ww, hh, dd = (t.to(dt).clone().requires_grad_(True) for t in (x, w, h, d))
            y, lad = ref_rq.rational_quadratic_spline(xx, ww, hh, dd, inverse=False, left=left, right=right, bottom=bottom, top=top)
            (y.sum() + lad.sum()).backward()
            out[f"{name}/fwd_y{suffix}"], out[f"{name}/fwd_lad{suffix}"] = npy(y), npy(lad)
            out[f"{name}/g_x{suffix}"], out[f"{name}/g_w{suffix}"] = npy(xx.grad), npy(ww.grad)
            out[f"{name}/g_h{suffix}"], out[f"{name}/g_d{suffix}"] = npy(hh.grad), npy(dd.grad)
            # inverse on the forward outputs clipped into [max(left, bottom), min(right, top)] when the boxes overlap,
            # else on fresh points of [bottom, top] (only valid when that lies inside [left, right], :85-86 checks x-range)
            yy = (bottom + (top - bottom) * torch.rand(n, generator=torch.Generator().manual_seed(k))).to(dt)
            if bottom >= left and top <= right:
                xi, ladi = ref_rq.rational_quadratic_spline(yy, ww.detach(), hh.detach(), dd.detach(), inverse=True, left=left, right=right,
                                                            bottom=bottom, top=top)
                out[f"{name}/inv_in{suffix}"], out[f"{name}/inv_x{suffix}"], out[f"{name}/inv_lad{suffix}"] = npy(yy), npy(xi), npy(ladi)
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "spline_box.npz"), **out)
    print("spline_box.npz:", len(names), "cases")


# ------------------------------------------------------------------ layer-level
def run_layer(name, layer, x, ctx, out, takes_ctx=True, has_inverse=True):
    layer.eval()
    sd = gu.perturb_state_dict(layer.state_dict(), seed=777)
    layer.load_state_dict(sd)
    for k, v in sd.items():
        out[f"{name}/sd/{k}"] = npy(v)
    out[f"{name}/x"] = npy(x)
    if ctx is not None:
        out[f"{name}/ctx"] = npy(ctx)
    with torch.no_grad():
        y, lad = layer(x, context=ctx) if takes_ctx else layer(x)
        out[f"{name}/y"], out[f"{name}/lad"] = npy(y), npy(lad)
        if has_inverse:
            xi, ladi = layer.inverse(y, context=ctx) if takes_ctx else layer.inverse(y)
            out[f"{name}/xi"], out[f"{name}/ladi"] = npy(xi), npy(ladi)


def layer_cases():
    out = {}
    torch.manual_seed(31)
    names = []

    def vec_coupling(d, k, t, ctx_features, even, hidden=16):
        net = lambda i, o: nn_.ResidualNet(i, o, hidden_features=hidden, context_features=ctx_features, num_blocks=2)
        return transforms.PiecewiseRationalQuadraticCouplingTransform(
            mask=various.create_alternating_binary_mask(d, even=even), transform_net_create_fn=net, num_bins=k, tails="linear", tail_bound=t
        )

    for d, k, t, cf in ((2, 8, 3.0, None), (3, 10, 6.0, 1), (14, 11, 10.0, 2), (40, 11, 10.0, None), (64, 11, 10.0, 1)):
        for even in (True, False):
            name = f"vcoup_D{d}_K{k}_T{int(t)}_ctx{cf or 0}_{'even' if even else 'odd'}"
            x = torch.randn(7, d) * (t / 2.5)
            ctx = torch.randn(7, cf) if cf else None
            run_layer(name, vec_coupling(d, k, t, cf, even), x, ctx, out)
            names.append(name)

    def img_coupling(c, k, t, ctx_channels, hidden):
        net = lambda i, o: nn_.ConvResidualNet(i, o, hidden_channels=hidden, context_channels=ctx_channels, num_blocks=2)
        return transforms.PiecewiseRationalQuadraticCouplingTransform(
            mask=various.create_mid_split_binary_mask(c), transform_net_create_fn=net, num_bins=k, tails="linear", tail_bound=t,
            min_bin_width=0.001, min_bin_height=0.001, min_derivative=0.001,
        )

    for c, hw, cf, hidden in ((12, 8, None, 16), (12, 8, 1, 16), (24, 4, None, 8), (48, 4, None, 8), (96, 2, None, 8)):
        name = f"icoup_C{c}_HW{hw}_ctx{cf or 0}_h{hidden}"
        x = torch.randn(3, c, hw, hw) * 4.5
        ctx = torch.randn(3, cf) if cf else None
        run_layer(name, img_coupling(c, 11, 10.0, cf, hidden), x, ctx, out)
        names.append(name)

    run_layer("actnorm_vec", _initialized(transforms.ActNorm(14)), torch.randn(9, 14), None, out)
    run_layer("actnorm_img", _initialized(transforms.ActNorm(12)), torch.randn(4, 12, 4, 4), None, out)
    run_layer("conv1x1_C12", transforms.OneByOneConvolution(12), torch.randn(3, 12, 4, 4), None, out)
    run_layer("conv1x1_C96", transforms.OneByOneConvolution(96), torch.randn(2, 96, 2, 2), None, out)
    run_layer("lu_D14", transforms.LULinear(14, identity_init=True), torch.randn(9, 14), None, out)
    run_layer("lu_D40", transforms.LULinear(40, identity_init=True), torch.randn(5, 40), None, out)
    run_layer("perm_D40", transforms.RandomPermutation(40), torch.randn(5, 40), None, out)
    run_layer("squeeze", transforms.SqueezeTransform(), torch.randn(3, 3, 8, 8), None, out, takes_ctx=True)
    run_layer("affine_scalar", transforms.AffineScalarTransform(scale=1.0 / 256, shift=-0.5), torch.rand(3, 3, 4, 4) * 256, None, out)
    run_layer("logtanh", transforms.LogTanh(cut_point=1), torch.randn(6, 20) * 2.0, None, out)
    mask = various.create_mlt_channel_mask(3 * 8 * 8, channels_per_level=2 * np.array([1, 2], dtype=int), resolution=8)
    pdim = int(mask.sum())
    partial = transforms.CompositeTransform(
        [
            transforms.PartialTransform(
                mask, transforms.CompositeTransform([transforms.LULinear(pdim), transforms.LogTanh(cut_point=1), transforms.LULinear(pdim)])
            ),
            transforms.MaskBasedPermutation(mask),
        ]
    )
    run_layer("partial_mlp", partial, torch.randn(4, 3 * 8 * 8), None, out)
    names += ["actnorm_vec", "actnorm_img", "conv1x1_C12", "conv1x1_C96", "lu_D14", "lu_D40", "perm_D40", "squeeze", "affine_scalar",
              "logtanh", "partial_mlp"]

    # ActNorm data-dependent initialisation (normalization.py:134-147)
    for nm, x in (("actnorm_init_vec", torch.randn(50, 6) * 3 + 1), ("actnorm_init_img", torch.randn(8, 12, 4, 4) * 2 - 0.5)):
        layer = transforms.ActNorm(x.shape[1])
        layer.train()
        y, lad = layer(x)
        out[f"{nm}/x"], out[f"{nm}/y"], out[f"{nm}/lad"] = npy(x), npy(y), npy(lad)
        out[f"{nm}/log_scale"], out[f"{nm}/shift"] = npy(layer.log_scale), npy(layer.shift)
        names.append(nm)
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "layers.npz"), **out)
    print("layers.npz:", len(names), "cases")


def _initialized(actnorm):
    actnorm.initialized = True
    return actnorm


# ------------------------------------------------------------------ model-level
def model_cases():
    from architectures.create_model import create_model  # the reference factory

    meta, arrays = {}, {}
    for name in ("cfg1", "cfg2", "cfg3", "cfg4s", "cfg5s", "cfg4", "cfg5"):
        args, sim = gu.config_args(name), gu.FakeSimulator(name)
        torch.manual_seed(gu.SEED)
        model = create_model(args, sim)
        sd0 = model.state_dict()
        entry = dict(
            keys=[[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd0.items()],
            hashes=gu.state_dict_hashes(sd0),
            num_parameters=int(sum(p.numel() for p in model.parameters())),
        )
        model.load_state_dict(gu.perturb_state_dict(sd0))
        for m in model.modules():
            if isinstance(m, transforms.ActNorm):
                m.initialized = True
        model.eval()
        x, ctx = gu.config_inputs(name)
        is_flow = args.algorithm == "flow"
        modes = [None] if is_flow else ["projection", "pie", "mf-fixed-manifold", "slice"]
        if not is_flow and not sim.is_image():
            modes.append("mf")
        for mode in modes:
            model.zero_grad()
            xx = x.clone()
            if is_flow:
                x_reco, log_prob, u = model(xx, context=ctx)
                tag = "flow"
            else:
                x_reco, log_prob, u = model(xx, mode=mode, context=ctx)
                tag = mode
            if tag in ("flow", "projection"):
                arrays[f"{name}/x_reco"] = npy(x_reco)
                arrays[f"{name}/u"] = npy(u)
            if log_prob is not None:
                arrays[f"{name}/{tag}/log_prob"] = npy(log_prob)
            if tag in ("flow", "mf-fixed-manifold"):
                (-log_prob.mean()).backward()
                gn = {}
                for k, p in model.named_parameters():
                    if p.grad is None:
                        continue
                    gn[k] = float(p.grad.double().norm())
                    if p.numel() <= 2048 and (k.endswith("bias") or k.endswith("log_scale") or k.endswith("shift")
                                              or k.endswith("unconstrained_upper_diag")) and ("_transforms.0." in k or "_transforms.1." in k):
                        arrays[f"{name}/grad/{k}"] = npy(p.grad)
                entry["grad_norms"] = gn
            if tag == "projection":
                model.zero_grad()
                x_reco2, _, _ = model(x.clone(), mode="projection", context=ctx)
                torch.nn.functional.mse_loss(x_reco2, x).backward()
                entry["grad_norms_projection"] = {k: float(p.grad.double().norm()) for k, p in model.named_parameters() if p.grad is not None}
        # float64 truth (tolerances of the fp32 kernels are judged against the reference's own
        # fp32-vs-fp64 error): likelihood, latents, reconstruction and gradient norms
        torch.set_default_dtype(torch.float64)
        try:
            m64 = model.double()
            m64.zero_grad()
            c64 = None if ctx is None else ctx.double()
            if is_flow:
                xr64, lp64, u64 = m64(x.double(), context=c64)
            else:
                xr64, lp64, u64 = m64(x.double(), mode="mf-fixed-manifold", context=c64)
            (-lp64.mean()).backward()
            entry["grad_norms64"] = {k: float(p.grad.norm()) for k, p in m64.named_parameters() if p.grad is not None}
            arrays[f"{name}/log_prob64"], arrays[f"{name}/u64"] = npy(lp64), npy(u64)
            arrays[f"{name}/x_reco64"] = npy(xr64).astype(np.float32) if xr64.numel() > 4096 else npy(xr64)
            for key in [k for k in arrays if k.startswith(f"{name}/grad/")]:
                pname = key.split("/grad/")[1]
                arrays[f"{name}/grad64/{pname}"] = npy(dict(m64.named_parameters())[pname].grad)
        finally:
            torch.set_default_dtype(torch.float32)
        meta[name] = entry
        print(name, "params", entry["num_parameters"], "modes", modes)
    with open(os.path.join(OUT, "models.json"), "w") as f:
        json.dump(meta, f)
    np.savez_compressed(os.path.join(OUT, "models.npz"), **arrays)


if __name__ == "__main__":
    torch.set_num_threads(8)
    spline_cases()
    spline_box_cases()
    layer_cases()
    model_cases()
    print("golden fixtures written to", OUT)
