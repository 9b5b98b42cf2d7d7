"""CPU: the oracle reproduces the golden vectors frozen from the live reference
(oracle/make_golden.py).  This is the pin that makes the oracle trustworthy."""
import numpy as np
import pytest
import torch

import golden_util as gu
import mflow_oracle as oracle

S = gu.load_npz("spline.npz")
L = gu.load_npz("layers.npz")


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", [str(c) for c in S["cases"]])
@pytest.mark.parametrize("inverse", [False, True])
def test_spline_matches_reference_bitwise(name, inverse):
    t = float(name.split("_")[1][1:])
    tag = "inv" if inverse else "fwd"
    x, w, h, d = (_t(S[f"{name}/{k}"]) for k in "xwhd")
    y, lad = oracle.rqs(x, w, h, d, inverse, t)
    # same ATen ops, same order -> identical bits on CPU
    assert torch.equal(y, _t(S[f"{name}/{tag}_y"]))
    assert torch.equal(lad, _t(S[f"{name}/{tag}_lad"]))
    y64, lad64 = oracle.rqs(x.double(), w.double(), h.double(), d.double(), inverse, t)
    assert torch.equal(y64, _t(S[f"{name}/{tag}_y64"]))
    assert torch.equal(lad64, _t(S[f"{name}/{tag}_lad64"]))
    # tails: exact identity with zero logabsdet
    out = (x < -t) | (x > t)
    assert out.any()
    assert torch.equal(y[out], x[out]) and torch.all(lad[out] == 0)


SB = gu.load_npz("spline_box.npz")


@pytest.mark.parametrize("name", [str(c) for c in SB["cases"]])
def test_box_spline_matches_reference_bitwise(name):
    """rational_quadratic_spline on an arbitrary box (rational_quadratic.py:67-168)."""
    left, right, bottom, top = (float(v) for v in SB[f"{name}/box"])
    x, w, h, d = (_t(SB[f"{name}/{k}"]) for k in "xwhd")
    y, lad = oracle.rqs_box(x, w, h, d, False, left, right, bottom, top)
    assert torch.equal(y, _t(SB[f"{name}/fwd_y"])) and torch.equal(lad, _t(SB[f"{name}/fwd_lad"]))
    y64, lad64 = oracle.rqs_box(x.double(), w.double(), h.double(), d.double(), False, left, right, bottom, top)
    assert torch.equal(y64, _t(SB[f"{name}/fwd_y64"])) and torch.equal(lad64, _t(SB[f"{name}/fwd_lad64"]))
    assert float(y64.min()) >= bottom - 1e-9 and float(y64.max()) <= top + 1e-9
    if f"{name}/inv_in" in SB:
        xi, ladi = oracle.rqs_box(_t(SB[f"{name}/inv_in"]), w, h, d, True, left, right, bottom, top)
        assert torch.equal(xi, _t(SB[f"{name}/inv_x"])) and torch.equal(ladi, _t(SB[f"{name}/inv_lad"]))
    # autograd of the restatement reproduces the reference's gradients
    xx, ww, hh, dd = (t.double().clone().requires_grad_(True) for t in (x, w, h, d))
    yy, ll = oracle.rqs_box(xx, ww, hh, dd, False, left, right, bottom, top)
    (yy.sum() + ll.sum()).backward()
    for got, key in ((xx.grad, "g_x64"), (ww.grad, "g_w64"), (hh.grad, "g_h64"), (dd.grad, "g_d64")):
        torch.testing.assert_close(got, _t(SB[f"{name}/{key}"]), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("t", [3, 10])
def test_searchsorted_bitexact(t):
    knots, v, bins = (_t(S[f"search_T{t}/{k}"]) for k in ("knots", "values", "bins"))
    kk = knots[:, None, :].expand(-1, v.shape[1], -1)
    assert torch.equal(oracle.search_bins(kk, v), bins)


def _layer_spec(name):
    if name.startswith("vcoup") or name.startswith("icoup"):
        parts = name.split("_")
        if name.startswith("vcoup"):
            k, t, hidden = int(parts[2][1:]), float(parts[3][1:]), 16
            ctx = parts[4] != "ctx0"
        else:
            k, t, hidden = 11, 10.0, int(parts[4][1:])
            ctx = parts[3] != "ctx0"
        return ("rq_coupling", dict(num_bins=k, tail_bound=t, hidden=hidden, context=ctx))
    return {
        "actnorm_vec": ("actnorm",), "actnorm_img": ("actnorm",), "conv1x1_C12": ("conv1x1",), "conv1x1_C96": ("conv1x1",),
        "lu_D14": ("lu",), "lu_D40": ("lu",), "perm_D40": ("perm",), "squeeze": ("squeeze",), "affine_scalar": ("affine_scalar",),
        "logtanh": ("logtanh", 1.0),
        "partial_mlp": ("composite", [("partial", ("composite", [("lu",), ("logtanh", 1.0), ("lu",)])), ("perm",)]),
    }[name]


LAYER_CASES = [str(c) for c in L["cases"] if not str(c).startswith("actnorm_init")]


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_matches_reference(name):
    sd = {k.split("/sd/")[1]: _t(v) for k, v in L.items() if k.startswith(name + "/sd/")}
    x = _t(L[f"{name}/x"])
    ctx = _t(L[f"{name}/ctx"]) if f"{name}/ctx" in L else None
    spec = _layer_spec(name)
    y, lad = oracle.apply(spec, sd, "", x, ctx, False)
    torch.testing.assert_close(y, _t(L[f"{name}/y"]), rtol=0, atol=0)
    torch.testing.assert_close(lad, _t(L[f"{name}/lad"]), rtol=0, atol=0)
    xi, ladi = oracle.apply(spec, sd, "", _t(L[f"{name}/y"]), ctx, True)
    # the reference inverts LU layers with torch.triangular_solve; solve_triangular is the same LAPACK call
    torch.testing.assert_close(xi, _t(L[f"{name}/xi"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ladi, _t(L[f"{name}/ladi"]), rtol=1e-6, atol=1e-6)
    # invariants: round trip and cancelling log-dets
    torch.testing.assert_close(xi, x, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(lad + ladi, torch.zeros_like(lad), rtol=0, atol=5e-4)


@pytest.mark.parametrize("name", ["actnorm_init_vec", "actnorm_init_img"])
def test_actnorm_init(name):
    x = _t(L[f"{name}/x"])
    log_scale, shift = oracle.actnorm_init(x)
    torch.testing.assert_close(log_scale, _t(L[f"{name}/log_scale"]), rtol=0, atol=0)
    torch.testing.assert_close(shift, _t(L[f"{name}/shift"]), rtol=0, atol=0)
    y, lad = oracle.actnorm({"log_scale": log_scale, "shift": shift}, "", x, False)
    torch.testing.assert_close(y, _t(L[f"{name}/y"]), rtol=0, atol=0)
    torch.testing.assert_close(lad, _t(L[f"{name}/lad"]), rtol=0, atol=0)
