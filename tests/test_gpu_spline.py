"""GPU parity of kernel (a), through the C ABI: fused RQ-spline forward / inverse / backward in both
data layouts against the golden vectors frozen from the reference and against the oracle.

Tolerances (BASELINE.json:north_star): bin indices bit-exact on given knots; outputs and logabsdet
within 1e-5 relative (+1e-5 absolute floor for values near zero) in fp32; tails exact identity."""
import numpy as np
import pytest
import torch

import golden_util as gu
import mflow_oracle as oracle

pytestmark = pytest.mark.gpu

S = gu.load_npz("spline.npz")
RTOL, ATOL = 1e-5, 1e-5


def _t(a, dev=None):
    t = torch.from_numpy(np.asarray(a))
    return t if dev is None else t.to(dev)


@pytest.mark.parametrize("t", [3, 10])
def test_search_bitexact_on_given_knots(cuda_device, t):
    from manifold_flow.b200 import ops

    knots, v, bins = (_t(S[f"search_T{t}/{k}"]) for k in ("knots", "values", "bins"))
    kk = knots[:, None, :].expand(-1, v.shape[1], -1).reshape(-1, knots.shape[1]).contiguous()
    got = ops.rqs_search(kk.to(cuda_device), v.reshape(-1).to(cuda_device)).cpu().view(bins.shape)
    assert torch.equal(got.long(), bins)


@pytest.mark.parametrize("name", [str(c) for c in S["cases"]])
@pytest.mark.parametrize("inverse", [False, True])
def test_spline_function_matches_golden(cuda_device, name, inverse):
    from manifold_flow.transforms.splines import unconstrained_rational_quadratic_spline as spline

    t = float(name.split("_")[1][1:])
    tag = "inv" if inverse else "fwd"
    x, w, h, d = (_t(S[f"{name}/{k}"], cuda_device) for k in "xwhd")
    y, lad = spline(x, w, h, d, inverse=inverse, tails="linear", tail_bound=t)
    y, lad = y.cpu(), lad.cpu()
    xc = x.cpu()
    gu.assert_as_accurate_as_reference(y, _t(S[f"{name}/{tag}_y"]), _t(S[f"{name}/{tag}_y64"]), what="outputs")
    gu.assert_as_accurate_as_reference(lad, _t(S[f"{name}/{tag}_lad"]), _t(S[f"{name}/{tag}_lad64"]), what="logabsdet", floor=5e-6)
    if "saturated" not in name:  # well-conditioned cases also meet the plain element-wise bound
        torch.testing.assert_close(y, _t(S[f"{name}/{tag}_y"]), rtol=RTOL, atol=ATOL * max(1.0, t / 3))
        torch.testing.assert_close(lad, _t(S[f"{name}/{tag}_lad"]), rtol=1e-4, atol=2e-4)
    out = (xc < -t) | (xc > t)
    assert out.any()
    assert torch.equal(y[out], xc[out]) and torch.all(lad[out] == 0)


@pytest.mark.parametrize("k,t", [(8, 3.0), (10, 6.0), (11, 10.0), (4, 1.0), (16, 5.0)])
@pytest.mark.parametrize("inverse", [False, True])
def test_spline_rows_layout_vs_oracle_large(cuda_device, k, t, inverse):
    from manifold_flow.b200 import ops

    g = torch.Generator().manual_seed(5 + k)
    n = 40013  # ragged: not a multiple of any tile size
    x = torch.randn(n, generator=g) * (t / 2)
    params = torch.randn(n, 3 * k - 1, generator=g) * 1.5
    split = lambda p: (p[:, :k], p[:, k : 2 * k], p[:, 2 * k :])
    ref_y, ref_l, ref_b = oracle.rqs(x.double(), *split(params.double()), inverse, t, return_bins=True)
    r32_y, r32_l, r32_b = oracle.rqs(x, *split(params), inverse, t, return_bins=True)
    y, lad, bins = ops.rqs_elementwise(x.to(cuda_device), params.to(cuda_device), k, t, inverse=inverse, return_bins=True)
    gu.assert_as_accurate_as_reference(y, r32_y, ref_y, what="outputs")
    gu.assert_as_accurate_as_reference(lad, r32_l, ref_l, what="logabsdet", floor=5e-6)
    # end-to-end bins against the reference's own fp32 computation: a flip is only legal within a
    # couple of ulps of a knot (~1 per 2e6 elements expected)
    mism = (bins.cpu().long() != r32_b).sum().item()
    assert mism <= max(2, n // 20000), f"{mism} bin mismatches out of {n}"


def _plane_case(k, t, b, c, hh, ww, seed):
    g = torch.Generator().manual_seed(seed)
    ct = (c + 1) // 2
    x = torch.randn(b, c, hh, ww, generator=g) * (t / 2)
    params = torch.randn(b, ct * (3 * k - 1), hh, ww, generator=g) * 8.0
    return x, params, ct


def _oracle_planes(x, params, ct, k, t, hidden, inverse):
    b, c, hh, ww = x.shape
    p = params.reshape(b, ct, -1, hh, ww).permute(0, 1, 3, 4, 2)
    scale = np.sqrt(hidden)
    y_t, lad = oracle.rqs(x[:, :ct], p[..., :k] / scale, p[..., k : 2 * k] / scale, p[..., 2 * k :], inverse, t)
    out = x.clone()
    out[:, :ct] = y_t
    return out, oracle.sum_except_batch(lad)


@pytest.mark.parametrize("shape", [(3, 12, 32, 32), (2, 24, 16, 16), (5, 48, 8, 8), (7, 96, 4, 4), (2, 6, 5, 3), (1, 2, 1, 1)])
@pytest.mark.parametrize("inverse", [False, True])
def test_coupling_planes_layout_vs_oracle(cuda_device, shape, inverse):
    from manifold_flow.b200 import ops

    k, t, hidden = 11, 10.0, 100
    b, c, hh, ww = shape
    x, params, ct = _plane_case(k, t, b, c, hh, ww, seed=b * 100 + c)
    geom = ops.SplineGeometry(ct, 0, 1, c - ct, ct, 1)
    y, lad = ops.rqs_coupling(x.to(cuda_device), params.to(cuda_device), geom, k, t, 1e-3, 1e-3, 1e-3, np.sqrt(hidden), inverse)
    ref_y, ref_l = _oracle_planes(x.double(), params.double(), ct, k, t, hidden, inverse)
    r32_y, r32_l = _oracle_planes(x, params, ct, k, t, hidden, inverse)
    gu.assert_as_accurate_as_reference(y, r32_y, ref_y, what="outputs")
    gu.assert_as_accurate_as_reference(lad, r32_l, ref_l, what="per-sample logabsdet", floor=(1e-4 if inverse else 3e-5) * max(1.0, ct * hh * ww / 100))
    # identity channels are copied bit-exactly
    assert torch.equal(y.cpu()[:, ct:], x[:, ct:])


def test_per_sample_logdet_is_deterministic(cuda_device):
    from manifold_flow.b200 import ops

    k, t = 11, 10.0
    x, params, ct = _plane_case(k, t, 25, 12, 32, 32, seed=3)
    geom = ops.SplineGeometry(ct, 0, 1, 12 - ct, ct, 1)
    xs, ps = x.to(cuda_device), params.to(cuda_device)
    first = ops.rqs_coupling(xs, ps, geom, k, t, 1e-3, 1e-3, 1e-3, 10.0, False)[1]
    for _ in range(5):
        again = ops.rqs_coupling(xs, ps, geom, k, t, 1e-3, 1e-3, 1e-3, 10.0, False)[1]
        assert torch.equal(first, again)


@pytest.mark.parametrize("k,t", [(8, 3.0), (11, 10.0)])
def test_round_trip_and_cancelling_logdets(cuda_device, k, t):
    from manifold_flow.b200 import ops

    g = torch.Generator().manual_seed(k)
    n = 1 << 16
    x = (torch.randn(n, generator=g) * (t / 2)).to(cuda_device)
    params = torch.randn(n, 3 * k - 1, generator=g).to(cuda_device)
    y, lad = ops.rqs_elementwise(x, params, k, t)
    xi, ladi = ops.rqs_elementwise(y, params, k, t, inverse=True)
    # the fp32 reference round-trips to ~1e-5; a few ill-conditioned elements
    # out of 65536 go to ~1e-3 in the reference's own arithmetic, hence quantile + loose max
    dx, dl = (xi - x).abs().cpu(), (lad + ladi).abs().cpu()
    # the same round trip through the reference algorithm in fp32 on the CPU sets the yardstick
    xc, pc = x.cpu(), params.cpu()
    split = lambda p: (p[:, :k], p[:, k : 2 * k], p[:, 2 * k :])
    yr, lr = oracle.rqs(xc, *split(pc), False, t)
    xr, lir = oracle.rqs(yr, *split(pc), True, t)
    rx, rl = (xr - xc).abs(), (lr + lir).abs()
    q = lambda v, f: float(v.kthvalue(int(f * n))[0])
    for f in (0.5, 0.99, 0.999):
        assert q(dx, f) <= 2.0 * q(rx, f) + 1e-6, (f, q(dx, f), q(rx, f))
        assert q(dl, f) <= 2.0 * q(rl, f) + 5e-6, (f, q(dl, f), q(rl, f))
    assert q(dx, 0.99) <= 2e-5 * max(1.0, t / 3)  # the bulk round-trips to the north star's 1e-5 scale
    assert float(dx.max()) <= 20 * float(rx.max()) + 1e-4 and float(dl.max()) <= 20 * float(rl.max()) + 1e-3


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("layout", ["rows", "planes"])
def test_backward_matches_oracle_autograd(cuda_device, inverse, layout):
    from manifold_flow.b200 import ops

    k, t, hidden = 11, 10.0, 100.0
    g = torch.Generator().manual_seed(17)
    if layout == "rows":
        b, d = 301, 14
        ct = 7
        x = torch.randn(b, d, generator=g) * (t / 2)
        params = torch.randn(b, ct * (3 * k - 1), generator=g) * 6
        geom = ops.SplineGeometry(ct, 0, 2, d - ct, 1, 2)  # even features transformed
        tr, idn = slice(0, None, 2), slice(1, None, 2)
    else:
        b, c, hh, ww = 3, 12, 8, 8
        ct = 6
        x = torch.randn(b, c, hh, ww, generator=g) * (t / 2)
        params = torch.randn(b, ct * (3 * k - 1), hh, ww, generator=g) * 6
        geom = ops.SplineGeometry(ct, 0, 1, c - ct, ct, 1)
        tr, idn = slice(0, ct), slice(ct, None)
    gy = torch.randn(x.shape, generator=g)
    gl = torch.randn(b, generator=g)

    xd = x.to(cuda_device).requires_grad_(True)
    pd = params.to(cuda_device).requires_grad_(True)
    y, lad = ops.rqs_coupling(xd, pd, geom, k, t, 1e-3, 1e-3, 1e-3, np.sqrt(hidden), inverse)
    ((y * gy.to(cuda_device)).sum() + (lad * gl.to(cuda_device)).sum()).backward()

    xo = x.double().requires_grad_(True)
    po = params.double().requires_grad_(True)
    if layout == "rows":
        pv = po.reshape(b, ct, -1)
    else:
        pv = po.reshape(b, ct, -1, hh, ww).permute(0, 1, 3, 4, 2)
    scale = np.sqrt(hidden)
    y_t, lad_o = oracle.rqs(xo[:, tr], pv[..., :k] / scale, pv[..., k : 2 * k] / scale, pv[..., 2 * k :], inverse, t)
    loss = (y_t * gy.double()[:, tr]).sum() + (xo[:, idn] * gy.double()[:, idn]).sum() + (oracle.sum_except_batch(lad_o) * gl.double()).sum()
    loss.backward()
    # the same graph in float32 on the CPU = what the reference's autograd computes
    x3 = x.clone().requires_grad_(True)
    p3 = params.clone().requires_grad_(True)
    pv3 = p3.reshape(b, ct, -1) if layout == "rows" else p3.reshape(b, ct, -1, hh, ww).permute(0, 1, 3, 4, 2)
    y3, l3 = oracle.rqs(x3[:, tr], pv3[..., :k] / scale, pv3[..., k : 2 * k] / scale, pv3[..., 2 * k :], inverse, t)
    ((y3 * gy[:, tr]).sum() + (x3[:, idn] * gy[:, idn]).sum() + (oracle.sum_except_batch(l3) * gl).sum()).backward()
    gu.assert_as_accurate_as_reference(xd.grad, x3.grad, xo.grad, what="grad inputs", floor=1e-5, rel=1e-4)
    gu.assert_as_accurate_as_reference(pd.grad, p3.grad, po.grad, what="grad params", floor=1e-5, rel=1e-4)


def test_elementwise_backward_matches_oracle(cuda_device):
    from manifold_flow.transforms.splines import unconstrained_rational_quadratic_spline as spline

    k, t = 8, 3.0
    g = torch.Generator().manual_seed(23)
    n = 777
    x = torch.randn(n, generator=g) * 1.5
    w, h, d = torch.randn(n, k, generator=g), torch.randn(n, k, generator=g), torch.randn(n, k - 1, generator=g)
    gy, gl = torch.randn(n, generator=g), torch.randn(n, generator=g)
    for inverse in (False, True):
        leaves = [v.to(cuda_device).requires_grad_(True) for v in (x, w, h, d)]
        y, lad = spline(*leaves, inverse=inverse, tails="linear", tail_bound=t)
        ((y * gy.to(cuda_device)).sum() + (lad * gl.to(cuda_device)).sum()).backward()
        ref = [v.double().requires_grad_(True) for v in (x, w, h, d)]
        yo, lo = oracle.rqs(*ref, inverse, t)
        ((yo * gy.double()).sum() + (lo * gl.double()).sum()).backward()
        r32 = [v.clone().requires_grad_(True) for v in (x, w, h, d)]
        y3, l3 = oracle.rqs(*r32, inverse, t)
        ((y3 * gy).sum() + (l3 * gl).sum()).backward()
        for a, b32, b64 in zip(leaves, r32, ref):
            gu.assert_as_accurate_as_reference(a.grad, b32.grad, b64.grad, what="grad", floor=1e-5, rel=1e-4)


def test_edge_cases(cuda_device):
    from manifold_flow.b200 import ops

    k, t = 8, 3.0
    # empty batch
    y, lad = ops.rqs_elementwise(torch.empty(0, device=cuda_device), torch.empty(0, 3 * k - 1, device=cuda_device), k, t)
    assert y.numel() == 0 and lad.numel() == 0
    # everything in the tails, NaN passes through like the reference's mask logic
    x = torch.tensor([-5.0, 4.0, float("inf"), float("nan")], device=cuda_device)
    y, lad = ops.rqs_elementwise(x, torch.zeros(4, 3 * k - 1, device=cuda_device), k, t)
    assert torch.equal(y[:3], x[:3]) and torch.isnan(y[3]) and torch.all(lad == 0)
    # uniform bins with unit knot derivatives are the identity map inside the box too
    x = torch.linspace(-t, t, 1001, device=cuda_device)
    ident = torch.zeros(1001, 3 * k - 1, device=cuda_device)
    ident[:, 2 * k :] = float(np.log(np.exp(1 - 1e-3) - 1))
    y, lad = ops.rqs_elementwise(x, ident, k, t)
    assert (y - x).abs().max().item() < 3e-6 and lad.abs().max().item() < 3e-6
    # unsupported bin count fails loudly (compiled in: 2..16, 20, 24, 32)
    with pytest.raises(ValueError):
        ops.rqs_elementwise(x, torch.zeros(1001, 3 * 17 - 1, device=cuda_device), 17, t)
    # too many bins for the minimal width (reference: ValueError, rational_quadratic.py:97-100)
    with pytest.raises(ValueError):
        ops.rqs_elementwise(x, torch.zeros(1001, 3 * 16 - 1, device=cuda_device), 16, t, min_w=0.1)


def test_large_batch_rows_and_planes(cuda_device):
    """Sizes of the spline micro-benchmark: 2^20 vector samples; 256 images."""
    from manifold_flow.b200 import ops

    k, t = 11, 10.0
    g = torch.Generator(device=cuda_device).manual_seed(1)
    n = 1 << 20
    x = torch.randn(n, 2, device=cuda_device, generator=g) * 5
    params = torch.randn(n, 32, device=cuda_device, generator=g)
    geom = ops.SplineGeometry(1, 0, 2, 1, 1, 2)
    y, lad = ops.rqs_coupling(x, params, geom, k, t, 1e-3, 1e-3, 1e-3, 10.0, False)
    xi, ladi = ops.rqs_coupling(y, params, geom, k, t, 1e-3, 1e-3, 1e-3, 10.0, True)
    kq = int(0.9999 * n)
    assert float((xi - x).abs().max()) < 5e-2 and float((xi - x).abs().reshape(-1).kthvalue(2 * kq)[0]) < 1e-4
    assert float((lad + ladi).abs().kthvalue(kq)[0]) < 1e-3
    assert torch.equal(y[:, 1], x[:, 1])
    # sortedness: the spline is monotone, so it preserves the order of inputs that share parameters
    xs = torch.sort(torch.randn(4096, device=cuda_device) * 5)[0]
    ys, _ = ops.rqs_elementwise(xs, params[:1].expand(4096, -1).contiguous(), k, t)
    assert torch.all(ys[1:] >= ys[:-1])
    b = 256
    xi4 = torch.randn(b, 12, 32, 32, device=cuda_device, generator=g) * 5
    p4 = torch.randn(b, 6 * 32, 32, 32, device=cuda_device, generator=g)
    geom4 = ops.SplineGeometry(6, 0, 1, 6, 6, 1)
    y4, lad4 = ops.rqs_coupling(xi4, p4, geom4, k, t, 1e-3, 1e-3, 1e-3, 10.0, False)
    x4, ladi4 = ops.rqs_coupling(y4, p4, geom4, k, t, 1e-3, 1e-3, 1e-3, 10.0, True)
    assert float((x4 - xi4).abs().reshape(-1).kthvalue(int(0.9999 * x4.numel()))[0]) < 1e-4
    assert (lad4 + ladi4).abs().max().item() < 5e-2  # 6144 elements per sample


# ---- rational_quadratic_spline on an arbitrary box (rational_quadratic.py:67-168) ----
SB = gu.load_npz("spline_box.npz")


@pytest.mark.parametrize("name", [str(c) for c in SB["cases"]])
def test_box_spline_matches_golden(cuda_device, name):
    from manifold_flow.transforms.splines import rational_quadratic_spline as spline

    left, right, bottom, top = (float(v) for v in SB[f"{name}/box"])
    box = dict(left=left, right=right, bottom=bottom, top=top)
    x, w, h, d = (_t(SB[f"{name}/{k}"], cuda_device).requires_grad_(True) for k in "xwhd")
    y, lad = spline(x, w, h, d, inverse=False, **box)
    scale = max(1.0, abs(bottom), abs(top))
    gu.assert_as_accurate_as_reference(y, _t(SB[f"{name}/fwd_y"]), _t(SB[f"{name}/fwd_y64"]), what="outputs", floor=1e-6 * scale)
    gu.assert_as_accurate_as_reference(lad, _t(SB[f"{name}/fwd_lad"]), _t(SB[f"{name}/fwd_lad64"]), what="logabsdet", floor=5e-6)
    torch.testing.assert_close(y.detach().cpu(), _t(SB[f"{name}/fwd_y"]), rtol=RTOL, atol=ATOL * scale)
    (y.sum() + lad.sum()).backward()
    for got, key in ((x.grad, "g_x"), (w.grad, "g_w"), (h.grad, "g_h"), (d.grad, "g_d")):
        gu.assert_grad_as_accurate_as_reference(got, _t(SB[f"{name}/{key}"]), _t(SB[f"{name}/{key}64"]), what=key)
    # inverse: golden where the reference can run it, and the round trip everywhere
    if f"{name}/inv_in" in SB:
        xi, ladi = spline(_t(SB[f"{name}/inv_in"], cuda_device), w.detach(), h.detach(), d.detach(), inverse=True, **box)
        gu.assert_as_accurate_as_reference(xi, _t(SB[f"{name}/inv_x"]), _t(SB[f"{name}/inv_x64"]), what="inverse outputs",
                                           floor=1e-6 * max(1.0, abs(left), abs(right)))
        gu.assert_as_accurate_as_reference(ladi, _t(SB[f"{name}/inv_lad"]), _t(SB[f"{name}/inv_lad64"]), what="inverse logabsdet", floor=2e-5)
    from manifold_flow.b200 import ops

    params = torch.cat((w, h, d), dim=-1).detach()
    y2 = y.detach().clamp(bottom, top)
    xr, ladr = ops.rqs_box(y2, params, w.shape[-1], left, right, bottom, top, inverse=True)
    # round trip, weighted by the local slope dy/dx = exp(lad): where the spline is nearly flat one ulp of y is many ulps of x
    back = (xr - x.detach()).abs() * torch.exp(lad.detach())
    assert float(back.max()) <= 2e-5 * scale + 1e-5 * max(abs(left), abs(right)), float(back.max())
    assert float((ladr + lad.detach()).abs().median()) < 1e-4


def test_box_spline_domain_and_argument_errors(cuda_device):
    from manifold_flow.transforms import InputOutsideDomain
    from manifold_flow.transforms.splines import rational_quadratic_spline as spline

    k = 6
    w, h, d = torch.zeros(4, k, device=cuda_device), torch.zeros(4, k, device=cuda_device), torch.zeros(4, k + 1, device=cuda_device)
    x = torch.tensor([0.1, 0.5, 0.9, 1.5], device=cuda_device)
    with pytest.raises(InputOutsideDomain):   # rational_quadratic.py:85-86
        spline(x, w, h, d)
    with pytest.raises(InputOutsideDomain):
        spline(-x, w, h, d)
    with pytest.raises(ValueError):           # :90-93
        spline(x.clamp(0, 1), w, h, d, min_bin_width=0.5)
    with pytest.raises(ValueError):
        spline(x.clamp(0, 1), w, h, d[:, :-1])
    # identity-like default (zero logits): monotone, maps the box ends onto the box ends exactly
    y, lad = spline(torch.tensor([2.0, 3.0, 5.0], device=cuda_device), w[:3], h[:3], d[:3], left=2.0, right=5.0, bottom=-1.0, top=1.0)
    assert float(y[0]) == -1.0 and float(y[2]) == 1.0 and float(y[1]) > -1.0
    e = torch.empty(0, device=cuda_device)
    y, lad = spline(e, torch.empty(0, k, device=cuda_device), torch.empty(0, k, device=cuda_device), torch.empty(0, k + 1, device=cuda_device))
    assert y.shape == (0,) and lad.shape == (0,)


@pytest.mark.parametrize("k", [2, 3, 7, 9, 13, 15, 20, 24, 32])
def test_every_bin_count_is_compiled_in(cuda_device, k):
    """--splinebins is a free integer in the reference (experiments/train.py:60)."""
    from manifold_flow.b200 import ops

    g = torch.Generator().manual_seed(k)
    n, t = 3001, 4.0
    x = torch.randn(n, generator=g) * (t / 2)
    params = torch.randn(n, 3 * k - 1, generator=g)
    split = lambda p: (p[:, :k], p[:, k : 2 * k], p[:, 2 * k :])
    for inverse in (False, True):
        ref_y, ref_l = oracle.rqs(x.double(), *split(params.double()), inverse, t)
        r32_y, r32_l = oracle.rqs(x, *split(params), inverse, t)
        y, lad = ops.rqs_elementwise(x.to(cuda_device), params.to(cuda_device), k, t, inverse=inverse)
        gu.assert_as_accurate_as_reference(y, r32_y, ref_y, what=f"K={k} outputs")
        gu.assert_as_accurate_as_reference(lad, r32_l, ref_l, what=f"K={k} logabsdet", floor=5e-6)


# ---- the exact-operation-order build (libmflow_b200_exact.so, -DMFLOW_RQS_EXACT_ORDER) ----
@pytest.mark.parametrize("name", [str(c) for c in S["cases"]])
@pytest.mark.parametrize("inverse", [False, True])
def test_exact_order_build_meets_1e5_wherever_the_reference_does(cuda_device, name, inverse):
    """the default build contracts FMAs and uses ex2.approx; this build evaluates every fp32
    operation of the reference separately, in the reference's order.  Wherever the reference's own fp32 result is
    within 1e-5 relative of its fp64 result (well-conditioned elements), BOTH builds must be within 1e-5 relative of the
    reference's fp32 output - a plain element-wise bound, no statistics."""
    from manifold_flow.b200 import cabi
    from manifold_flow.transforms.splines import unconstrained_rational_quadratic_spline as spline

    t = float(name.split("_")[1][1:])
    tag = "inv" if inverse else "fwd"
    x, w, h, d = (_t(S[f"{name}/{k}"], cuda_device) for k in "xwhd")
    results = {"default": spline(x, w, h, d, inverse=inverse, tails="linear", tail_bound=t)}
    with cabi.library_variant("exact"):
        results["exact"] = spline(x, w, h, d, inverse=inverse, tails="linear", tail_bound=t)
    # outputs: plain 1e-5; logabsdet = log(dnum) - 2 log(den) inherits 1-2 ulp differences between the CPU's and
    # libdevice's logf / log1pf / expf at |log| ~ 10..30, which no operation order removes: 2e-5 there
    # (profiles/r2_spline_exact_order_stats.txt: 0 of 1911 outputs and 6 / 32 of 1600 logabsdets beyond 1e-5, max 1.8e-5)
    for key, bar in (("y", 1e-5), ("lad", 2e-5)):
        ref32, ref64 = _t(S[f"{name}/{tag}_{key}"]).double(), _t(S[f"{name}/{tag}_{key}64"])
        scale = ref64.abs().clamp_min(1.0)
        well = (ref32 - ref64).abs() <= 0.2e-5 * scale       # the reference itself is 5x inside the bar here
        assert well.float().mean() > 0.25
        for build, (y, lad) in results.items():
            got = (y if key == "y" else lad).cpu().double()
            err = (got - ref32).abs()
            assert torch.all(err[well] <= bar * scale[well]), \
                f"{build} build, {key}: max rel err {float((err / scale)[well].max()):.2e} on well-conditioned elements"
    # the exact build reproduces a large share of the reference's fp32 outputs bit for bit
    same = (results["exact"][0].cpu() == _t(S[f"{name}/{tag}_y"])).float().mean().item()
    assert same > 0.25, same


@pytest.mark.parametrize("rows,d,ct,k,w", [(1024, 14, 7, 11, 32), (512, 3, 1, 8, 16), (2048 + 256, 40, 20, 11, 16), (192, 2, 1, 10, 8),
                                           (64, 5, 2, 4, 4), (4096, 64, 32, 11, 32)])
@pytest.mark.parametrize("inverse", [False, True])
def test_rows_x_planes_layout_equals_rows_layout(cuda_device, rows, d, ct, k, w, inverse):
    """Vectors whose conditioner ran on the convolution kernels hand the spline their parameters as [rows / hw, C_t*P, h, w]
    (mflow_rqs_desc.p_hw).  Per element the arithmetic is the rows kernel's, so outputs, input gradients and parameter
    gradients are bit-identical; a row's log-det is summed over two channel groups (order differs: 1e-6)."""
    from manifold_flow.b200 import ops

    t, hidden, hw = 5.0, 100.0, w * w
    p = 3 * k - 1
    g = torch.Generator().manual_seed(rows + d + k)
    x = (torch.randn(rows, d, generator=g) * (t / 2)).to(cuda_device)
    params = (torch.randn(rows, ct * p, generator=g) * 6).to(cuda_device)
    planes = params.view(rows // hw, hw, ct * p).transpose(1, 2).contiguous().view(rows // hw, ct * p, w, w)
    step = 2 if 2 * ct <= d + 1 and ct > 1 else 1
    n_id = min(d - ct, ct) if step == 2 else d - ct
    geom = ops.SplineGeometry(ct, 0, step, n_id, 1 if step == 2 else ct, step)
    gy = torch.randn(rows, d, generator=g).to(cuda_device)
    gl = torch.randn(rows, generator=g).to(cuda_device)
    outs = []
    for par in (params, planes):
        xd, pd = x.clone().requires_grad_(True), par.clone().requires_grad_(True)
        y, lad = ops.rqs_coupling(xd, pd, geom, k, t, 1e-3, 1e-3, 1e-3, np.sqrt(hidden), inverse)
        ((y * gy).sum() + (lad * gl).sum()).backward()
        outs.append((y.detach(), lad.detach(), xd.grad, pd.grad))
    (y0, l0, gx0, gp0), (y1, l1, gx1, gp1) = outs
    touched = torch.zeros(d, dtype=torch.bool)
    touched[[geom.t_offset + c * geom.t_step for c in range(ct)]] = True
    touched[[geom.i_offset + c * geom.i_step for c in range(n_id)]] = True
    cols = touched.to(cuda_device)
    assert torch.equal(y0[:, cols], y1[:, cols])
    torch.testing.assert_close(l1, l0, rtol=1e-6, atol=1e-5)
    assert torch.equal(gx0[:, cols], gx1[:, cols])
    gp1_rows = gp1.view(rows // hw, ct * p, hw).transpose(1, 2).reshape(rows, ct * p)
    assert torch.equal(gp0, gp1_rows)


def test_vector_model_planes_path_matches_fused_path(cuda_device):
    """cfg3-shaped M-flow at 32768 rows: conditioners on the convolution kernels with the spline reading their NCHW output
    (rows x planes) against the same model through the fused per-coupling MLP kernel (the path of the small-batch parity
    tests): log_prob and parameter gradients agree to fp32 accuracy."""
    import golden_util as gu
    from manifold_flow.b200 import config, factory

    torch.manual_seed(gu.SEED)
    model = factory.create_model("cfg3").to(cuda_device)
    model.train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(32768, 40, generator=g).to(cuda_device)
    c = torch.randn(32768, 2, generator=g).to(cuda_device)
    with torch.no_grad():
        model(x, mode="projection", context=c)

    def run(min_rows):
        prev, config.linear_tc_min_rows = config.linear_tc_min_rows, min_rows
        try:
            model.zero_grad(set_to_none=True)
            lp = model(x, mode="mf-fixed-manifold", context=c)[1]
            (-lp.mean()).backward()
            return lp.detach(), [p.grad.clone() for p in model.parameters() if p.grad is not None]
        finally:
            config.linear_tc_min_rows = prev

    lp_tc, g_tc = run(32768)
    lp_mlp, g_mlp = run(1 << 40)
    torch.testing.assert_close(lp_tc, lp_mlp, rtol=2e-5, atol=2e-3)
    num = sum(float((a.double() - b.double()).pow(2).sum()) for a, b in zip(g_tc, g_mlp)) ** 0.5
    den = sum(float(b.double().pow(2).sum()) for b in g_mlp) ** 0.5
    assert num / den < 1e-4, num / den
