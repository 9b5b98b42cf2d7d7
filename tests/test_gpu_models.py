"""GPU parity of whole models (cfg1..cfg5 of BASELINE.json) through the drop-in API: the reference's
seeded weights (+ a shared deterministic perturbation) are loaded into the product model and every
supported mode is compared with the outputs frozen from the reference and with the fp64 truth.

Tolerances: u / x_reco 1e-5 relative (+ small absolute floor scaled to the data range); log_prob
1e-4 absolute per sample for the vector configs (north star) - for the 12288-dimensional image
configs |log_prob| ~ 7e4 where one fp32 ulp is 8e-3, so the bound there is 2e-6 relative."""
import numpy as np
import pytest
import torch

import golden_util as gu
import mflow_oracle as oracle
import model_util as mu

pytestmark = pytest.mark.gpu

META = gu.load_json("models.json")
GOLD = gu.load_npz("models.npz")
VECTOR = ["cfg1", "cfg2", "cfg3"]
IMAGE_MF = ["cfg4s", "cfg4"]
IMAGE_FLOW = ["cfg5s", "cfg5"]


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _model(name, device):
    model, sd0, sd = mu.seeded_product_model(name)
    return model.to(device), sd


def _lp_tol(name):
    return dict(rtol=0, atol=1e-4) if name in VECTOR else dict(rtol=2e-6, atol=5e-3)


@pytest.mark.parametrize("name", VECTOR + IMAGE_MF)
def test_manifold_flow_modes_match_reference(cuda_device, name):
    model, _ = _model(name, cuda_device)
    x, ctx = gu.config_inputs(name)
    xd, cd = x.to(cuda_device), (None if ctx is None else ctx.to(cuda_device))
    img = name not in VECTOR
    with torch.no_grad():
        for mode in ("projection", "pie", "mf-fixed-manifold", "slice"):
            x_reco, log_prob, u = model(xd, mode=mode, context=cd)
            torch.testing.assert_close(u.cpu(), _t(GOLD[f"{name}/u"]), rtol=1e-5, atol=2e-5)
            torch.testing.assert_close(x_reco.cpu(), _t(GOLD[f"{name}/x_reco"]), rtol=1e-5, atol=2e-3 if img else 2e-5)
            if mode == "projection":
                assert log_prob is None
            else:
                torch.testing.assert_close(log_prob.cpu(), _t(GOLD[f"{name}/{mode}/log_prob"]), **_lp_tol(name))
        # fp64 truth
        _, lp, u = model(xd, mode="mf-fixed-manifold", context=cd)
        torch.testing.assert_close(lp.cpu().double(), _t(GOLD[f"{name}/log_prob64"]), **_lp_tol(name))
        torch.testing.assert_close(u.cpu().double(), _t(GOLD[f"{name}/u64"]), rtol=1e-5, atol=2e-5)
        # pie == mf-fixed-manifold; encode/decode/log_prob/sample API
        assert torch.equal(model(xd, mode="pie", context=cd)[1], lp)
        torch.testing.assert_close(model.encode(xd, context=cd), u)
        torch.testing.assert_close(model.log_prob(xd, mode="pie", context=cd), lp)
        torch.testing.assert_close(model.decode(u, context=cd), x_reco)
        torch.testing.assert_close(model.sample(u=u, context=cd), x_reco)
        assert model.sample(n=5, context=None if cd is None else cd[:1].expand(5, -1)).shape[0] == 5


@pytest.mark.parametrize("name", VECTOR)
def test_full_jacobian_likelihood_matches_reference(cuda_device, name):
    """mode="mf": chain of per-layer Jacobians -> J^T J -> Cholesky log-det kernel (d)."""
    model, _ = _model(name, cuda_device)
    x, ctx = gu.config_inputs(name)
    xd, cd = x.to(cuda_device), (None if ctx is None else ctx.to(cuda_device))
    x_reco, log_prob, u = model(xd, mode="mf", context=cd)
    torch.testing.assert_close(log_prob.cpu(), _t(GOLD[f"{name}/mf/log_prob"]), rtol=1e-5, atol=2e-4)
    torch.testing.assert_close(x_reco.cpu(), _t(GOLD[f"{name}/x_reco"]), rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("name", IMAGE_FLOW)
def test_ambient_flow_matches_reference(cuda_device, name):
    model, _ = _model(name, cuda_device)
    x, _ = gu.config_inputs(name)
    xd = x.to(cuda_device)
    with torch.no_grad():
        x_reco, log_prob, u = model(xd)
        torch.testing.assert_close(u.cpu(), _t(GOLD[f"{name}/u"]), rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(log_prob.cpu(), _t(GOLD[f"{name}/flow/log_prob"]), **_lp_tol(name))
        torch.testing.assert_close(log_prob.cpu().double(), _t(GOLD[f"{name}/log_prob64"]), **_lp_tol(name))
        torch.testing.assert_close(x_reco.cpu(), _t(GOLD[f"{name}/x_reco"]), rtol=1e-5, atol=2e-3)
        torch.testing.assert_close(model.log_prob(xd), log_prob)
        torch.testing.assert_close(x_reco.cpu(), x, rtol=1e-4, atol=5e-3)  # an ambient flow reconstructs its input


@pytest.mark.parametrize("name", VECTOR + IMAGE_MF + IMAGE_FLOW)
def test_parameter_gradients_match_reference(cuda_device, name):
    model, _ = _model(name, cuda_device)
    x, ctx = gu.config_inputs(name)
    xd, cd = x.to(cuda_device), (None if ctx is None else ctx.to(cuda_device))
    if name in IMAGE_FLOW:
        log_prob = model(xd)[1]
    else:
        log_prob = model(xd, mode="mf-fixed-manifold", context=cd)[1]
    (-log_prob.mean()).backward()
    norms = META[name]["grad_norms"]
    params = dict(model.named_parameters())
    assert set(norms) <= set(params)
    worst = 0.0
    for k, ref in norms.items():
        g = params[k].grad
        got = 0.0 if g is None else float(g.double().norm())
        assert got == pytest.approx(ref, rel=1e-3, abs=1e-5), k
    for key in [k for k in GOLD if k.startswith(f"{name}/grad/")]:
        pname = key.split("/grad/")[1]
        ref = _t(GOLD[key])
        torch.testing.assert_close(params[pname].grad.cpu(), ref, rtol=1e-3, atol=1e-5 * max(1.0, float(ref.abs().max())))


@pytest.mark.parametrize("name", ["cfg1", "cfg4s"])
def test_projection_training_gradients_match_reference(cuda_device, name):
    """Phase-1 loss of the sequential / alternating M-flow training: MSE(x_reco, x) (train.py:273-283)."""
    model, _ = _model(name, cuda_device)
    x, ctx = gu.config_inputs(name)
    xd, cd = x.to(cuda_device), (None if ctx is None else ctx.to(cuda_device))
    x_reco, _, _ = model(xd, mode="projection", context=cd)
    torch.nn.functional.mse_loss(x_reco, xd).backward()
    norms = META[name]["grad_norms_projection"]
    params = dict(model.named_parameters())
    for k, ref in norms.items():
        g = params[k].grad
        got = 0.0 if g is None else float(g.double().norm())
        assert got == pytest.approx(ref, rel=2e-3, abs=1e-5 * max(1.0, ref)), k


def test_actnorm_initialises_on_first_training_batch(cuda_device):
    from manifold_flow import transforms

    model, _ = _model("cfg5s", cuda_device)
    for m in model.modules():
        if isinstance(m, transforms.ActNorm):
            m.initialized = False
    model.train()
    x, _ = gu.config_inputs("cfg5s", batch=16)
    model(x.to(cuda_device))
    assert all(m.initialized for m in model.modules() if isinstance(m, transforms.ActNorm))


def test_state_dict_round_trip_and_reference_keys(cuda_device):
    model, sd = _model("cfg4s", cuda_device)
    assert list(model.state_dict().keys()) == [k for k, _, _ in META["cfg4s"]["keys"]]
    from manifold_flow.b200 import factory

    other = factory.create_model("cfg4s")
    other.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    factory.mark_actnorm_initialized(other)
    other.to(cuda_device).eval()
    x, ctx = gu.config_inputs("cfg4s")
    with torch.no_grad():
        a = model(x.to(cuda_device), mode="pie", context=ctx.to(cuda_device))[1]
        b = other(x.to(cuda_device), mode="pie", context=ctx.to(cuda_device))[1]
    assert torch.equal(a, b)
