"""CPU: (1) the product package rebuilds the reference's models - same state_dict keys, shapes and
seeded initial weights (SHA-256 against the live reference); (2) the oracle, fed those weights,
reproduces the reference's outputs for every supported mode of cfg1..cfg5."""
import numpy as np
import pytest
import torch

import golden_util as gu
import mflow_oracle as oracle
import model_util as mu

META = gu.load_json("models.json")
GOLD = gu.load_npz("models.npz")
NAMES = ["cfg1", "cfg2", "cfg3", "cfg4s", "cfg5s", "cfg4", "cfg5"]


@pytest.fixture(scope="module")
def built():
    return {}


def _get(built, name):
    if name not in built:
        built[name] = mu.seeded_product_model(name)
    return built[name]


@pytest.mark.parametrize("name", NAMES)
def test_state_dict_matches_reference(built, name):
    _, sd0, _ = _get(built, name)
    ref = META[name]
    assert [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd0.items()] == ref["keys"]
    assert gu.state_dict_hashes(sd0) == ref["hashes"]
    model = _get(built, name)[0]
    assert sum(p.numel() for p in model.parameters()) == ref["num_parameters"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_models(built, name):
    _, _, sd = _get(built, name)
    model = mu.oracle_model(name)
    x, ctx = gu.config_inputs(name)
    tol = dict(rtol=1e-5, atol=1e-5)
    with torch.no_grad():
        if model["kind"] == "image_flow":
            x_reco, log_prob, u = oracle.flow_forward(model, sd, x, ctx)
            torch.testing.assert_close(u, torch.from_numpy(GOLD[f"{name}/u"]), **tol)
            torch.testing.assert_close(log_prob, torch.from_numpy(GOLD[f"{name}/flow/log_prob"]), rtol=1e-6, atol=1e-2)
            torch.testing.assert_close(x_reco, torch.from_numpy(GOLD[f"{name}/x_reco"]), rtol=1e-4, atol=2e-3)
            return
        for mode in ("projection", "pie", "mf-fixed-manifold", "slice"):
            x_reco, log_prob, u = oracle.mflow_forward(model, sd, x, mode, ctx)
            if mode == "projection":
                torch.testing.assert_close(u, torch.from_numpy(GOLD[f"{name}/u"]), **tol)
                torch.testing.assert_close(x_reco, torch.from_numpy(GOLD[f"{name}/x_reco"]), rtol=1e-4, atol=2e-3)
                assert log_prob is None
            else:
                ref = torch.from_numpy(GOLD[f"{name}/{mode}/log_prob"])
                torch.testing.assert_close(log_prob, ref, rtol=1e-6, atol=1e-3 if len(model["data_shape"]) == 1 else 1e-1)
    if model["kind"] == "mf":
        x_reco, log_prob, u = oracle.mflow_forward(model, sd, x, "mf", ctx)
        torch.testing.assert_close(log_prob, torch.from_numpy(GOLD[f"{name}/mf/log_prob"]), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4s", "cfg5s"])
def test_oracle_gradients_match_reference(built, name):
    _, _, sd = _get(built, name)
    model = mu.oracle_model(name)
    x, ctx = gu.config_inputs(name)
    sdg = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and k.split(".")[-1] not in ("_shift", "_scale") else v) for k, v in sd.items()}
    if model["kind"] == "image_flow":
        _, log_prob, _ = oracle.flow_forward(model, sdg, x, ctx, decode=False)
    else:
        _, log_prob, _ = oracle.mflow_forward(model, sdg, x, "mf-fixed-manifold", ctx)
    (-log_prob.mean()).backward()
    norms = META[name]["grad_norms"]
    checked = 0
    for k, ref in norms.items():
        g = sdg[k].grad
        got = 0.0 if g is None else float(g.double().norm())
        assert got == pytest.approx(ref, rel=2e-4, abs=1e-6), k
        checked += 1
    assert checked > 10
    for key in [k for k in GOLD if k.startswith(f"{name}/grad/")]:
        pname = key.split("/grad/")[1]
        torch.testing.assert_close(sdg[pname].grad, torch.from_numpy(GOLD[key]), rtol=2e-4, atol=1e-5)
