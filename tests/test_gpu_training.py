"""GPU: the synchronisation-free training step (manifold_flow.b200.training.GraphedTrainStep) takes the
same optimisation trajectory as the reference trainer's batch_train sequence (experiments/training/trainer.py:177-184,
512-521: zero_grad, backward, clip_grad_norm_, optimizer.step) executed eagerly."""
import copy

import pytest
import torch

import golden_util as gu
import model_util as mu

pytestmark = pytest.mark.gpu


def nll(x_pred, x_true, log_p, hidden=None):   # experiments/training/losses.py:10-13
    return -torch.mean(log_p)


def mse(x_pred, x_true, log_p, hidden=None):   # :16-19
    return torch.nn.functional.mse_loss(x_pred, x_true)


@pytest.mark.parametrize("name,batch,mode,losses,weights", [
    ("cfg3", 100, "mf-fixed-manifold", (nll,), (1.0,)),
    ("cfg1", 100, "projection", (mse,), (100.0,)),
    ("cfg4s", 16, "mf-fixed-manifold", (nll, mse), (1.0, 0.1)),
])
def test_graphed_step_follows_the_eager_trajectory(cuda_device, name, batch, mode, losses, weights):
    from manifold_flow.b200.training import GraphedTrainStep

    ref_model, _, _ = mu.seeded_product_model(name)
    ref_model.to(cuda_device).train()
    model = copy.deepcopy(ref_model)
    batches = [gu.config_inputs(name, batch=batch, seed=100 + i) for i in range(7)]
    to_dev = lambda t: None if t is None else t.to(cuda_device)
    trained = lambda m: list(m.inner_transform.parameters()) if mode == "mf-fixed-manifold" else list(m.parameters())

    # reference-style eager loop
    opt_ref = torch.optim.AdamW(trained(ref_model), lr=torch.tensor(3e-4), capturable=True)
    ref_losses = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for x, c in batches:
            xd, cd = to_dev(x), to_dev(c)
            kwargs = {"mode": mode} if cd is None else {"mode": mode, "context": cd}
            x_reco, log_prob, u = ref_model(xd, **kwargs)
            terms = [f(x_reco, xd, log_prob) for f in losses]
            loss = sum(w * t for w, t in zip(weights, terms))
            opt_ref.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(trained(ref_model), 1.0)
            opt_ref.step()
            ref_losses.append(float(loss))
    torch.cuda.current_stream().wait_stream(side)

    # graphed: 3 eager warm-up steps, capture at the 4th call, replays afterwards
    opt = torch.optim.AdamW(trained(model), lr=torch.tensor(3e-4), capturable=True)
    step = GraphedTrainStep(model, opt, losses, weights, {"mode": mode}, clip_gradient=1.0, parameters=trained(model), warmup=3)
    got_losses = []
    for x, c in batches:
        out = step(to_dev(x), to_dev(c))
        got_losses.append(step.loss_values(out)[0])
    assert len(step._graphs) == 1 and step.steps_done == len(batches)
    assert step.check_nans() == dict(x_reco=0, log_prob=0, loss=0)
    torch.testing.assert_close(torch.tensor(got_losses), torch.tensor(ref_losses), rtol=1e-5, atol=1e-5)
    for (k, p), (_, q) in zip(model.named_parameters(), ref_model.named_parameters()):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")


def test_graphed_step_needs_a_capturable_optimizer(cuda_device):
    from manifold_flow.b200.training import GraphedTrainStep

    model, _, _ = mu.seeded_product_model("cfg1")
    model.to(cuda_device)
    with pytest.raises(ValueError, match="capturable"):
        GraphedTrainStep(model, torch.optim.AdamW(model.parameters(), lr=1e-3), (nll,))
