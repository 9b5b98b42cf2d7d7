"""Shared by CPU and GPU model-level tests: seeded product models, perturbed weights, oracle specs."""
import torch

import golden_util as gu
import mflow_oracle as oracle

ORACLE_OVERRIDES = {
    "cfg4s": ("cfg4", dict(data_shape=(3, 16, 16), latent_dim=16, levels=2, outer_layers=4, inner_layers=2)),
    "cfg5s": ("cfg5", dict(data_shape=(3, 16, 16), levels=2, outer_layers=2, inner_layers=2)),
}


def oracle_model(name):
    base, over = ORACLE_OVERRIDES.get(name, (name, {}))
    return oracle.config_model(base, **over)


def seeded_product_model(name):
    """Product model with the reference's seeded initial weights (hash-checked elsewhere), then the
    shared deterministic perturbation; ActNorm marked initialised; eval mode."""
    from manifold_flow.b200 import factory

    torch.manual_seed(gu.SEED)
    model = factory.create_model(name)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sd = gu.perturb_state_dict(sd0)
    model.load_state_dict(sd)
    factory.mark_actnorm_initialized(model)
    model.eval()
    return model, sd0, sd
