"""Model builders for the five BASELINE.json configurations.

The reference builds its models in ``experiments/architectures`` (create_model.py:15-62,
vector_transforms.py:139-177, image_transforms.py:116-286); those factories run unchanged against
this package.  ``bench.py``, ``smoke()`` and the tests cannot import the reference on the GPU box,
so this module assembles the same module trees - same classes, same construction order (hence the
same seeded initial weights and ``state_dict`` keys) - from a plain config dict.
"""
import numpy as np
from torch.nn import functional as F

from .. import nn as nn_
from .. import transforms
from ..flows import Flow, ManifoldFlow
from ..utils import various

CONFIGS = {
    # spherical_gaussian, experiments/train.py:41-60 defaults
    "cfg1": dict(kind="vector_mf", data_dim=3, latent_dim=2, outer_layers=5, inner_layers=5, linear="permutation", num_bins=8,
                 tail_bound=3.0, context_features=None, pie_epsilon=0.01, clip_pie=None),
    # power, experiments/configs/train_power_march.config
    "cfg2": dict(kind="vector_mf", data_dim=3, latent_dim=2, outer_layers=5, inner_layers=5, linear="permutation", num_bins=10,
                 tail_bound=6.0, context_features=1, pie_epsilon=0.01, clip_pie=None),
    # lhc40d, experiments/configs/train_mf_lhc_june.config
    "cfg3": dict(kind="vector_mf", data_dim=40, latent_dim=14, outer_layers=20, inner_layers=15, linear="lu", num_bins=11,
                 tail_bound=10.0, context_features=2, pie_epsilon=0.01, clip_pie=None),
    # gan64d, experiments/configs/train_mf_gan64d_april.config
    "cfg4": dict(kind="image_mf", data_dim=(3, 64, 64), latent_dim=64, levels=4, outer_layers=20, inner_layers=8, linear="lu",
                 num_bins=11, tail_bound=10.0, context_features=1, pie_epsilon=0.01, clip_pie=None, actnorm=True,
                 lin_layers=2, lin_channel_factor=2),
    # celeba, experiments/configs/train_flow_celeba_april.config
    "cfg5": dict(kind="image_flow", data_dim=(3, 64, 64), levels=4, outer_layers=20, inner_layers=8, num_bins=11, tail_bound=10.0,
                 context_features=None, actnorm=True),
    # reduced image variants used by quick tests (same code path)
    "cfg4s": dict(kind="image_mf", data_dim=(3, 16, 16), latent_dim=16, levels=2, outer_layers=4, inner_layers=2, linear="lu",
                  num_bins=11, tail_bound=10.0, context_features=1, pie_epsilon=0.01, clip_pie=None, actnorm=True,
                  lin_layers=2, lin_channel_factor=2),
    "cfg5s": dict(kind="image_flow", data_dim=(3, 16, 16), levels=2, outer_layers=2, inner_layers=2, num_bins=11, tail_bound=10.0,
                  context_features=None, actnorm=True),
}


def vector_transform(dim, flow_steps, linear="permutation", num_bins=8, tail_bound=3.0, context_features=None, hidden_features=100,
                     num_blocks=2):
    """[linear glue, RQ coupling] x flow_steps + linear glue (vector_transforms.py:139-177)."""

    def glue():
        perm = transforms.RandomPermutation(features=dim)
        if linear == "permutation":
            return perm
        if linear == "lu":
            return transforms.CompositeTransform([perm, transforms.LULinear(dim, identity_init=True)])
        raise ValueError(linear)

    def conditioner(n_in, n_out):
        return nn_.ResidualNet(n_in, n_out, hidden_features=hidden_features, context_features=context_features, num_blocks=num_blocks,
                               activation=F.relu, dropout_probability=0.0, use_batch_norm=False)

    steps = []
    for i in range(flow_steps):
        lin = glue()
        coupling = transforms.PiecewiseRationalQuadraticCouplingTransform(
            mask=various.create_alternating_binary_mask(dim, even=(i % 2 == 0)), transform_net_create_fn=conditioner,
            num_bins=num_bins, tails="linear", tail_bound=tail_bound, apply_unconditional_transform=False,
        )
        steps.append(transforms.CompositeTransform([lin, coupling]))
    return transforms.CompositeTransform(steps + [glue()])


def image_transform(c, h, w, levels, steps_per_level, num_bins, tail_bound, postprocessing, use_actnorm=True, context_features=None,
                    hidden_channels=100, num_res_blocks=2, postprocessing_layers=2, postprocessing_channel_factor=2):
    """Glow-style multi-scale stack (image_transforms.py:116-218): per level squeeze, then
    steps_per_level x [ActNorm, 1x1 conv, RQ coupling], then a 1x1 conv; post-processing
    'partial_mlp' (:240-258) or 'permutation' (:277-280)."""
    assert h == w
    res, dim = h, c * h * w

    def conditioner(n_in, n_out):
        return nn_.ConvResidualNet(n_in, n_out, hidden_channels=hidden_channels, context_channels=context_features,
                                   num_blocks=num_res_blocks, use_batch_norm=False, dropout_probability=0.0)

    def step(channels):
        coupling = transforms.PiecewiseRationalQuadraticCouplingTransform(
            mask=various.create_mid_split_binary_mask(channels), transform_net_create_fn=conditioner, tails="linear",
            tail_bound=tail_bound, num_bins=num_bins, apply_unconditional_transform=False, min_bin_width=0.001,
            min_bin_height=0.001, min_derivative=0.001,
        )
        layers = [transforms.ActNorm(channels)] if use_actnorm else []
        layers.extend([transforms.OneByOneConvolution(channels), coupling])
        return transforms.CompositeTransform(layers)

    pre = transforms.AffineScalarTransform(scale=(1.0 / 2 ** 8), shift=-0.5)  # "glow" preprocessing, :289-296
    mct = transforms.MultiscaleCompositeTransform(num_transforms=levels)
    for _ in range(levels):
        squeeze = transforms.SqueezeTransform()
        c, h, w = squeeze.get_output_shape(c, h, w)
        level = transforms.CompositeTransform([squeeze] + [step(c) for _ in range(steps_per_level)] + [transforms.OneByOneConvolution(c)])
        new_shape = mct.add_transform(level, (c, h, w))
        if new_shape:
            c, h, w = new_shape
    if postprocessing == "partial_mlp":
        mask = various.create_mlt_channel_mask(dim, channels_per_level=postprocessing_channel_factor * np.array([1, 2, 4, 8], dtype=int),
                                               resolution=res)
        partial_dim = int(mask.sum())
        layers = [transforms.LULinear(partial_dim, identity_init=True)]
        for _ in range(postprocessing_layers - 1):
            layers.append(transforms.LogTanh(cut_point=1))
            layers.append(transforms.LULinear(partial_dim, identity_init=True))
        post = transforms.CompositeTransform(
            [transforms.PartialTransform(mask, transforms.CompositeTransform(layers)), transforms.MaskBasedPermutation(mask)]
        )
    elif postprocessing == "permutation":
        post = transforms.RandomPermutation(dim)
    else:
        raise NotImplementedError(postprocessing)
    return transforms.CompositeTransform([pre, mct, post])


def create_model(cfg):
    """cfg: a key of CONFIGS or a dict of the same form."""
    m = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    kind = m["kind"]
    if kind == "vector_mf":  # create_vector_mf, create_model.py:121-162 (outer is unconditional unless --conditionalouter)
        outer = vector_transform(m["data_dim"], m["outer_layers"], m["linear"], m["num_bins"], m["tail_bound"], None)
        inner = vector_transform(m["latent_dim"], m["inner_layers"], m["linear"], m["num_bins"], m["tail_bound"], m["context_features"])
        return ManifoldFlow(m["data_dim"], m["latent_dim"], outer, inner, apply_context_to_outer=False, pie_epsilon=m["pie_epsilon"],
                            clip_pie=m["clip_pie"])
    if kind == "image_mf":  # create_image_mf, create_model.py:65-118
        c, h, w = m["data_dim"]
        outer = image_transform(c, h, w, m["levels"], m["outer_layers"] // m["levels"], m["num_bins"], m["tail_bound"], "partial_mlp",
                                use_actnorm=m["actnorm"], context_features=None, postprocessing_layers=m["lin_layers"],
                                postprocessing_channel_factor=m["lin_channel_factor"])
        inner = vector_transform(m["latent_dim"], m["inner_layers"], m["linear"], m["num_bins"], m["tail_bound"], m["context_features"])
        return ManifoldFlow(m["data_dim"], m["latent_dim"], outer, inner, apply_context_to_outer=False, pie_epsilon=m["pie_epsilon"],
                            clip_pie=m["clip_pie"])
    if kind == "image_flow":  # create_image_flow, create_model.py:378-408
        c, h, w = m["data_dim"]
        steps = (m["inner_layers"] + m["outer_layers"]) // m["levels"]
        transform = image_transform(c, h, w, m["levels"], steps, m["num_bins"], m["tail_bound"], "permutation", use_actnorm=m["actnorm"],
                                    context_features=m["context_features"])
        return Flow(m["data_dim"], transform)
    raise ValueError(kind)


def mark_actnorm_initialized(model):
    """What experiments/train.py:fix_act_norm_issue does after loading a checkpoint (:520-526)."""
    for module in model.modules():
        if isinstance(module, transforms.ActNorm):
            module.initialized = True
