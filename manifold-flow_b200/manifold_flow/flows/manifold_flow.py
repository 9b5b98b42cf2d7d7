"""Manifold-learning flow: outer flow f, projection onto n latents, inner flow h
(reference: manifold_flow/flows/manifold_flow.py:12-165)."""
import logging

import torch

from .. import distributions, transforms
from ..b200 import ops
from ..transforms import ProjectionSplit, fusion
from ..transforms.base import _expand_lad
from ..utils.various import product
from .base import BaseFlow, host_staged

logger = logging.getLogger(__name__)

_MODES = ("mf", "pie", "slice", "projection", "mf-fixed-manifold")


class ManifoldFlow(BaseFlow):
    """M-flow / PIE.  Same constructor, attributes (``outer_transform``, ``inner_transform``,
    ``projection``, the two latent distributions) and call signatures as the reference (:15-94)."""

    def __init__(self, data_dim, latent_dim, outer_transform, inner_transform=None, pie_epsilon=1.0e-2, apply_context_to_outer=True,
                 clip_pie=False):
        super().__init__()
        self.data_dim = data_dim
        self.latent_dim = latent_dim
        self.apply_context_to_outer = apply_context_to_outer
        self.total_data_dim = product(data_dim)
        self.total_latent_dim = product(latent_dim)
        assert self.total_latent_dim < self.total_data_dim
        self.manifold_latent_distribution = distributions.StandardNormal((self.total_latent_dim,))
        self.orthogonal_latent_distribution = distributions.RescaledNormal(
            (self.total_data_dim - self.total_latent_dim,), std=pie_epsilon, clip=None if not clip_pie else clip_pie * pie_epsilon
        )
        self.projection = ProjectionSplit(self.total_data_dim, self.total_latent_dim)
        self.outer_transform = outer_transform
        self.inner_transform = transforms.IdentityTransform() if inner_transform is None else inner_transform
        self._report_model_parameters()

    # ---- public API ------------------------------------------------------------------------
    @host_staged
    def forward(self, x, mode="mf", context=None, return_hidden=False):
        """Encode, decode and evaluate the likelihood; returns (x_reco, log_prob, u[, hidden]).

        Modes (:40-64, :124-157): "projection" (no density), "pie" == "mf-fixed-manifold"
        (density of PIE in the ambient space), "slice", and "mf" (exact manifold density with the
        full Jacobian of the decoder; evaluation only: its log_prob carries no autograd graph).
        "pie-inv" raises TypeError in the reference and is not offered."""
        if mode not in _MODES:
            raise ValueError("mode must be one of {}, got {!r}".format(_MODES, mode))
        fusion.begin_pass(self)
        u, h_manifold, h_orthogonal, log_det_outer, log_det_inner = self._encode(x, context)
        x_reco, inv_log_det_inner, inv_log_det_outer, inv_jacobian_outer, _ = self._decode(u, mode=mode, context=context)
        log_prob = self._log_prob(mode, u, h_orthogonal, log_det_inner, log_det_outer, inv_log_det_inner, inv_log_det_outer, inv_jacobian_outer)
        if return_hidden:
            return x_reco, log_prob, u, torch.cat((h_manifold, h_orthogonal), -1)
        return x_reco, log_prob, u

    @host_staged
    def encode(self, x, context=None):
        fusion.begin_pass(self)
        return self._encode(x, context=context)[0]

    @host_staged
    def decode(self, u, u_orthogonal=None, context=None):
        return self._decode(u, mode="projection", u_orthogonal=u_orthogonal, context=context)[0]

    @host_staged
    def log_prob(self, x, mode="mf", context=None):
        return self.forward(x, mode, context)[1]

    @host_staged
    def sample(self, u=None, n=1, context=None, sample_orthogonal=False):
        if u is None:
            u = self.manifold_latent_distribution.sample(n, context=None)
        u_orthogonal = self.orthogonal_latent_distribution.sample(n, context=None) if sample_orthogonal else None
        return self.decode(u, u_orthogonal=u_orthogonal, context=context)

    # ---- pieces ------------------------------------------------------------------------------
    def _outer_context(self, context):
        return context if self.apply_context_to_outer else None

    def _encode(self, x, context=None):
        h, log_det_outer = self.outer_transform._apply_transform(x, self._outer_context(context), False)
        h_manifold, h_orthogonal = self.projection(h)
        u, log_det_inner = self.inner_transform._apply_transform(h_manifold, context, False)
        return u, h_manifold, h_orthogonal, _expand_lad(log_det_outer, x), _expand_lad(log_det_inner, x)

    def _decode(self, u, mode, u_orthogonal=None, context=None):
        h, inv_log_det_inner = self.inner_transform._apply_transform(u, context, True)
        inv_log_det_inner = _expand_lad(inv_log_det_inner, u)
        h = self.projection.inverse(h, orthogonal_inputs=u_orthogonal) if u_orthogonal is not None else self.projection.inverse(h)
        octx = self._outer_context(context)
        if mode != "mf":
            x, inv_log_det_outer = self.outer_transform._apply_transform(h, octx, True)
            return x, inv_log_det_inner, _expand_lad(inv_log_det_outer, u), None, h
        # chain of per-layer Jacobians dx/dh of the outer decoder (base.py:52-69), each layer by one
        # reverse pass per coordinate over the whole batch instead of B*D autograd calls (various.py:22-45)
        x, jac = self.outer_transform._jacobian(h, octx, True)
        return ops.eval_only(x, u, 'x_reco of mode="mf"'), inv_log_det_inner, None, jac, h

    def _log_prob(self, mode, u, h_orthogonal, log_det_inner, log_det_outer, inv_log_det_inner, inv_log_det_outer, inv_jacobian_outer):
        dist = self.orthogonal_latent_distribution
        if mode in ("pie", "mf-fixed-manifold"):  # :125-128, :149-152 - one fused epilogue launch
            return ops.gauss_logprob(u, h_orthogonal, eps=dist.std, clip=dist._clip, add0=log_det_outer, add1=log_det_inner)
        if mode == "slice":  # :135-138
            return ops.gauss_logprob(u, torch.zeros_like(h_orthogonal), eps=dist.std, clip=dist._clip,
                                     add0=-inv_log_det_outer, add1=-inv_log_det_inner)
        if mode == "mf":  # :140-147
            jac = inv_jacobian_outer[:, :, : self.total_latent_dim]
            half_logdet = 0.5 * ops.jtj_logdet(jac)
            log_prob = ops.gauss_logprob(u.detach(), add0=-half_logdet, add1=-inv_log_det_inner.detach())
            return ops.eval_only(log_prob, u, 'log_prob of mode="mf"')
        return None

    def _report_model_parameters(self):
        super()._report_model_parameters()
        logger.info("  Outer transform: %.1f M parameters", sum(p.numel() for p in self.outer_transform.parameters()) / 1e6)
        logger.info("  Inner transform: %.1f M parameters", sum(p.numel() for p in self.inner_transform.parameters()) / 1e6)
