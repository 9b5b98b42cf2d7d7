"""M-flow with a separate encoder network (reference: manifold_flow/flows/encoder_manifold_flow.py:13-151).
Shares decoder and likelihood code with ``ManifoldFlow``; only the encoding path differs (:110-114)
and "mf-fixed-manifold" drops the orthogonal and outer terms (:134-136)."""
import torch

from ..b200 import ops
from ..transforms import fusion
from ..transforms.base import _expand_lad
from .base import host_staged
from .manifold_flow import ManifoldFlow


class EncoderManifoldFlow(ManifoldFlow):
    def __init__(self, data_dim, latent_dim, encoder, outer_transform, inner_transform=None, pie_epsilon=1.0e-2, apply_context_to_outer=True,
                 clip_pie=False):
        super().__init__(data_dim, latent_dim, outer_transform, inner_transform, pie_epsilon, apply_context_to_outer, clip_pie)
        self.encoder = encoder

    @host_staged
    def forward(self, x, mode="mf", context=None, return_hidden=False):
        if mode not in ("mf", "projection", "pie", "mf-fixed-manifold"):
            raise ValueError("unsupported mode {!r}".format(mode))
        fusion.begin_pass(self)
        u, h_manifold, log_det_inner = self._encode(x, context)
        x_reco, inv_log_det_inner, inv_log_det_outer, inv_jacobian_outer, _ = self._decode(u, mode=mode, context=context)
        if mode == "mf":
            log_prob = self._log_prob("mf", u, None, None, None, inv_log_det_inner, None, inv_jacobian_outer)
        elif mode == "mf-fixed-manifold":
            log_prob = ops.gauss_logprob(u, add0=-inv_log_det_inner)
        else:
            log_prob = None
        if return_hidden:
            return x_reco, log_prob, u, h_manifold
        return x_reco, log_prob, u

    @host_staged
    def encode(self, x, context=None):
        fusion.begin_pass(self)
        return self._encode(x, context)[0]

    def _encode(self, x, context=None):
        h_manifold = self.encoder(x, context=self._outer_context(context))
        u, log_det_inner = self.inner_transform._apply_transform(h_manifold, context, False)
        return u, h_manifold, _expand_lad(log_det_inner, x)
