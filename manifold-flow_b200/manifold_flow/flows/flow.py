"""Ambient normalizing flow (reference: manifold_flow/flows/flow.py:10-75)."""
import torch

from .. import distributions
from ..b200 import ops
from ..transforms import fusion
from ..transforms.base import _expand_lad
from ..utils.various import product
from .base import BaseFlow, host_staged


class Flow(BaseFlow):
    def __init__(self, data_dim, transform):
        super().__init__()
        self.data_dim = data_dim
        self.latent_dim = data_dim
        self.total_data_dim = product(data_dim)
        self.total_latent_dim = product(self.latent_dim)
        self.latent_distribution = distributions.StandardNormal((self.total_latent_dim,))
        self.transform = transform
        self._report_model_parameters()

    @host_staged
    def forward(self, x, context=None):
        """(x_reco, log_prob, u) - flow.py:26-39.  The reconstruction only feeds diagnostics in the
        reference's trainers (the NLL loss does not depend on it), so it is computed without
        recording an autograd graph."""
        fusion.begin_pass(self)
        u, log_det = self._encode(x, context=context)
        with torch.no_grad():
            x_reco = self.decode(u.detach(), context=context)
        return x_reco, self._density(u, log_det), u

    @host_staged
    def encode(self, x, context=None):
        fusion.begin_pass(self)
        return self._encode(x, context=context)[0]

    @host_staged
    def decode(self, u, context=None):
        return self.transform.inverse(u, context=context)[0]

    @host_staged
    def log_prob(self, x, context=None):
        """flow.py:53-63"""
        fusion.begin_pass(self)
        u, log_det = self._encode(x, context)
        return self._density(u, log_det)

    @host_staged
    def sample(self, u=None, n=1, context=None):
        if u is None:
            u = self.latent_distribution.sample(n, context=None)
        return self.decode(u, context=context)

    def _encode(self, x, context=None):
        u, log_det = self.transform._apply_transform(x, context, False)
        return u, _expand_lad(log_det, x)

    def _density(self, u, log_det):
        return ops.gauss_logprob(u.reshape(u.shape[0], -1), add0=log_det)
