"""Abstract flow (reference: manifold_flow/flows/base.py:7-34)."""
import logging

from torch import nn

logger = logging.getLogger(__name__)


class BaseFlow(nn.Module):
    def forward(self, x, context=None):
        raise NotImplementedError

    def encode(self, x, context=None):
        raise NotImplementedError

    def decode(self, u, context=None):
        raise NotImplementedError

    def project(self, x, context=None):
        return self.decode(self.encode(x, context), context)

    def log_prob(self, x, context=None):
        raise NotImplementedError

    def sample(self, u=None, n=1, context=None):
        raise NotImplementedError

    def _report_model_parameters(self):
        n_all = sum(p.numel() for p in self.parameters())
        n_train = sum(p.numel() for p in self.parameters() if p.requires_grad)
        logger.info("Model has %.1f M parameters (%.1f M trainable) with an estimated size of %.1f MB", n_all / 1e6, n_train / 1e6, n_all * 4 / 1e6)
