"""Distribution protocol (reference: manifold_flow/distributions/base.py)."""
import torch
from torch import nn

from ..utils import various


class NoMeanException(Exception):
    pass


class Distribution(nn.Module):
    def forward(self, *args):
        raise RuntimeError("Forward method cannot be called for a Distribution object.")

    def log_prob(self, inputs, context=None):
        inputs = torch.as_tensor(inputs)
        if context is not None:
            context = torch.as_tensor(context)
            if inputs.shape[0] != context.shape[0]:
                raise ValueError("Number of input items must be equal to number of context items.")
        return self._log_prob(inputs, context)

    def _log_prob(self, inputs, context):
        raise NotImplementedError()

    def sample(self, num_samples, context=None, batch_size=None):
        if not various.is_positive_int(num_samples):
            raise TypeError("Number of samples must be a positive integer.")
        if context is not None:
            context = torch.as_tensor(context)
        if batch_size is None:
            return self._sample(num_samples, context)
        if not various.is_positive_int(batch_size):
            raise TypeError("Batch size must be a positive integer.")
        sizes = [batch_size] * (num_samples // batch_size)
        if num_samples % batch_size:
            sizes.append(num_samples % batch_size)
        return torch.cat([self._sample(s, context) for s in sizes], dim=0)

    def _sample(self, num_samples, context):
        raise NotImplementedError()
