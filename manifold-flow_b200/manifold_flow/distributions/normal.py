"""Base densities of the flows (reference: manifold_flow/distributions/normal.py:10-75).
The log-densities are one fused reduction launch (``mflow_gauss_logprob``)."""
import torch

from ..b200 import ops
from ..utils import various
from .base import Distribution


def _default_device():
    # samples are created on the GPU; honour an explicit global default device if one was set
    # (the reference trainer flips the default tensor type to CUDA, training/trainer.py:32-39)
    d = torch.empty(0).device
    return d if d.type == "cuda" else torch.device("cuda", torch.cuda.current_device())


class StandardNormal(Distribution):
    """N(0, 1) on ``shape`` (normal.py:10-39)."""

    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)

    def _log_prob(self, inputs, context):
        if inputs.shape[1:] != self._shape:
            raise ValueError("Expected input of shape {}, got {}".format(self._shape, inputs.shape[1:]))
        return ops.gauss_logprob(inputs.reshape(inputs.shape[0], -1))

    def _sample(self, num_samples, context):
        if context is None:
            return torch.randn(num_samples, *self._shape, device=_default_device())
        samples = torch.randn(context.shape[0] * num_samples, *self._shape, device=context.device)
        return various.split_leading_dim(samples, [context.shape[0], num_samples])


class RescaledNormal(Distribution):
    """N(0, std^2) with optional clipping of the inputs (normal.py:42-75)."""

    def __init__(self, shape, std=1.0, clip=10.0):
        super().__init__()
        self._shape = torch.Size(shape)
        self.std = std
        self._clip = clip

    def _log_prob(self, inputs, context):
        if inputs.shape[1:] != self._shape:
            raise ValueError("Expected input of shape {}, got {}".format(self._shape, inputs.shape[1:]))
        flat = inputs.reshape(inputs.shape[0], -1)
        empty = flat[:, :0]
        return ops.gauss_logprob(empty, flat, eps=self.std, clip=self._clip)

    def _sample(self, num_samples, context):
        if context is None:
            return self.std * torch.randn(num_samples, *self._shape, device=_default_device())
        samples = self.std * torch.randn(context.shape[0] * num_samples, *self._shape, device=context.device)
        return various.split_leading_dim(samples, [context.shape[0], num_samples])
