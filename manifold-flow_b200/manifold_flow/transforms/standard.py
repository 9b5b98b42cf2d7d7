"""Identity and scalar affine transforms (reference: manifold_flow/transforms/standard.py)."""
import numpy as np
import torch

from ..b200 import ops
from .base import Transform


class IdentityTransform(Transform):
    def _apply_transform(self, inputs, context, inverse):
        return inputs, None


class AffineScalarTransform(Transform):
    """outputs = inputs * scale + shift with scalar buffers ``_shift`` / ``_scale``
    (standard.py:23-57); logabsdet = D * log|scale|."""

    def __init__(self, shift=None, scale=None):
        super().__init__()
        if shift is None and scale is None:
            raise ValueError("At least one of scale and shift must be provided.")
        if scale == 0.0:
            raise ValueError("Scale cannot be zero.")
        self.register_buffer("_shift", torch.tensor(shift if (shift is not None) else 0.0))
        self.register_buffer("_scale", torch.tensor(scale if (scale is not None) else 1.0))
        self._host = None

    @property
    def _log_scale(self):
        return torch.log(torch.abs(self._scale))

    def _host_values(self):
        key = (self._scale._version, self._shift._version, self._scale.data_ptr())
        if self._host is None or self._host[0] != key:
            self._host = (key, float(self._scale), float(self._shift))  # one sync, then cached
        return self._host[1], self._host[2]

    def _apply_transform(self, inputs, context, inverse):
        scale, shift = self._host_values()
        num_dims = int(np.prod(inputs.shape[1:]))
        lad = torch.full((), float(np.log(abs(np.float32(scale)))) * num_dims, dtype=inputs.dtype, device=inputs.device)
        return ops.affine_scalar(inputs, scale, shift, inverse), (-lad if inverse else lad)
