"""ActNorm (reference: manifold_flow/transforms/normalization.py:70-147).  ``BatchNorm`` as a
transform is never constructed by experiments/architectures and is out of scope."""
import torch
from torch import nn

from ..utils import various
from .base import Transform


class ActNorm(Transform):
    """y = exp(log_scale) * x + shift per channel / feature, with data-dependent initialisation on
    the first training batch.  ``initialized`` stays a plain Python attribute, exactly like the
    reference (:83), so ``experiments/train.py:fix_act_norm_issue`` keeps working."""

    def __init__(self, features):
        if not various.is_positive_int(features):
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.initialized = False
        self.log_scale = nn.Parameter(torch.zeros(features))
        self.shift = nn.Parameter(torch.zeros(features))

    @property
    def scale(self):
        return torch.exp(self.log_scale)

    def _can_fuse(self, x):
        return x.dim() in (2, 4) and self.log_scale.numel() <= 96

    def _needs_data_init(self):
        return self.training and not self.initialized

    def _affine_map(self, inverse):
        if not inverse:  # :99-116
            return torch.diag(self.scale), self.shift, torch.sum(self.log_scale)
        inv_scale = torch.exp(-self.log_scale)  # :118-131
        return torch.diag(inv_scale), -self.shift * inv_scale, -torch.sum(self.log_scale)

    def _apply_transform(self, inputs, context, inverse):
        if inputs.dim() not in (2, 4):
            raise ValueError("Expecting inputs to be a 2D or a 4D tensor.")
        if not self._can_fuse(inputs):
            raise NotImplementedError("ActNorm over more than 96 features is not supported")
        if not inverse and self._needs_data_init():  # forward only, like the reference (:103-104)
            self._initialize(inputs)
        from .fusion import apply_affine

        return apply_affine([self._affine_map(inverse)], inputs)

    def _initialize(self, inputs):
        """log_scale = -log std, shift = -mean(x / std) over batch (and pixels).  :134-147.
        Under torch.distributed the statistics are those of rank 0's batch, broadcast to every
        rank, so all replicas start from identical parameters."""
        with torch.no_grad():
            flat = inputs.permute(0, 2, 3, 1).reshape(-1, inputs.shape[1]) if inputs.dim() == 4 else inputs
            std = flat.std(dim=0)
            mu = (flat / std).mean(dim=0)
            log_scale, shift = -torch.log(std), -mu
            if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
                torch.distributed.broadcast(log_scale, src=0)
                torch.distributed.broadcast(shift, src=0)
            self.log_scale.data.copy_(log_scale)
            self.shift.data.copy_(shift)
        self.initialized = True


class BatchNorm:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("transforms.BatchNorm is never constructed by the reference factories; out of scope")
