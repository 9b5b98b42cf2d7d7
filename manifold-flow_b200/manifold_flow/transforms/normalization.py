"""ActNorm (reference: manifold_flow/transforms/normalization.py:70-147).  ``BatchNorm`` as a
transform is never constructed by experiments/architectures and is out of scope."""
import torch
from torch import nn

from ..b200.config import invalidate_caches
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
        group = getattr(self, "_an_group", None)
        if group is not None and group.ready():
            from .fusion import member_views

            mat, bias, lad = member_views(group, "_members_inv" if inverse else "_members_fwd", group.affine_maps(inverse))
            return mat[self._an_index], bias[self._an_index], lad[self._an_index]
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
            # written through the parameters themselves (not .data) so that their version counters move and every
            # cache keyed on them (composed channel mixes, grouped maps) is rebuilt
            self.log_scale.copy_(log_scale)
            self.shift.copy_(shift)
        invalidate_caches()
        self.initialized = True


class ActNormGroup:
    """Stacked affine maps of all same-width ActNorm layers of a model (one exp / diag_embed / sum instead of one
    per layer).  Used only once every member has had its data-dependent initialisation (which writes the
    parameters behind autograd's back, so a cache built earlier could go stale)."""

    def __init__(self, layers):
        from .fusion import VersionedCache

        self.layers = list(layers)
        self._maps = {False: VersionedCache(), True: VersionedCache()}

    def ready(self):
        return all(layer.initialized or not layer.training for layer in self.layers)

    def affine_maps(self, inverse):
        def build():
            log_scale = torch.stack([layer.log_scale for layer in self.layers])
            shift = torch.stack([layer.shift for layer in self.layers])
            if not inverse:
                return torch.diag_embed(torch.exp(log_scale)), shift, log_scale.sum(dim=1)
            inv_scale = torch.exp(-log_scale)
            return torch.diag_embed(inv_scale), -shift * inv_scale, -log_scale.sum(dim=1)

        params = tuple(p for layer in self.layers for p in (layer.log_scale, layer.shift))
        return self._maps[bool(inverse)].get(params, build)


class BatchNorm:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("transforms.BatchNorm is never constructed by the reference factories; out of scope")
