"""Fixed permutations of the feature / channel axis
(reference: manifold_flow/transforms/permutations.py)."""
import torch

from ..b200 import ops
from ..b200.cabi import require_cuda
from ..b200.config import config
from ..utils import various
from .base import Transform


class Permutation(Transform):
    """outputs = inputs.index_select(dim, permutation)  (:73); inverse with the arg-sorted
    permutation (:27-29), computed once per device instead of on every call."""

    def __init__(self, permutation, dim=1):
        if permutation.ndimension() != 1:
            raise ValueError("Permutation must be a 1D tensor.")
        if not various.is_positive_int(dim):
            raise ValueError("dim must be a positive integer.")
        super().__init__()
        self._dim = dim
        self.register_buffer("_permutation", permutation)
        self._inv_cache = None

    @property
    def _inverse_permutation(self):
        perm = self._permutation
        c = self._inv_cache
        if c is None or c[0] != (perm.data_ptr(), perm._version, perm.device):
            self._inv_cache = ((perm.data_ptr(), perm._version, perm.device), torch.argsort(perm))
        return self._inv_cache[1]

    def _can_fuse(self, x):
        # tiny vectors: fold the permutation into the neighbouring dense channel mix
        return self._dim == 1 and x.dim() == 2 and self._permutation.numel() <= 96

    def _affine_map(self, inverse):
        n = self._permutation.numel()
        idx = self._inverse_permutation if inverse else self._permutation
        eye = torch.eye(n, dtype=torch.float32, device=idx.device)
        zero = torch.zeros((), dtype=torch.float32, device=idx.device)
        return eye[idx, :], torch.zeros(n, dtype=torch.float32, device=idx.device), zero

    def _jacobian(self, inputs, context, inverse):
        """Jacobian as the reference states it (permutations.py:38-71): ones at
        [permutation[j], j].  For outputs[j] = inputs[permutation[j]] the exact Jacobian is the
        transpose of that matrix (they coincide only for involutions); the reference's mode="mf"
        likelihoods are built from its own matrix, so it is the default here.  Set
        ``manifold_flow.b200.config.exact_permutation_jacobian = True`` for the exact one."""
        outputs, _ = self._apply_transform(inputs, context, inverse)
        if inputs.dim() != 2 or self._dim != 1:
            raise NotImplementedError("full_jacobian is only available for [B, D] inputs")
        idx = self._inverse_permutation if inverse else self._permutation
        n = idx.numel()
        # ones at [idx[j], j] (the reference's matrix) or at [j, idx[j]] (exact); scatter with a scalar value - an indexed
        # assignment of 1.0 would stage a CPU scalar tensor, which a CUDA-graph capture of the evaluation rejects
        jac = torch.zeros(n, n, dtype=inputs.dtype, device=inputs.device)
        jac.scatter_(1 if config.exact_permutation_jacobian else 0, idx.view(n, 1) if config.exact_permutation_jacobian else idx.view(1, n), 1.0)
        return outputs.detach(), jac.unsqueeze(0).expand(inputs.shape[0], -1, -1)

    def _apply_transform(self, inputs, context, inverse):
        require_cuda(inputs)
        dim = self._dim
        if dim >= inputs.ndimension():
            raise ValueError("No dimension {} in inputs.".format(dim))
        if inputs.shape[dim] != len(self._permutation):
            raise ValueError("Dimension {} in inputs must be of size {}.".format(dim, len(self._permutation)))
        fwd, inv = self._permutation, self._inverse_permutation
        if inverse:
            fwd, inv = inv, fwd
        if dim == 1 and inputs.dim() == 2:
            return ops.permute_rows(inputs, fwd, inv), None
        return torch.index_select(inputs, dim, fwd), None


class MaskBasedPermutation(Permutation):
    """Active entries of a binary mask first, then the inactive ones (:84-90)."""

    def __init__(self, mask, dim=1):
        mask = torch.as_tensor(mask)
        idx = torch.cat((torch.nonzero(mask).squeeze(), torch.nonzero(1 - mask).squeeze()), 0)
        super().__init__(idx, dim)


class RandomPermutation(Permutation):
    def __init__(self, features, dim=1):
        if not various.is_positive_int(features):
            raise ValueError("Number of features must be a positive integer.")
        super().__init__(torch.randperm(features), dim)


class ReversePermutation(Permutation):
    def __init__(self, features, dim=1):
        if not various.is_positive_int(features):
            raise ValueError("Number of features must be a positive integer.")
        super().__init__(torch.arange(features - 1, -1, -1), dim)
