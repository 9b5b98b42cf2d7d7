"""Invertible 1x1 convolution (reference: manifold_flow/transforms/conv.py:5-55): a fixed random
channel permutation followed by the LU map applied at every pixel."""
import torch

from .fusion import MAX_FUSED_CHANNELS
from .lu import LULinear
from .permutations import RandomPermutation


class OneByOneConvolution(LULinear):
    def __init__(self, num_channels, using_cache=False, identity_init=True):
        super().__init__(num_channels, using_cache, identity_init)
        self.permutation = RandomPermutation(num_channels, dim=1)

    def _can_fuse(self, x):
        return x.dim() == 4

    def _affine_map(self, inverse):
        """y = W x[perm] + b  (conv.py:33-41)  /  x = (W^-1 (y - b))[inv_perm]  (:43-55)."""
        group = getattr(self, "_lu_group", None)
        if group is not None:   # all same-width 1x1 convolutions of the model at once (lu.LUGroup)
            from .fusion import member_views

            mat, bias, lad = member_views(group, "_members_inv" if inverse else "_members_fwd", group.affine_maps(inverse))
            return mat[self._lu_index], bias[self._lu_index], lad[self._lu_index]
        perm = self.permutation._permutation
        inv_perm = self.permutation._inverse_permutation

        def build():
            if not inverse:
                return self.weight()[:, inv_perm], self.bias, self.logabsdet()
            m_inv = self.weight_inverse()[inv_perm, :]
            return m_inv, -(m_inv @ self.bias), -self.logabsdet()

        return self._map_cache[bool(inverse)].get(self._lu_params() + (self.bias, perm), build)

    def _apply_transform(self, inputs, context, inverse):
        if inputs.dim() != 4:
            raise ValueError("Inputs must be a 4D tensor.")
        if inputs.shape[1] != self.features:
            raise ValueError("Expected {} channels, got {}".format(self.features, inputs.shape[1]))
        if self.features > MAX_FUSED_CHANNELS:
            raise NotImplementedError("1x1 convolutions with more than {} channels are not supported".format(MAX_FUSED_CHANNELS))
        from .fusion import apply_affine

        return apply_affine([self._affine_map(inverse)], inputs)
