"""Rational-quadratic spline coupling layer (reference: manifold_flow/transforms/coupling.py).

``PiecewiseRationalQuadraticCouplingTransform`` (:458-533) is the only coupling family any
BASELINE config selects; the affine / additive / linear / quadratic / cubic families are out of
scope (stubs in ``out_of_scope.py``).  One layer evaluation is: slice the identity features (a view),
run the conditioner, and hand the full activation tensor plus the raw conditioner output to ONE
``mflow_rqs_apply`` launch that does parameter slicing, the 1/sqrt(hidden) pre-scale, bin
construction, search, the spline (or its inverse), the tails, the per-sample log-det reduction and
the identity pass-through (:58-129, :131-185, :264-279, :501-533).
"""
import warnings

import numpy as np
import torch

from ..b200 import conditioner, ops
from ..b200.cabi import require_cuda
from .base import Transform
from .splines import rational_quadratic as rq


class CouplingTransform(Transform):
    """Splits features by a 1-D mask: mask > 0 transformed, mask <= 0 passed through and fed to
    the conditioner.  2-D ([B, D]) and 4-D ([B, C, H, W], split on channels) inputs.  :14-197"""

    def __init__(self, mask, transform_net_create_fn, unconditional_transform=None):
        mask = torch.as_tensor(mask)
        if mask.dim() != 1:
            raise ValueError("Mask must be a 1-dim tensor.")
        if mask.numel() <= 0:
            raise ValueError("Mask can't be empty.")
        super().__init__()
        self.features = len(mask)
        index = torch.arange(self.features)
        self.register_buffer("identity_features", index.masked_select(mask <= 0))
        self.register_buffer("transform_features", index.masked_select(mask > 0))
        assert self.num_identity_features + self.num_transform_features == self.features
        self.transform_net = transform_net_create_fn(self.num_identity_features, self.num_transform_features * self._transform_dim_multiplier())
        if unconditional_transform is not None:
            raise NotImplementedError("apply_unconditional_transform=True is never used by the reference factories; out of scope")
        self.unconditional_transform = None
        self._geometry = ops.SplineGeometry.from_indices(self.transform_features.tolist(), self.identity_features.tolist())

    @property
    def num_identity_features(self):
        return len(self.identity_features)

    @property
    def num_transform_features(self):
        return len(self.transform_features)

    def _check(self, inputs):
        if inputs.dim() not in [2, 4]:
            raise ValueError("Inputs must be a 2D or a 4D tensor.")
        if inputs.shape[1] != self.features:
            raise ValueError("Expected features = {}, got {}.".format(self.features, inputs.shape[1]))

    def _identity_view(self, inputs):
        g = self._geometry
        if g is None:
            return torch.index_select(inputs, 1, self.identity_features)
        return inputs[:, g.i_offset :: g.i_step][:, : g.n_identity]

    def _transform_dim_multiplier(self):
        raise NotImplementedError()


class PiecewiseRationalQuadraticCouplingTransform(CouplingTransform):
    def __init__(
        self,
        mask,
        transform_net_create_fn,
        num_bins=10,
        tails=None,
        tail_bound=1.0,
        apply_unconditional_transform=False,
        img_shape=None,
        min_bin_width=rq.DEFAULT_MIN_BIN_WIDTH,
        min_bin_height=rq.DEFAULT_MIN_BIN_HEIGHT,
        min_derivative=rq.DEFAULT_MIN_DERIVATIVE,
    ):
        if tails != "linear":
            raise NotImplementedError("only tails='linear' is in scope (every BASELINE config; coupling.py:495-499)")
        if apply_unconditional_transform:
            raise NotImplementedError("apply_unconditional_transform=True is never used by the reference factories; out of scope")
        self.num_bins = num_bins
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.min_derivative = min_derivative
        self.tails = tails
        self.tail_bound = tail_bound
        super().__init__(mask, transform_net_create_fn, unconditional_transform=None)

    def _transform_dim_multiplier(self):
        return self.num_bins * 3 - 1  # K widths, K heights, K-1 interior derivatives (:495-499)

    def _divisor(self):
        net = self.transform_net
        if hasattr(net, "hidden_features"):
            return float(np.sqrt(net.hidden_features))  # :506-508
        if hasattr(net, "hidden_channels"):
            return float(np.sqrt(net.hidden_channels))  # :509-511
        warnings.warn("Inputs to the softmax are not scaled down: initialization might be bad.")
        return 1.0

    def _apply_transform(self, inputs, context, inverse):
        self._check(inputs)
        require_cuda(inputs)
        args = (self.num_bins, self.tail_bound, self.min_bin_width, self.min_bin_height, self.min_derivative, self._divisor(), inverse)
        if self._geometry is not None:
            if inputs.dim() == 2 and inputs.shape[1] <= ops.ROWTILE_MAX_FEATURES:
                # a vector conditioner that runs on the convolution kernels may hand over its NCHW result as it is
                with conditioner.planes_output():
                    params = self.transform_net(self._identity_view(inputs), context)
            else:
                net = self.transform_net
                final = getattr(net, "final_layer", None)
                if (hasattr(net, "hidden") and isinstance(final, torch.nn.Conv2d)
                        and ops.conv_rqs_supported(inputs, final.weight, inputs, self.num_bins)):
                    # evaluation pass of an image coupling: the spline runs in the epilogue of the final 1x1 convolution
                    hidden = net.hidden(self._identity_view(inputs), context)
                    return ops.conv_rqs(hidden, final.weight, final.bias, inputs, self._geometry, *args)
                params = self.transform_net(self._identity_view(inputs), context)
            return ops.rqs_coupling(inputs, params, self._geometry, *args)
        params = self.transform_net(self._identity_view(inputs), context)
        # irregular mask: pack the transformed features, run the same kernel, scatter back
        packed = torch.index_select(inputs, 1, self.transform_features)
        geom = ops.SplineGeometry(self.num_transform_features, 0, 1, 0, 0, 1)
        out_t, lad = ops.rqs_coupling(packed, params, geom, *args)
        return inputs.index_copy(1, self.transform_features, out_t), lad
