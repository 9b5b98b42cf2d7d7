"""Function-level API of the rational-quadratic spline
(reference: manifold_flow/transforms/splines/rational_quadratic.py:16-28,67-81).

One fused launch of ``mflow_rqs_apply`` (elementwise mode) instead of ~60 ATen launches with
boolean-mask gathers and host syncs.  The coupling layers do not go through here: they hand the
conditioner output to the kernel directly (no concatenation, no pre-scale pass).
"""
import torch

from ...b200 import ops

DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3


def unconstrained_rational_quadratic_spline(
    inputs,
    unnormalized_widths,
    unnormalized_heights,
    unnormalized_derivatives,
    inverse=False,
    tails="linear",
    tail_bound=1.0,
    min_bin_width=DEFAULT_MIN_BIN_WIDTH,
    min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
    min_derivative=DEFAULT_MIN_DERIVATIVE,
    full_jacobian=False,
):
    """RQ spline on [-tail_bound, tail_bound]^2 with identity (linear) tails; returns
    (outputs, logabsdet), both shaped like ``inputs``.  widths/heights: [..., K], derivatives:
    [..., K-1] (the two boundary derivatives are fixed to 1, :38-41)."""
    assert not full_jacobian
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    num_bins = unnormalized_widths.shape[-1]
    if unnormalized_heights.shape[-1] != num_bins or unnormalized_derivatives.shape[-1] != num_bins - 1:
        raise ValueError("expected K widths, K heights and K-1 derivatives")
    if min_bin_width * num_bins > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    params = torch.cat((unnormalized_widths, unnormalized_heights, unnormalized_derivatives), dim=-1).reshape(-1, 3 * num_bins - 1)
    outputs, logabsdet = ops.rqs_elementwise(
        inputs.reshape(-1), params, num_bins, tail_bound, min_bin_width, min_bin_height, min_derivative, inverse
    )
    return outputs.view(inputs.shape), logabsdet.view(inputs.shape)


def rational_quadratic_spline(
    inputs,
    unnormalized_widths,
    unnormalized_heights,
    unnormalized_derivatives,
    inverse=False,
    left=0.0,
    right=1.0,
    bottom=0.0,
    top=1.0,
    min_bin_width=DEFAULT_MIN_BIN_WIDTH,
    min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
    min_derivative=DEFAULT_MIN_DERIVATIVE,
    full_jacobian=False,
):
    """RQ spline mapping [left, right] onto [bottom, top] (:67-168): widths/heights [..., K], derivatives [..., K+1]
    (every knot derivative is a parameter here - the unconstrained wrapper is what pads the two boundary ones).
    Returns (outputs, logabsdet) shaped like ``inputs``; ``InputOutsideDomain`` if an input lies outside
    [left, right] (:85-86, the same check for both directions, as in the reference) or, for the inverse, outside
    [bottom, top] (where the reference's gather fails with an index error).  One ``mflow_rqs_box_apply`` launch."""
    from ..base import InputOutsideDomain

    assert not full_jacobian
    if inputs.numel() and (torch.min(inputs) < left or torch.max(inputs) > right):
        raise InputOutsideDomain()
    if inverse and inputs.numel() and (torch.min(inputs) < bottom or torch.max(inputs) > top):
        raise InputOutsideDomain()
    num_bins = unnormalized_widths.shape[-1]
    if unnormalized_heights.shape[-1] != num_bins or unnormalized_derivatives.shape[-1] != num_bins + 1:
        raise ValueError("expected K widths, K heights and K+1 derivatives")
    if min_bin_width * num_bins > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    params = torch.cat((unnormalized_widths, unnormalized_heights, unnormalized_derivatives), dim=-1).reshape(-1, 3 * num_bins + 1)
    outputs, logabsdet = ops.rqs_box(inputs.reshape(-1), params, num_bins, left, right, bottom, top, min_bin_width, min_bin_height,
                                     min_derivative, inverse)
    return outputs.view(inputs.shape), logabsdet.view(inputs.shape)
