"""Squeeze / reshape transforms (reference: manifold_flow/transforms/reshape.py)."""
import torch

from ..b200 import ops
from ..utils import various
from .base import Transform


class SqueezeTransform(Transform):
    """2x2 space-to-depth: [B, C, H, W] <-> [B, 4C, H/2, W/2] with channel order (c, fh, fw)
    (reshape.py:29-57) as one gather launch instead of view/permute/contiguous."""

    def __init__(self, factor=2):
        super().__init__()
        if not various.is_int(factor) or factor <= 1:
            raise ValueError("Factor must be an integer > 1.")
        if factor != 2:
            raise NotImplementedError("only factor=2 is used by the reference architectures")
        self.factor = factor

    def get_output_shape(self, c, h, w):
        return (c * self.factor * self.factor, h // self.factor, w // self.factor)

    def _apply_transform(self, inputs, context, inverse):
        if inputs.dim() != 4:
            raise ValueError("Expecting inputs with 4 dimensions")
        return ops.squeeze2x2(inputs, inverse=inverse), None


class ReshapeTransform(Transform):
    """reshape.py:60-74"""

    def __init__(self, input_shape, output_shape):
        super().__init__()
        self.input_shape = input_shape
        self.output_shape = output_shape

    def _apply_transform(self, inputs, context, inverse):
        src, dst = (self.output_shape, self.input_shape) if inverse else (self.input_shape, self.output_shape)
        if tuple(inputs.shape[1:]) != tuple(src):
            raise RuntimeError("Unexpected inputs shape ({}, but expecting {})".format(tuple(inputs.shape[1:]), src))
        return inputs.reshape(-1, *dst), None
