"""Transform protocol, composition and multi-scale split
(reference: manifold_flow/transforms/base.py).

Public contract (base.py:25-32): ``forward(inputs, context=None, full_jacobian=False)`` and
``inverse(...)`` return ``(outputs, logabsdet[B])``, or ``(outputs, jacobian[B, D, D])`` when
``full_jacobian=True``.

Internally every transform implements ``_apply(inputs, context, inverse) -> (outputs, logabsdet)``
where ``logabsdet`` may be ``None`` (zero), a 0-dim tensor (the same value for every sample: ActNorm,
LU layers, scalar affine maps) or a ``[B]`` tensor; compositions add the cheap forms first and
broadcast once, and fuse neighbouring linear layers into a single channel-mix launch.
"""
import numpy as np
import torch
from torch import nn

from ..utils import various


class InverseNotAvailable(Exception):
    """Raised by transforms that have no inverse (base.py:13-16)."""


class InputOutsideDomain(Exception):
    """Raised when an input is outside a transform's domain (base.py:19-22)."""


def _expand_lad(lad, inputs):
    b = inputs.shape[0]
    if lad is None:
        return torch.zeros(b, dtype=inputs.dtype, device=inputs.device)
    if lad.dim() == 0:
        return lad.expand(b).clone() if not lad.requires_grad else lad * torch.ones(b, dtype=inputs.dtype, device=inputs.device)
    return lad


def _add_lad(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return a + b


def batch_jacobian(fn, inputs):
    """[B, D_out, D_in] Jacobian of a per-sample map by D_out reverse passes over the batch sum
    (the reference differentiates B * D_out scalars one by one: utils/various.py:22-45,68-87)."""
    from ..b200 import ops

    with torch.enable_grad(), ops.input_grads_only():   # d outputs / d inputs only: no parameter-gradient kernels
        x = inputs.detach().requires_grad_(True)
        y = fn(x)
        flat = y.reshape(y.shape[0], -1)
        rows = []
        for i in range(flat.shape[1]):
            (g,) = torch.autograd.grad(flat[:, i].sum(), x, retain_graph=True)
            rows.append(g.reshape(x.shape[0], -1))
    return y.detach(), torch.stack(rows, dim=1)


class Transform(nn.Module):
    """Base class of all transforms (base.py:25-32)."""

    def _apply_transform(self, inputs, context, inverse):
        raise NotImplementedError()

    def forward(self, inputs, context=None, full_jacobian=False):
        return self._public(inputs, context, full_jacobian, False)

    def inverse(self, inputs, context=None, full_jacobian=False):
        return self._public(inputs, context, full_jacobian, True)

    def _public(self, inputs, context, full_jacobian, inverse):
        if full_jacobian:
            return self._jacobian(inputs, context, inverse)
        outputs, lad = self._apply_transform(inputs, context, inverse)
        return outputs, _expand_lad(lad, inputs)

    def _jacobian(self, inputs, context, inverse):
        """(outputs, jacobian[B, D, D]) of this layer: one reverse pass per output coordinate over
        the whole batch (coupling.py:71-105,141-170 does B * D autograd calls instead)."""
        return batch_jacobian(lambda x: self._apply_transform(x, context, inverse)[0], inputs)


class CompositeTransform(Transform):
    """Chains transforms in order (base.py:35-84); state_dict prefix ``_transforms.{i}``."""

    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    def _apply_transform(self, inputs, context, inverse):
        from .fusion import apply_chain

        return apply_chain(list(self._transforms), inputs, context, inverse)

    def _jacobian(self, inputs, context, inverse):
        """Per-layer Jacobians multiplied in evaluation order, as base.py:52-69 does."""
        seq = list(self._transforms)[::-1] if inverse else list(self._transforms)
        outputs, total = inputs, None
        for transform in seq:
            outputs, jac = transform._jacobian(outputs, context, inverse)
            total = jac if total is None else torch.bmm(jac, total)
        return outputs, total


class MultiscaleCompositeTransform(Transform):
    """RealNVP-style multi-scale composition (base.py:87-245): after every transform but the last,
    half of the channels leave the chain; all outputs are flattened and concatenated."""

    def __init__(self, num_transforms, split_dim=1):
        if not various.is_positive_int(split_dim):
            raise TypeError("Split dimension must be a positive integer.")
        super().__init__()
        self._transforms = nn.ModuleList()
        self._output_shapes = []
        self._num_transforms = num_transforms
        self._split_dim = split_dim

    def add_transform(self, transform, transform_output_shape):
        """Returns the input shape of the next transform (None after the last one); base.py:112-145."""
        if len(self._transforms) == self._num_transforms:
            raise RuntimeError("Adding more than {} transforms is not allowed.".format(self._num_transforms))
        axis = self._split_dim - 1
        if axis >= len(transform_output_shape):
            raise ValueError("No split_dim in output shape")
        if transform_output_shape[axis] < 2:
            raise ValueError("Size of dimension {} must be at least 2.".format(self._split_dim))
        self._transforms.append(transform)
        shape = tuple(transform_output_shape)
        if len(self._transforms) == self._num_transforms:
            self._output_shapes.append(shape)
            return None
        size = shape[axis]
        self._output_shapes.append(shape[:axis] + ((size + 1) // 2,) + shape[axis + 1 :])
        return shape[:axis] + (size // 2,) + shape[axis + 1 :]

    def _check(self):
        if self._num_transforms != len(self._transforms):
            raise RuntimeError("Expecting exactly {} transform(s) to be added.".format(self._num_transforms))

    def _apply_transform(self, inputs, context, inverse):
        self._check()
        b = inputs.shape[0]
        total = None
        last = len(self._transforms) - 1
        if not inverse:  # base.py:147-196
            if self._split_dim >= inputs.dim():
                raise ValueError("No split_dim in inputs.")
            pieces, hidden = [], inputs
            for i, transform in enumerate(self._transforms):
                out, lad = transform._apply_transform(hidden, context, False)
                total = _add_lad(total, lad)
                if i < last:
                    out, hidden = torch.chunk(out, chunks=2, dim=self._split_dim)
                    assert out.shape[1:] == self._output_shapes[i]
                pieces.append(out.reshape(b, -1))
            return torch.cat(pieces, dim=-1), total
        # base.py:198-245
        if inputs.dim() != 2:
            raise ValueError("Expecting NxD inputs")
        sizes = [int(np.prod(s)) for s in self._output_shapes]
        chunks = [c.reshape(b, *s) for c, s in zip(torch.split(inputs, sizes, dim=1), self._output_shapes)]
        hidden, total = self._transforms[last]._apply_transform(chunks[last], context, True)
        for i in range(last - 1, -1, -1):
            hidden, lad = self._transforms[i]._apply_transform(torch.cat([chunks[i], hidden], dim=self._split_dim), context, True)
            total = _add_lad(total, lad)
        return hidden, total


class InverseTransform(Transform):
    """Swaps forward and inverse of a transform (base.py:248-264)."""

    def __init__(self, transform):
        super().__init__()
        self._transform = transform

    def _apply_transform(self, inputs, context, inverse):
        return self._transform._apply_transform(inputs, context, not inverse)
