"""Projection onto the first n latent coordinates and zero-padding back
(reference: manifold_flow/transforms/projections.py:10-47)."""
import torch
from torch.nn import functional as F

from ..utils.various import product
from .base import Transform


class ProjectionSplit(Transform):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.input_dim_total = product(input_dim)
        self.output_dim_total = product(output_dim)
        self.mode_in = "vector" if isinstance(input_dim, int) else "image"
        self.mode_out = "vector" if isinstance(input_dim, int) else "image"
        assert self.input_dim_total >= self.output_dim_total, "Input dimension has to be larger than output dimension"

    def forward(self, inputs, **kwargs):
        """(u, rest): two views of the flattened input, no copies (:25-35)."""
        h = inputs.reshape(inputs.size(0), -1)
        return h[:, : self.output_dim_total], h[:, self.output_dim_total :]

    def inverse(self, inputs, **kwargs):
        """Concatenate with the orthogonal coordinates (zeros by default) and restore the data shape (:37-47)."""
        orthogonal_inputs = kwargs.get("orthogonal_inputs", None)
        if orthogonal_inputs is None:
            x = F.pad(inputs, (0, self.input_dim_total - self.output_dim_total))
        else:
            x = torch.cat((inputs, orthogonal_inputs), dim=1)
        if self.mode_in == "image":
            x = x.view(inputs.size(0), *self.input_dim)
        return x


class Projection(ProjectionSplit):
    def forward(self, inputs, **kwargs):
        return super().forward(inputs)[0]
