"""Transform applied to a masked subset of the features
(reference: manifold_flow/transforms/partial.py:11-88)."""
import torch

from .base import Transform


class PartialTransform(Transform):
    def __init__(self, mask, transform):
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
        self.transform = transform

    @property
    def num_identity_features(self):
        return len(self.identity_features)

    @property
    def num_transform_features(self):
        return len(self.transform_features)

    def _apply_transform(self, inputs, context, inverse):
        if inputs.shape[1] != self.features:
            raise ValueError("Expected features = {}, got {}.".format(self.features, inputs.shape[1]))
        part = torch.index_select(inputs, 1, self.transform_features)
        part, lad = self.transform._apply_transform(part, context, inverse)
        outputs = inputs.index_copy(1, self.transform_features, part)
        return outputs, lad
