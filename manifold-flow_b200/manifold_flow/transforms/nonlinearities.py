"""LogTanh (reference: manifold_flow/transforms/nonlinearities.py:38-95) - the only elementwise
bijection the BASELINE configs reach (image post-processing, image_transforms.py:251)."""
import numpy as np

from ..b200 import ops
from .base import Transform


class LogTanh(Transform):
    """tanh inside [-cut, cut], +-alpha*log(+-beta*x) outside, value- and slope-matched at the cut."""

    def __init__(self, cut_point=1):
        if cut_point <= 0:
            raise ValueError("Cut point must be positive.")
        super().__init__()
        self.cut_point = cut_point
        self.inv_cut_point = np.tanh(cut_point)
        self.alpha = (1 - np.tanh(np.tanh(cut_point))) / cut_point
        self.beta = np.exp((np.tanh(cut_point) - self.alpha * np.log(cut_point)) / self.alpha)

    def _apply_transform(self, inputs, context, inverse):
        return ops.logtanh(inputs, inverse=inverse, cut_point=float(self.cut_point))
