"""``Linear`` base class lives in lu.py; ``NaiveLinear`` is dead code in the reference
(linear.py:135-239 calls the removed ``torch.gesv``) and is not provided."""
from .lu import Linear  # noqa: F401


class NaiveLinear:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("NaiveLinear is unreachable in the reference (uses removed torch.gesv); out of scope")
