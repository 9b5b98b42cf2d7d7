from .base import BaseFlow  # noqa: F401
from .flow import Flow  # noqa: F401
from .manifold_flow import ManifoldFlow  # noqa: F401
from .encoder_manifold_flow import EncoderManifoldFlow  # noqa: F401


def _out_of_scope(name):
    class _Stub:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"manifold_flow.flows.{name} is another algorithm, outside the B200 hot-path scope (SURVEY.md section 2, row 19)")

    _Stub.__name__ = _Stub.__qualname__ = name
    return _Stub


VariableDimensionManifoldFlow = _out_of_scope("VariableDimensionManifoldFlow")
ProbabilisticAutoEncoder = _out_of_scope("ProbabilisticAutoEncoder")
