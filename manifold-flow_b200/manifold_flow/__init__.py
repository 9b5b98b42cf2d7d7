"""Drop-in ``manifold_flow`` package for B200: same module paths, class names, constructor
signatures and ``state_dict`` keys as johannbrehmer/manifold-flow (reference
``manifold_flow/__init__.py:1``), with the neural-spline coupling hot path executed by the
hand-written sm_100a kernels of ``libmflow_b200.so`` (``manifold_flow.b200``)."""
from . import distributions, flows, nn, transforms  # noqa: F401
